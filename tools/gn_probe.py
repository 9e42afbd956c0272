"""GroupNorm apply-kernel timing by statistics source (CUDA events, mean of 50 launches after warm-up):
own = statistics pass + apply, rb = row-block sums (finalize kernel + apply), sums = fixed-point per-image sums (apply only).
usage: python tools/gn_probe.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=20):
    """device time per call: n calls captured in one CUDA graph (no host launch overhead), 10 replays"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (10 * n) * 1e3


import os as _os
SHAPES = [(4, 262144, 128, 0), (4, 65536, 256, 0), (4, 16384, 512, 0)] if _os.environ.get('GN_PROBE_VAE') else None
for (B, hw, c0, c1) in SHAPES or [(8, 4096, 320, 0), (8, 4096, 640, 320), (8, 1024, 640, 0), (8, 1024, 1280, 640), (8, 256, 1280, 1280), (8, 64, 1280, 1280)]:
    x0 = torch.randn(B, hw, c0, device=dev)
    x1 = torch.randn(B, hw, c1, device=dev) if c1 else None
    C = c0 + c1
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    ws = ops.groupnorm_workspace(B, 32, dev)
    out = torch.empty(B, hw, C, dtype=torch.bfloat16, device=dev)

    def rb(x):
        v = x.view(B * hw // 32, 32, -1)
        return torch.stack([v.sum(1), (v * v).sum(1)], -1).contiguous()

    import math
    gran = math.gcd(c0, c1 or c0) // 32

    def fx(x):
        d = x.double().view(B, hw, -1, gran)
        return torch.stack([(d.sum((1, 3)) * 2.0 ** 32).round().long(), ((d * d).sum((1, 3)) * 2.0 ** 24).round().long()], -1).contiguous()
    rb0, rb1 = rb(x0), (rb(x1) if c1 else None)
    s0, s1 = (fx(x0), (fx(x1) if c1 else None)) if hw <= 16384 else (rb0, rb1)
    kw = dict(groups=32, eps=1e-5, silu=True, x1=x1, out_norm=out, partials=ws)
    res = {"B": B, "hw": hw, "c0": c0, "c1": c1,
           "own_us": timeit(lambda: ops.groupnorm(x0, gamma, beta, **kw)),
           "rb_us": timeit(lambda: ops.groupnorm(x0, gamma, beta, x0_stats=rb0, x1_stats=rb1, **kw)),
           "sums_us": timeit(lambda: ops.groupnorm(x0, gamma, beta, x0_stats=s0, x1_stats=s1, **kw))}
    y_own = ops.groupnorm(x0, gamma, beta, **kw)[0].clone()
    y_sums = ops.groupnorm(x0, gamma, beta, x0_stats=s0, x1_stats=s1, **kw)[0]
    res["max_diff"] = float((y_own.float() - y_sums.float()).abs().max())
    res["GBps_sums"] = B * hw * C * 6 / res["sums_us"] / 1e3
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
