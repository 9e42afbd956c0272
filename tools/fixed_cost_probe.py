"""Fixed cost of one GEMM launch: the smallest transformer linear (M = 2048, N = K = 1280, one tile per CTA pair) back to
back in a CUDA graph, with parts of the kernel switched off (IDB_GEMM_DEBUG: 1 = no MMA issue, 2 = no TMA loads, 3 = both,
5 = no epilogue (no TMEM read / stores), 6 = no TMA stores), and a trivial kernel (cast of 1 KB) for the launch floor.
usage: IDB_GEMM_DEBUG=k python tools/fixed_cost_probe.py"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf16, f32 = torch.bfloat16, torch.float32


def timeit(fn, n=40):
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
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * n) * 1e3


tiny = torch.randn(256, device=dev)
out = {"debug": os.environ.get("IDB_GEMM_DEBUG", "0"), "trivial_kernel_us": round(timeit(lambda: ops.cast_bf16(tiny)), 2)}
for (M, N, K) in [(2048, 1280, 1280), (2048, 1280, 320), (8192, 640, 640), (32768, 320, 320)]:
    x = torch.randn(M, K, device=dev).to(bf16)
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf16)
    of = torch.empty(M, N, dtype=f32, device=dev)
    ob = torch.empty(M, N, dtype=bf16, device=dev)
    out[f"{M}x{N}x{K}_f32"] = round(timeit(lambda: ops.gemm_conv(x, w, out_f32=of)), 2)
    out[f"{M}x{N}x{K}_bf16"] = round(timeit(lambda: ops.gemm_conv(x, w, out_bf16=ob)), 2)
print(json.dumps(out))
