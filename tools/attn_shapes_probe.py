"""Self-attention at the UNet's smaller rasters (T = 1024 / 256 / 64, B = 8), CUDA-graph timed, per kernel variant
(IDB_ATTN_VARIANT = 128 / 256 forces the 128-query kernel / the row-split kernel).  usage: python tools/attn_shapes_probe.py"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=20):
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


for (B, h, T) in [(8, 5, 4096), (8, 10, 1024), (8, 20, 256), (8, 20, 64), (2, 5, 4096), (2, 10, 1024), (2, 20, 256)]:
    C = h * 64
    qkv = torch.randn(B * T, 3 * C, device=dev).bfloat16()
    out = torch.empty(B * T, C, dtype=torch.bfloat16, device=dev)
    fn = lambda: ops.attention(qkv, qkv, qkv, out, batch=B, heads=h, t_q=T, t_kv=T, scale=0.125, col0_k=C, col0_v=2 * C)
    us = timeit(fn)
    q, k, v = qkv.float().view(B, T, 3, h, 64).unbind(2)
    ref = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2).reshape(B * T, C)
    err = float((out.float() - ref).norm() / ref.norm())
    print(json.dumps({"variant": os.environ.get("IDB_ATTN_VARIANT", "auto"), "B": B, "heads": h, "T": T, "us": round(us, 2),
                      "tflops": round(4.0 * B * h * T * T * 64 / us / 1e6, 1), "rel_err": round(err, 5)}), flush=True)
    if T >= 256:   # the cross-attention over the 77-token text context at the same raster
        kv = torch.randn(B * 77, 2 * C, device=dev).bfloat16()
        fx = lambda: ops.attention(qkv, kv, kv, out, batch=B, heads=h, t_q=T, t_kv=77, scale=0.125, col0_q=0, col0_k=0, col0_v=C)
        us = timeit(fx)
        kx, vx = kv.float().view(B, 77, 2, h, 64).unbind(2)
        refx = F.scaled_dot_product_attention(q.transpose(1, 2), kx.transpose(1, 2), vx.transpose(1, 2)).transpose(1, 2).reshape(B * T, C)
        err = float((out.float() - refx).norm() / refx.norm())
        print(json.dumps({"variant": os.environ.get("IDB_ATTN_VARIANT", "auto"), "cross": True, "B": B, "heads": h, "T": T, "us": round(us, 2),
                          "hbm_floor_us": round(B * T * C * 4 / 6545e3, 2), "rel_err": round(err, 5)}), flush=True)
