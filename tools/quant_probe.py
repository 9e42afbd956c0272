"""Wave quantisation probe: the UNet's conv GEMMs and the T = 4096 self-attention at B = 8 (the bench batch: 64 / 128 work
items on 74 CTA pairs, 640 on 148 SMs) against B = 9 / 10 (more items, same number of waves).  If the time does not move
between B = 8 and B = 9, the idle SMs of the last wave are free capacity (what a stream-K schedule could use); if it grows
with the work, the kernels are bound by something chip-wide (power, L2 feed).  usage: python tools/quant_probe.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=10):
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


only = int(sys.argv[1]) if len(sys.argv) > 1 else 0     # e.g. 16: only the 16x16 convs at B = 8
ws = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
for (hw, cin, cout) in [(16, 1280, 1280), (16, 2560, 1280), (32, 640, 640), (32, 1280, 640), (64, 320, 320), (64, 640, 320), (8, 1280, 1280)]:
    if only and hw != only:
        continue
    w = (torch.randn(cout, 9 * cin, device=dev) * 0.02).bfloat16()
    for B in ((8,) if only else (8, 9, 10, 16)):
        x = torch.randn(B, hw, hw, cin, device=dev).bfloat16()
        res = torch.randn(B * hw * hw, cout, device=dev)
        out = torch.empty(B * hw * hw, cout, device=dev)
        fn = lambda: ops.gemm_conv(x, w, mode=ops.A_3X3, residual=res, out_f32=out, workspace=ws, k_splits=0)
        us = timeit(fn)
        fl = 2.0 * B * hw * hw * cout * 9 * cin
        print(json.dumps({"conv3x3": [hw, cin, cout], "B": B, "us": round(us, 2), "tflops": round(fl / us / 1e6, 1),
                          "us_per_image": round(us / B, 3)}), flush=True)

for (h, T) in ([] if only else [(5, 4096), (10, 1024)]):
    C = h * 64
    for B in (8, 9, 10, 16):
        qkv = torch.randn(B * T, 3 * C, device=dev).bfloat16()
        out = torch.empty(B * T, C, dtype=torch.bfloat16, device=dev)
        fn = lambda: ops.attention(qkv, qkv, qkv, out, batch=B, heads=h, t_q=T, t_kv=T, scale=0.125, col0_k=C, col0_v=2 * C)
        us = timeit(fn)
        print(json.dumps({"attention": [h, T], "B": B, "items_256q": B * h * T // 256, "us": round(us, 2),
                          "tflops": round(4.0 * B * h * T * T * 64 / us / 1e6, 1), "us_per_image": round(us / B, 3)}), flush=True)

from faceposegenerator_b200 import _lib  # noqa: E402
print(json.dumps({"stream_k_launches": int(_lib.load().idb_stream_k_launch_count()), "stream_k_mode": int(_lib.load().idb_stream_k_mode())}))
