"""LayerNorm (fp32 token matrix -> bf16 operand) at the UNet's three widths, CUDA-graph timed.  (A variant
with 2 - 4 rows per warp was slower, 17.8 vs 12.9 us at 32768 x 320: registers cost more occupancy than the extra loads in
flight bring; one row per warp streams at 4.9 TB/s.)  Buffer sets larger than L2 alternate so the reads come from HBM.  usage: python tools/ln_probe.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for rows, C in [(32768, 320), (8192, 640), (2048, 1280), (512, 1280)]:
    nbuf = max(2, int(300e6 // (rows * C * 4)))
    xs = [torch.randn(rows, C, device=dev) for _ in range(nbuf)]
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    outs = [torch.empty(rows, C, dtype=torch.bfloat16, device=dev) for _ in range(nbuf)]
    for i in range(nbuf):
        ops.layernorm(xs[i], g, b, out=outs[i])
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(nbuf):
            ops.layernorm(xs[i], g, b, out=outs[i])
    gr.replay()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        gr.replay()
    e.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(e) / (5 * nbuf) * 1e3
    ref = torch.nn.functional.layer_norm(xs[0], (C,), g, b, 1e-5)
    err = float((outs[0].float() - ref).norm() / ref.norm())
    print(json.dumps({"rows": rows, "C": C, "us": round(us, 2), "GBps": round(rows * C * 6 / us / 1e3, 0), "rel_err": round(err, 5)}), flush=True)
