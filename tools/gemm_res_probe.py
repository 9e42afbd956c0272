"""The HBM-bound residual-stream GEMMs of the transformer blocks (to_out + fused LoRA + fp32 residual; proj_out + residual +
GroupNorm sums; proj_in), CUDA-graph timed (20 launches per replay, device time per launch) with their HBM floor.
usage: python tools/gemm_res_probe.py"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf16, f32 = torch.bfloat16, torch.float32
WS = torch.empty((96 << 20) // 4, dtype=f32, device=dev)
HBM = 6545.0


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


for (M, C, T) in [(32768, 320, 4096), (8192, 640, 1024), (2048, 1280, 256)]:
    x = torch.randn(M, C, device=dev).to(bf16)
    w = (torch.randn(C, C, device=dev) / math.sqrt(C)).to(bf16)
    bias = torch.randn(C, device=dev)
    res = torch.randn(M, C, device=dev)
    ld = torch.randn(16, C, device=dev).to(bf16)
    lu = torch.nn.functional.pad(torch.randn(C, 4, device=dev) * 0.05, (0, 60)).to(bf16)
    of, ob = torch.empty(M, C, dtype=f32, device=dev), torch.empty(M, C, dtype=bf16, device=dev)
    pool = ops.SumsPool(dev, capacity=1 << 22)
    cases = {
        "to_out lora f32 res": (lambda: ops.gemm_conv(x, w, bias=bias, residual=res, lora_down=ld, lora_up=lu, lora_seg_n=C, out_f32=of, k_splits=0, workspace=WS), 2 + 4 + 4),
        "q2 lora bf16": (lambda: ops.gemm_conv(x, w, lora_down=ld, lora_up=lu, lora_seg_n=C, out_bf16=ob, k_splits=0, workspace=WS), 2 + 2),
        "proj_out f32 res sums": (lambda: ops.gemm_conv(x, w, bias=bias, residual=res, out_f32=of, want_stats=True, stats_hw=T, stats_gran=10, sums_pool=pool, k_splits=0, workspace=WS), 2 + 4 + 4),
        "proj_in f32": (lambda: ops.gemm_conv(x, w, bias=bias, out_f32=of, k_splits=0, workspace=WS), 2 + 4),
        "plain f32 res (no lora)": (lambda: ops.gemm_conv(x, w, bias=bias, residual=res, out_f32=of, k_splits=0, workspace=WS), 2 + 4 + 4),
    }
    for name, (fn, bpe) in cases.items():
        us = timeit(fn)
        floor = M * C * bpe / HBM / 1e3
        print(json.dumps({"M": M, "C": C, "case": name, "us": round(us, 2), "tflops": round(2.0 * M * C * C / us / 1e6, 1),
                          "hbm_floor_us": round(floor, 2), "frac_of_hbm_floor": round(floor / us, 3)}), flush=True)
