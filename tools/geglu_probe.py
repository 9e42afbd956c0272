"""The GEGLU FF-in GEMMs of the transformer blocks ([M, 2 * 4C] <- [M, C], epilogue a * gelu(g) -> bf16 [M, 4C]) back to back
in a CUDA graph; with IDB_GEMM_DEBUG (4 = TMEM read only, 5 = no epilogue, 6 = no TMA stores; any value selects the generic
epilogue, so compare against IDB_GEMM_NOSPEC=1 rather than the default) the pieces of the epilogue can be switched off.
usage: [IDB_GEMM_DEBUG=k | IDB_GEMM_NOSPEC=1] python tools/geglu_probe.py"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops  # noqa: E402
from faceposegenerator_b200.packing import interleave_geglu  # noqa: E402

dev = torch.device("cuda:0")
bf16 = torch.bfloat16


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


out = {"debug": os.environ.get("IDB_GEMM_DEBUG", "0"), "nospec": os.environ.get("IDB_GEMM_NOSPEC", "0")}
for (M, C) in [(32768, 320), (8192, 640), (2048, 1280)]:
    x = torch.randn(M, C, device=dev).to(bf16)
    w = (torch.randn(8 * C, C, device=dev) / math.sqrt(C)).to(bf16)
    b = torch.randn(8 * C, device=dev)
    wi, bi = interleave_geglu(w, b)
    o = torch.empty(M, 4 * C, dtype=bf16, device=dev)
    us = timeit(lambda: ops.gemm_conv(x, wi, bias=bi, geglu=True, out_bf16=o))
    out[f"{M}x{8 * C}x{C}"] = {"us": round(us, 2), "tflops": round(2.0 * M * 8 * C * C / us / 1e6, 1)}
print(json.dumps(out))
