"""UNet step timing (eager vs CUDA graph) at several batch sizes; prints JSON lines."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import _lib  # noqa: E402
from faceposegenerator_b200.unet import UNet2DConditionModel  # noqa: E402
from faceposegenerator_b200.weights import random_lora  # noqa: E402

FLOP_PER_ROW = 0.8054e12
dev = torch.device("cuda:0")
t0 = time.time()
unet = UNet2DConditionModel.from_random(0, device=dev)
unet.set_lora(random_lora(seed=0))
print(json.dumps({"build_s": round(time.time() - t0, 1)}), flush=True)
for B in [int(a) for a in sys.argv[1:]] or [2, 8]:
    x = torch.randn(B, 4, 64, 64, device=dev)
    ctx = torch.randn(B, 77, 1024, device=dev)
    t = torch.full((B,), 500.0, device=dev)
    context = unet.encode_context(ctx)
    temb = unet.time_embedding(t) if os.environ.get("IDB_BENCH_TEMB", "1") == "1" else None   # hoisted by the pipeline
    for _ in range(2):
        unet.forward(x, t, context=context, temb=temb)
    torch.cuda.synchronize()
    n0 = _lib.launch_count
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        unet.forward(x, t, context=context, temb=temb)
    b.record()
    torch.cuda.synchronize()
    eager = a.elapsed_time(b) / 5
    launches = (_lib.launch_count - n0) // 5
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = unet.forward(x, t, context=context, temb=temb, return_dict=False)[0]
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    gms = a.elapsed_time(b) / 10
    print(json.dumps({"B": B, "eager_ms": round(eager, 3), "graph_ms": round(gms, 3), "launches": launches,
                      "tflops_graph": round(B * FLOP_PER_ROW / gms / 1e9, 1),
                      "frac_of_1672.7": round(B * FLOP_PER_ROW / gms / 1e9 / 1672.7, 4)}), flush=True)
    del g
