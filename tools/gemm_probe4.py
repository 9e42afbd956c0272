"""LoRA GEMM probes inside a CUDA graph (50 launches)."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
def graph_time(fn, n=50):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n * 1e3, 2)
tag = f"DEBUG={os.environ.get('IDB_GEMM_DEBUG','0')}"
res = {"tag": tag}
for (M, K, C, nseg) in [(32768, 320, 320, 1), (32768, 320, 320, 3), (8192, 640, 640, 3), (2048, 1280, 1280, 3)]:
    N = C * nseg
    x = torch.randn(M, K, device=dev).to(bf16)
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf16)
    ld = torch.randn(16 * nseg, K, device=dev).to(bf16)
    lu = torch.randn(N, 4, device=dev) * 0.05
    out = torch.empty(M, N, dtype=bf16, device=dev)
    res[f"lora M{M} K{K} N{N}"] = graph_time(lambda: ops.gemm_conv(x, w, lora_down=ld, lora_up=lu, lora_seg_n=C, out_bf16=out))
    res[f"plain M{M} K{K} N{N}"] = graph_time(lambda: ops.gemm_conv(x, w, out_bf16=out))
print(json.dumps(res))
