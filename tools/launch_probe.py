"""Fixed per-launch cost of the kernels inside a CUDA graph (tiny problems, 100 launches per graph)."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
def graph_time(fn, n=100):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
x = torch.randn(256, 64, device=dev).to(bf16); w = torch.randn(128, 64, device=dev).to(bf16)
o = torch.empty(256, 128, dtype=bf16, device=dev)
xs = torch.randn(128, 64, device=dev).to(bf16); os_ = torch.empty(128, 128, dtype=bf16, device=dev)
f = torch.randn(4096, device=dev); fb = torch.empty(4096, dtype=bf16, device=dev)
res = {"gemm_cg2_tiny_us": graph_time(lambda: ops.gemm_conv(x, w, out_bf16=o)),
       "gemm_cg1_tiny_us": graph_time(lambda: ops.gemm_conv(xs, w, out_bf16=os_)),
       "cast_tiny_us": graph_time(lambda: ops.cast_bf16(f, fb))}
xm = torch.randn(32768, 320, device=dev).to(bf16); wm = torch.randn(320, 320, device=dev).to(bf16)
om = torch.empty(32768, 320, dtype=bf16, device=dev)
res["gemm_proj320_us"] = graph_time(lambda: ops.gemm_conv(xm, wm, out_bf16=om))
qkv = torch.randn(2 * 64, 3 * 1280, device=dev).to(bf16)
res["attn_tiny_us"] = graph_time(lambda: ops.attention(qkv, qkv, qkv, batch=2, heads=20, t_q=64, t_kv=64, scale=0.125, col0_k=1280, col0_v=2560))
xg = torch.randn(2, 64, 1280, device=dev); gm = torch.ones(1280, device=dev); bt = torch.zeros(1280, device=dev)
og = torch.empty(2, 64, 1280, dtype=bf16, device=dev); ws = ops.groupnorm_workspace(2, 32, dev)
res["gn_tiny_us(2 kernels)"] = graph_time(lambda: ops.groupnorm(xg, gm, bt, groups=32, eps=1e-5, silu=True, out_norm=og, partials=ws))
print(json.dumps(res))
