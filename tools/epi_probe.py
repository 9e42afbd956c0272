"""Epilogue-bound GEMM shapes of the 64x64 transformer blocks, timed inside a CUDA graph with
rotating buffer sets (working set > L2, as inside the UNet step)."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16; f32 = torch.float32
R = 4
def graph_time(fns, n=24):
    for f in fns: f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n): fns[i % len(fns)]()
    for _ in range(2): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n * 1e3, 1)
res = {"tag": os.environ.get("IDB_TAG", "") + f" DEBUG={os.environ.get('IDB_GEMM_DEBUG','0')}"}
sel = os.environ.get("IDB_CASES", "").split(",") if os.environ.get("IDB_CASES") else None
def case(name, M, K, N, *, f32out=False, b16out=False, res_=False, stats=False, lora=0, geglu=False, bias=True):
    if sel and name not in sel: return
    n_out = N // 2 if geglu else N
    fns = []
    for r in range(R):
        x = torch.randn(M, K, device=dev).to(bf16)
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf16)
        kw = {}
        if bias: kw["bias"] = torch.randn(N, device=dev)
        if res_: kw["residual"] = torch.randn(M, n_out, device=dev)
        if f32out: kw["out_f32"] = torch.empty(M, n_out, dtype=f32, device=dev)
        if b16out: kw["out_bf16"] = torch.empty(M, n_out, dtype=bf16, device=dev)
        if stats: kw["stats"] = torch.empty((M + 31) // 32, n_out, 2, dtype=f32, device=dev)
        if lora:
            kw["lora_down"] = torch.randn(16 * lora, K, device=dev).to(bf16)
            up = torch.zeros(N, 64, device=dev); up[:, :4] = torch.randn(N, 4, device=dev) * 0.05
            kw["lora_up"] = up.to(bf16)
            kw["lora_seg_n"] = N // lora
        if geglu: kw["geglu"] = True
        fns.append(lambda x=x, w=w, kw=kw: ops.gemm_conv(x, w, **kw))
    us = graph_time(fns)
    res[name] = [us, round(2.0 * M * K * N / us / 1e6)]
case("proj_in", 32768, 320, 320, f32out=True)
case("proj_out", 32768, 320, 320, f32out=True, res_=True, stats=True)
case("res_nostats", 32768, 320, 320, f32out=True, res_=True)
case("o1_lora", 32768, 320, 320, f32out=True, res_=True, lora=1)
case("q2_lora", 32768, 320, 320, b16out=True, lora=1, bias=False)
case("qkv_lora", 32768, 320, 960, b16out=True, lora=3, bias=False)
case("qkv_plain", 32768, 320, 960, b16out=True, bias=False)
case("ff1", 32768, 320, 2560, b16out=True, geglu=True)
case("ff1_nogeglu", 32768, 320, 2560, b16out=True)
case("ff2", 32768, 1280, 320, b16out=True, res_=True)
case("m_o1_lora", 8192, 640, 640, f32out=True, res_=True, lora=1)
case("m_ff1", 8192, 640, 5120, b16out=True, geglu=True)
case("conv320", 32768, 2880, 320, f32out=True, stats=True) if False else None
print(json.dumps(res))
