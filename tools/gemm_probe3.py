"""Small-K GEMM probes (epilogue-bound regime)."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=8, warm=3, do_flush=True):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        if do_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
tag = f"CG={os.environ.get('IDB_GEMM_CG','auto')} DEBUG={os.environ.get('IDB_GEMM_DEBUG','0')}"
for (M, K, N, geglu) in [(32768, 320, 320, False), (32768, 320, 960, False), (32768, 320, 2560, True), (8192, 640, 640, False)]:
    x = torch.randn(M, K, device=dev).to(bf16)
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf16)
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N // 2 if geglu else N, dtype=bf16, device=dev)
    fn = lambda: ops.gemm_conv(x, w, bias=bias, geglu=geglu, out_bf16=out, k_splits=1)
    ms = timeit(fn); ms_hot = timeit(fn, do_flush=False)
    # 20 back-to-back launches (amortises launch latency)
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"tag": tag, "case": f"M{M} K{K} N{N}{' geglu' if geglu else ''}", "ms_cold": round(ms, 4), "ms_hot": round(ms_hot, 4),
                      "ms_b2b": round(a.elapsed_time(b) / 20, 4)}), flush=True)
