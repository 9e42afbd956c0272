"""Feed-rate / issue-rate probes for gemm_tc (IDB_GEMM_DEBUG / IDB_GEMM_CG set by the caller's env)."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=8, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
tag = f"CG={os.environ.get('IDB_GEMM_CG','auto')} DEBUG={os.environ.get('IDB_GEMM_DEBUG','0')}"
for (B, H, C, N) in [(8, 64, 640, 640), (8, 64, 320, 320), (8, 32, 1280, 1280), (8, 16, 1280, 1280)]:
    x = torch.randn(B, H, H, C, device=dev).to(bf16)
    w = (torch.randn(N, 9 * C, device=dev) / math.sqrt(9 * C)).to(bf16)
    out = torch.empty(B * H * H, N, dtype=torch.float32, device=dev)
    ms = timeit(lambda: ops.gemm_conv(x, w, mode=ops.A_3X3, out_f32=out, k_splits=1))
    print(json.dumps({"tag": tag, "case": f"conv B{B} {H}x{H} {C}->{N}", "ms": round(ms, 4),
                      "tflops_equiv": round(2.0 * B * H * H * N * 9 * C / ms / 1e9, 1)}), flush=True)
for (M, K, N) in [(32768, 1280, 2560), (32768, 2560, 1280), (8192, 5120, 5120)]:
    x = torch.randn(M, K, device=dev).to(bf16)
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf16)
    out = torch.empty(M, N, dtype=bf16, device=dev)
    ms = timeit(lambda: ops.gemm_conv(x, w, out_bf16=out, k_splits=1))
    print(json.dumps({"tag": tag, "case": f"gemm M{M} K{K} N{N}", "ms": round(ms, 4),
                      "tflops_equiv": round(2.0 * M * K * N / ms / 1e9, 1)}), flush=True)
