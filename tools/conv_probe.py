"""Big-K conv GEMM throughput vs number of SMs used / debug modes (feed-limit experiments)."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16; f32 = torch.float32
def graph_time(fn, n=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    for _ in range(2): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
res = {"tag": f"MAXSM={os.environ.get('IDB_GEMM_MAXSM','0')} DEBUG={os.environ.get('IDB_GEMM_DEBUG','0')} BN={os.environ.get('IDB_GEMM_BN','0')}"}
for (B, H, C, N) in [(8, 64, 320, 320), (8, 32, 640, 640), (8, 16, 1280, 1280), (8, 32, 1280, 1280)]:
    x = torch.randn(B, H, H, C, device=dev).to(bf16)
    w = (torch.randn(N, 9 * C, device=dev) / math.sqrt(9 * C)).to(bf16)
    bias = torch.randn(N, device=dev)
    out = torch.empty(B * H * H, N, dtype=f32, device=dev)
    us = graph_time(lambda: ops.gemm_conv(x, w, mode=ops.A_3X3, bias=bias, out_f32=out))
    res[f"conv B{B} {H}x{H} {C}->{N}"] = [round(us, 1), round(2.0 * B * H * H * 9 * C * N / us / 1e6)]
print(json.dumps(res))
# cuBLAS reference on the same GEMM shapes (dense [M, 9C] x [9C, N]; an upper bound for the implicit-GEMM kernel)
if os.environ.get("IDB_CUBLAS_REF") == "1":
    ref = {"tag": "cuBLAS (torch.matmul) bf16, same shapes, dense A"}
    for (B, H, C, N) in [(8, 64, 320, 320), (8, 32, 640, 640), (8, 16, 1280, 1280), (8, 32, 1280, 1280)]:
        M, K = B * H * H, 9 * C
        a = torch.randn(M, K, device=dev).to(bf16); wt = torch.randn(K, N, device=dev).to(bf16)
        o = torch.empty(M, N, dtype=bf16, device=dev)
        us = graph_time(lambda: torch.matmul(a, wt, out=o))
        ref[f"gemm M{M} K{K} N{N}"] = [round(us, 1), round(2.0 * M * K * N / us / 1e6)]
    print(json.dumps(ref))
    # plain linear through our kernel on the same dense shapes
    mine = {"tag": "idb_gemm_conv as a Linear on the same dense shapes"}
    for (B, H, C, N) in [(8, 64, 320, 320), (8, 32, 640, 640), (8, 16, 1280, 1280), (8, 32, 1280, 1280)]:
        M, K = B * H * H, 9 * C
        a = torch.randn(M, K, device=dev).to(bf16); w2 = torch.randn(N, K, device=dev).to(bf16)
        o = torch.empty(M, N, dtype=bf16, device=dev)
        us = graph_time(lambda: ops.gemm_conv(a, w2, out_bf16=o))
        mine[f"gemm M{M} K{K} N{N}"] = [round(us, 1), round(2.0 * M * K * N / us / 1e6)]
    print(json.dumps(mine))
