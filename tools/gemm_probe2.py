"""BN sweep: MMA-only / TMA-only / handshake-only / full per-k-block cost, with SM clock sampling."""
import json, math, os, subprocess, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
bn = int(os.environ.get("IDB_GEMM_BN", "256"))
M, K = 16384, 4096
N = bn * 40
x = torch.randn(M, K, device=dev).to(bf16)
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf16)
out = torch.empty(M, N, dtype=bf16, device=dev)
fn = lambda: ops.gemm_conv(x, w, out_bf16=out, k_splits=1)
for _ in range(3): fn()
torch.cuda.synchronize()
# run ~0.6 s continuously, sample the SM clock in the middle
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n_it = 200
a.record()
for _ in range(n_it): fn()
b.record()
time.sleep(0.15)
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / n_it
tiles = (M // 128) * 40
kb_per_cta = math.ceil(tiles / 148) * (K // 64)
print(json.dumps({"bn": bn, "cg": os.environ.get("IDB_GEMM_CG"), "debug": os.environ.get("IDB_GEMM_DEBUG", "0"), "ms": round(ms, 4),
                  "tflops_equiv": round(2.0 * M * K * N / ms / 1e9, 1), "ns_per_kblock": round(ms * 1e6 / kb_per_cta, 1),
                  "clk_mhz,power_w": clk}), flush=True)
