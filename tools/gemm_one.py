import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
M, K, N = 32768, 320, 320
x = torch.randn(M, K, device=dev).to(bf16)
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf16)
bias = torch.randn(N, device=dev)
out = torch.empty(M, N, dtype=bf16, device=dev)
for _ in range(5):
    ops.gemm_conv(x, w, bias=bias, out_bf16=out, k_splits=1)
torch.cuda.synchronize()
