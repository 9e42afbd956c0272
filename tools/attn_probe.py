import json, os, sys, math
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
B, h, T = 2, 5, 4096; C = h * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B * T, 3 * C, device=dev, generator=g) * 1.5).to(bf16)
out = ops.attention(qkv, qkv, qkv, batch=B, heads=h, t_q=T, t_kv=T, scale=0.125, col0_k=C, col0_v=2 * C)
q = qkv.float().view(B, T, 3, h, 64)
ref = F.scaled_dot_product_attention(q[:, :, 0].transpose(1, 2), q[:, :, 1].transpose(1, 2), q[:, :, 2].transpose(1, 2)).transpose(1, 2).reshape(B * T, C)
err = rel(out.float(), ref)
B = 8
qkv = torch.randn(B * T, 3 * C, device=dev).to(bf16); o = torch.empty(B * T, C, dtype=bf16, device=dev)
fn = lambda: ops.attention(qkv, qkv, qkv, o, batch=B, heads=h, t_q=T, t_kv=T, scale=0.125, col0_k=C, col0_v=2 * C)
for _ in range(3): fn()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): fn()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(json.dumps({"poly": os.environ.get("IDB_ATTN_POLY", "3"), "variant": os.environ.get("IDB_ATTN_VARIANT", "auto"), "packed": os.environ.get("IDB_ATTN_PACKED", "1"), "rel_err": err, "ms_T4096_B8": round(ms, 4),
                  "tflops": round(4.0 * B * h * T * T * 64 / ms / 1e9, 1)}))
