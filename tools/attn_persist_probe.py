"""Row-split 256-query attention kernel: device time per launch at the UNet's self-attention shapes.
Run once with IDB_ATTN_PERSISTENT=0 and once with =1 (the switch is read once per process).
usage: python tools/attn_persist_probe.py"""
import json, os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"persistent": os.environ.get("IDB_ATTN_PERSISTENT", "1")}
for (B, h, T) in [(8, 5, 4096), (8, 10, 1024), (16, 5, 9216), (16, 10, 2304), (32, 5, 4096)]:
    C = h * 64
    qkv = torch.randn(B * T, 3 * C, device=dev).to(bf16); o = torch.empty(B * T, C, dtype=bf16, device=dev)
    fn = lambda: ops.attention(qkv, qkv, qkv, o, batch=B, heads=h, t_q=T, t_kv=T, scale=0.125, col0_k=C, col0_v=2 * C)
    for _ in range(3): fn()
    n = 10
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    ms = ts[n // 2]
    if B * T <= 8 * 4096:
        q = qkv.float().view(B, T, 3, h, 64)
        ref = F.scaled_dot_product_attention(q[:, :, 0].transpose(1, 2), q[:, :, 1].transpose(1, 2), q[:, :, 2].transpose(1, 2)).transpose(1, 2).reshape(B * T, C)
        err = float((o.float() - ref).norm() / ref.norm())
    else:
        err = None
    res[f"B{B}h{h}T{T}"] = {"ms": round(ms, 4), "tflops": round(4.0 * B * h * T * T * 64 / ms / 1e9, 1), "rel_err": err}
print(json.dumps(res))
