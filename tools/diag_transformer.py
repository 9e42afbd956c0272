"""Localise run-to-run divergence inside the first transformer block, with memory poisoning."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops  # noqa: E402
from faceposegenerator_b200.unet import UNet2DConditionModel, HEAD_DIM  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def poison():
    big = torch.full((6 << 30,), 0xFF, dtype=torch.uint8, device=dev)  # 0xFFFF bf16 = NaN, fp32 0xFFFFFFFF = NaN
    del big


unet = UNet2DConditionModel.from_random(0, device=dev)
t = unet.transformers[0]
g = torch.Generator().manual_seed(1)
h = (torch.randn(2, 64, 64, 320, generator=g) * 1.2).to(dev)
ctx = torch.randn(2 * 77, 1024, generator=g).to(dev).to(torch.bfloat16)


def run():
    taps = {}
    gnws = ops.groupnorm_workspace(2, 32, dev)
    kv = unet._context_kv(t, ctx)
    taps["kv"] = kv.float().clone()
    B, H, W, Cc = h.shape
    M, T = B * H * W, H * W
    n, _ = ops.groupnorm(h, t.gn_g, t.gn_b, groups=32, eps=1e-6, silu=False, partials=gnws)
    taps["gn"] = n.float().clone()
    x0, _ = unet._gemm(n.view(M, Cc), t.w_in, bias=t.b_in, want_f32=True)
    taps["x0"] = x0.clone()
    a = ops.layernorm(x0, *t.ln[0])
    taps["ln1"] = a.float().clone()
    _, qkv = unet._lin_lora(a, t.w_qkv, t.lora["qkv"], Cc, want_bf16=True)
    taps["qkv"] = qkv.float().clone()
    o = ops.attention(qkv, qkv, qkv, batch=B, heads=t.heads, t_q=T, t_kv=T, scale=HEAD_DIM ** -0.5, col0_q=0, col0_k=Cc, col0_v=2 * Cc)
    taps["attn1"] = o.float().clone()
    x1, _ = unet._lin_lora(o, t.w_o1, t.lora["o1"], Cc, bias=t.b_o1, residual=x0, want_f32=True)
    taps["x1"] = x1.clone()
    a = ops.layernorm(x1, *t.ln[1])
    _, q = unet._lin_lora(a, t.w_q2, t.lora["q2"], Cc, want_bf16=True)
    taps["q2"] = q.float().clone()
    o = ops.attention(q, kv, kv, batch=B, heads=t.heads, t_q=T, t_kv=77, scale=HEAD_DIM ** -0.5, col0_q=0, col0_k=0, col0_v=Cc)
    taps["attn2"] = o.float().clone()
    x2, _ = unet._lin_lora(o, t.w_o2, t.lora["o2"], Cc, bias=t.b_o2, residual=x1, want_f32=True)
    taps["x2"] = x2.clone()
    a = ops.layernorm(x2, *t.ln[2])
    _, gg = unet._gemm(a, t.w_ff1, bias=t.b_ff1, geglu=True, want_bf16=True)
    taps["geglu"] = gg.float().clone()
    _, x3 = unet._gemm(gg, t.w_ff2, bias=t.b_ff2, residual=x2, want_bf16=True)
    taps["x3"] = x3.float().clone()
    out, _ = unet._gemm(x3, t.w_out, bias=t.b_out, residual=h.view(M, Cc), want_f32=True)
    taps["out"] = out.clone()
    torch.cuda.synchronize()
    return taps


a = run()
poison()
b = run()
for k in a:
    nan = bool(torch.isnan(b[k]).any())
    print(f"{k}: repeat diff {rel(b[k], a[k]):.3e} nan={nan}", flush=True)

# fp32 torch reference of the cross attention on the same operands
import torch.nn.functional as F
kvf = a["kv"].view(2, 77, 2, 5, 64)
qf = a["q2"].view(2, 4096, 5, 64)
ref = F.scaled_dot_product_attention(qf.transpose(1, 2), kvf[:, :, 0].transpose(1, 2), kvf[:, :, 1].transpose(1, 2))
print("attn2 vs sdpa:", rel(a["attn2"], ref.transpose(1, 2).reshape(8192, 320)))
q = a["qkv"].view(2, 4096, 3, 5, 64)
ref = F.scaled_dot_product_attention(q[:, :, 0].transpose(1, 2), q[:, :, 1].transpose(1, 2), q[:, :, 2].transpose(1, 2))
print("attn1 vs sdpa:", rel(a["attn1"], ref.transpose(1, 2).reshape(8192, 320)))
