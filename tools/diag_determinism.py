"""Diagnostics: run-to-run determinism of each kernel and of the UNet; localise base vs zero-LoRA divergence."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops  # noqa: E402
from faceposegenerator_b200.unet import UNet2DConditionModel  # noqa: E402
from faceposegenerator_b200.weights import random_lora  # noqa: E402

dev = torch.device("cuda:0")
bf16 = torch.bfloat16


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rep(name, fn, n=4):
    outs = [fn().clone() for _ in range(n)]
    torch.cuda.synchronize()
    print(f"{name}: max repeat diff {max(rel(o, outs[0]) for o in outs[1:]):.3e}", flush=True)


g = torch.Generator(device="cuda").manual_seed(0)
B, T, C = 2, 4096, 320
qkv = torch.randn(B * T, 3 * C, device=dev, generator=g).to(bf16)
rep("attention T4096", lambda: ops.attention(qkv, qkv, qkv, batch=B, heads=5, t_q=T, t_kv=T, scale=0.125, col0_k=C, col0_v=2 * C))
q = qkv.float().view(B, T, 3, 5, 64)
ref = torch.nn.functional.scaled_dot_product_attention(q[:, :, 0].transpose(1, 2), q[:, :, 1].transpose(1, 2), q[:, :, 2].transpose(1, 2))
out = ops.attention(qkv, qkv, qkv, batch=B, heads=5, t_q=T, t_kv=T, scale=0.125, col0_k=C, col0_v=2 * C)
print("attention vs sdpa", rel(out.float(), ref.transpose(1, 2).reshape(B * T, C)), flush=True)
kv = torch.randn(B * 77, 2 * C, device=dev, generator=g).to(bf16)
rep("attention S77", lambda: ops.attention(qkv, kv, kv, batch=B, heads=5, t_q=T, t_kv=77, scale=0.125, col0_v=C))
x = torch.randn(8192, 640, device=dev, generator=g).to(bf16)
w = (torch.randn(1920, 640, device=dev, generator=g) / 25).to(bf16)
rep("gemm", lambda: ops.gemm_conv(x, w, want_f32=True)[0])
ws = torch.empty(24 << 20, dtype=torch.float32, device=dev)
x2 = torch.randn(128, 1280, device=dev, generator=g).to(bf16)
w2 = (torch.randn(1280, 1280, device=dev, generator=g) / 36).to(bf16)
rep("gemm splitk auto", lambda: ops.gemm_conv(x2, w2, want_f32=True, k_splits=0, workspace=ws)[0])
a = ops.gemm_conv(x2, w2, want_f32=True, k_splits=0, workspace=ws)[0]
b = ops.gemm_conv(x2, w2, want_f32=True, k_splits=1)[0]
print("splitk auto vs none", rel(a, b), flush=True)
xi = torch.randn(2, 64, 64, 320, device=dev, generator=g).to(bf16)
wc = (torch.randn(320, 2880, device=dev, generator=g) / 54).to(bf16)
rep("conv", lambda: ops.gemm_conv(xi, wc, mode=ops.A_3X3, want_f32=True)[0])
xf = torch.randn(2, 4096, 320, device=dev, generator=g)
gm, bt = torch.ones(320, device=dev), torch.zeros(320, device=dev)
rep("groupnorm", lambda: ops.groupnorm(xf, gm, bt, groups=32, eps=1e-5, silu=True)[0].float())
rep("layernorm", lambda: ops.layernorm(xf.view(-1, 320), gm, bt).float())

unet = UNet2DConditionModel.from_random(0, device=dev)
lora = random_lora(seed=1)
gg = torch.Generator().manual_seed(123)
x = torch.randn(2, 4, 64, 64, generator=gg).to(dev)
ctx = torch.randn(2, 77, 1024, generator=gg).to(dev)
unet.set_lora(None)
t1, t2, t3 = {}, {}, {}
o1 = unet.forward(x, 500, ctx, return_dict=False, taps=t1)[0].clone()
o2 = unet.forward(x, 500, ctx, return_dict=False, taps=t2)[0].clone()
print("unet repeat (no lora):", rel(o1, o2), flush=True)
for k in t1:
    d = rel(t1[k], t2[k])
    if d > 0:
        print("  first repeat divergence at", k, d)
        break
zero = {k: (d, torch.zeros_like(u), s) for k, (d, u, s) in lora.items()}
unet.set_lora(zero)
o3 = unet.forward(x, 500, ctx, return_dict=False, taps=t3)[0].clone()
print("unet zero-lora vs none:", rel(o3, o1), flush=True)
for k in t1:
    print(f"  {k}: {rel(t3[k], t1[k]):.3e}")
