"""LoRA-only training step timing at the config-5 batch (B = 4, 64x64 latents, no CFG): forward with tape,
backward, and the torch-eager comparator (autograd through the oracle module graph in bf16 on the same GPU).
usage: python tools/train_step_bench.py [B]"""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200.lora_backward import UNetLoRAGrad  # noqa: E402
from faceposegenerator_b200.unet import UNet2DConditionModel  # noqa: E402
from faceposegenerator_b200.weights import random_lora, random_state_dict, unet_manifest  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
sd = random_state_dict(unet_manifest(), 0)
lora = random_lora(seed=1)
unet = UNet2DConditionModel(sd, device=dev)
unet.set_lora(lora)
eng = UNetLoRAGrad(unet, lora)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 4, 64, 64, generator=g).to(dev)
ctx = torch.randn(B, 77, 1024, generator=g).to(dev)
t = torch.randint(0, 1000, (B,), generator=g).float().to(dev)
tgt = torch.randn(B, 4, 64, 64, generator=g).to(dev)


def step():
    eps = eng.forward(x, t, ctx)
    return eng.backward(2.0 * (eps - tgt) / eps.numel())


for _ in range(2):
    step()
torch.cuda.synchronize()
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
n = 5
fw = bw = 0.0
for _ in range(n):
    a.record()
    eps = eng.forward(x, t, ctx)
    b.record()
    eng.backward(2.0 * (eps - tgt) / eps.numel())
    c.record()
    torch.cuda.synchronize()
    fw += a.elapsed_time(b)
    bw += b.elapsed_time(c)
line = {"B": B, "forward_ms": round(fw / n, 2), "backward_ms": round(bw / n, 2), "step_ms": round((fw + bw) / n, 2)}
print(json.dumps(line), flush=True)
# the whole optimisation step of the reference (add_noise, forward, MSE, backward, clip, AdamW, re-install of the adapters)
from faceposegenerator_b200 import DDPMScheduler  # noqa: E402
from faceposegenerator_b200.lora_backward import LoRATrainer  # noqa: E402
tr = LoRATrainer(unet, lora, DDPMScheduler.from_pretrained("stabilityai/stable-diffusion-2-1-base", subfolder="scheduler"))
ti = t.long()
for _ in range(2):
    tr.step(x, tgt, ti, ctx)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    tr.step(x, tgt, ti, ctx)
torch.cuda.synchronize()
line["trainer_step_wall_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 2)
print(json.dumps(line), flush=True)
unet.set_lora(lora)
# comparator: torch autograd through the oracle module graph, bf16 weights / activations, fp32 adapters (the reference trains
# with fp16 autocast + fp32 adapters, train_ID-Booth.py:779-785), SDPA attention, eager
try:
    from oracle import sd21
    sd21.USE_SDPA = True
    sdb = {k: (v.to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v.to(dev, torch.bfloat16)) for k, v in sd.items()}
    leaf = {k: (d.to(dev).requires_grad_(True), u.to(dev).requires_grad_(True), s) for k, (d, u, s) in lora.items()}

    def ref_step():
        pred = sd21.unet_forward(sdb, x.to(torch.bfloat16), t, ctx.to(torch.bfloat16), {k: (d.to(torch.bfloat16), u.to(torch.bfloat16), s) for k, (d, u, s) in leaf.items()})
        loss = F.mse_loss(pred.float(), tgt)
        return torch.autograd.grad(loss, [p for d, u, _ in leaf.values() for p in (d, u)])
    for _ in range(2):
        ref_step()
    torch.cuda.synchronize()
    a.record()
    for _ in range(3):
        ref_step()
    b.record()
    torch.cuda.synchronize()
    line["torch_eager_bf16_autograd_step_ms"] = round(a.elapsed_time(b) / 3, 2)
except Exception as e:  # noqa: BLE001
    line["torch_eager_bf16_autograd_step_ms"] = f"unavailable: {type(e).__name__}: {e}"[:160]
print(json.dumps(line))
