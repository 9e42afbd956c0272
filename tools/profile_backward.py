"""One LoRA-only training step (forward with tape + backward, B = 4) between cudaProfilerStart / Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --csv`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200.lora_backward import UNetLoRAGrad  # noqa: E402
from faceposegenerator_b200.unet import UNet2DConditionModel  # noqa: E402
from faceposegenerator_b200.weights import random_lora, random_state_dict, unet_manifest  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
lora = random_lora(seed=1)
unet = UNet2DConditionModel(random_state_dict(unet_manifest(), 0), device=dev)
unet.set_lora(lora)
eng = UNetLoRAGrad(unet, lora)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 4, 64, 64, generator=g).to(dev)
ctx = torch.randn(B, 77, 1024, generator=g).to(dev)
t = torch.randint(0, 1000, (B,), generator=g).float().to(dev)
tgt = torch.randn(B, 4, 64, 64, generator=g).to(dev)
for i in range(2):
    if i == 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    eps = eng.forward(x, t, ctx)
    torch.cuda.synchronize()
    if i == 1:
        print("forward done", file=sys.stderr)
    eng.backward(2.0 * (eps - tgt) / eps.numel())
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
