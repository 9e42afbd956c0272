"""One eager UNet forward (B=8 rows, 64x64 latent, LoRA) for ncu: 2 warm-up forwards, then 1 profiled.
Kernel launches before the profiled forward: 16 (encode_context) + 2 * L, printed on stderr.
`python tools/profile_step.py 8 pair`: the forward the pipeline runs under classifier-free guidance (B / 2 latents, 2 contexts
each: `forward(cfg_pair=True)`, the layers in front of the first cross-attention evaluated once)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import _lib  # noqa: E402
from faceposegenerator_b200.unet import UNet2DConditionModel  # noqa: E402
from faceposegenerator_b200.weights import random_lora  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
PAIR = len(sys.argv) > 2 and sys.argv[2] == "pair"
dev = torch.device("cuda:0")
unet = UNet2DConditionModel.from_random(0, device=dev)
unet.set_lora(random_lora(seed=0))
x = torch.randn(B // 2 if PAIR else B, 4, 64, 64, device=dev)
ctx = torch.randn(B, 77, 1024, device=dev)
t = torch.full((B,), 500.0, device=dev)
context = unet.encode_context(ctx)
temb = unet.time_embedding(t)   # hoisted by the pipeline (one batched call per image batch)
n0 = _lib.launch_count
for i in range(3):
    unet.forward(x, t, context=context, temb=temb, cfg_pair=PAIR)
    torch.cuda.synchronize()
    if i == 0:
        print(f"launches per forward: {_lib.launch_count - n0}; before profiled forward: {n0 + 2 * (_lib.launch_count - n0)}",
              file=sys.stderr)
# shape trace of one more forward (not profiled by `-c`), joined with the ncu launch list by tools/join_trace.py
import json  # noqa: E402
_lib.trace = []
unet.forward(x, t, context=context, temb=temb, cfg_pair=PAIR)
torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
with open(os.environ.get("IDB_TRACE_OUT", "gpurun_out/step_trace.json"), "w") as f:
    json.dump(_lib.trace, f)
