"""Generates tests/golden/iresnet100_golden.pt by running the REFERENCE module
(/root/reference/ArcFace_files/backbones/iresnet.py) in this container.  Only outputs are stored;
weights are regenerated deterministically from key names (oracle.iresnet.keyed_state_dict).
    python tests/golden/make_iresnet_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/ArcFace_files")
from backbones.iresnet import iresnet100  # noqa: E402  (the reference's own code)
from oracle.iresnet import keyed_state_dict  # noqa: E402

torch.manual_seed(0)
model = iresnet100(fp16=False).eval()
sd = keyed_state_dict(model.state_dict(), seed=0)
model.load_state_dict(sd)
x = torch.randn(2, 3, 112, 112, generator=torch.Generator().manual_seed(0))
feats = {}
model.layer1.register_forward_hook(lambda m, i, o: feats.__setitem__("layer1", o))
model.layer3.register_forward_hook(lambda m, i, o: feats.__setitem__("layer3", o))
with torch.no_grad():
    y = model(x)
out = {"embedding": y, "n_params": sum(p.numel() for p in model.parameters()), "n_state": len(sd),
       "layer1_mean_abs": feats["layer1"].abs().mean(), "layer3_mean_abs": feats["layer3"].abs().mean(),
       "layer3_slice": feats["layer3"][:, :8, :4, :4].clone()}
torch.save(out, os.path.join(os.path.dirname(os.path.abspath(__file__)), "iresnet100_golden.pt"))
print({k: (v.shape if hasattr(v, "shape") and v.dim() else v) for k, v in out.items()}, y.abs().mean())
