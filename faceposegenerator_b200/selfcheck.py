"""smoke(): one small invocation of the hot path on cuda:0, checked against the oracle.
Imported only by __graft_entry__.smoke() (the oracle is test infrastructure)."""
from __future__ import annotations

import time

import torch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def smoke_check(verbose: bool = False, tol: float = 1e-2) -> dict:
    from oracle import sd21  # checker only
    from . import ops
    from .scheduler import DDPMScheduler
    from .unet import UNet2DConditionModel
    from .weights import random_lora, random_state_dict, unet_manifest

    dev = torch.device("cuda:0")
    t0 = time.time()
    sd = random_state_dict(unet_manifest(), 0)
    lora = random_lora(seed=0)
    unet = UNet2DConditionModel(sd, device=dev)
    unet.set_lora(lora)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 64, 64, generator=g)
    ctx = torch.randn(2, 77, 1024, generator=g)
    noise = torch.randn(1, 4, 64, 64, generator=g)
    x2 = torch.cat([x, x])
    t = 958
    eps2 = unet.forward(x2.to(dev), t, ctx.to(dev), return_dict=False)[0]
    sch = DDPMScheduler()
    sch.set_timesteps(30, device=dev)
    prev = ops.cfg_ddpm_step(eps2, x.to(dev), noise.to(dev), sch.coef_row(0, t, dev), guidance_scale=5.0, use_cfg=True)
    torch.cuda.synchronize()
    t1 = time.time()
    with torch.no_grad():
        ref2 = sd21.unet_forward(sd, x2, t, ctx, lora)
        ref_sch = sd21.DDPMSchedulerRef()
        ref_sch.set_timesteps(30)
        e = ref2[:1] + 5.0 * (ref2[1:] - ref2[:1])
        ref_prev, _ = ref_sch.step(e, t, x, noise)
    res = {"unet_rel_l2": rel_l2(eps2, ref2), "latent_rel_l2": rel_l2(prev, ref_prev),
           "gpu_s": t1 - t0, "oracle_s": time.time() - t1}
    if verbose:
        print("smoke:", res)
    if not (res["unet_rel_l2"] < tol and res["latent_rel_l2"] < tol):
        raise AssertionError(f"smoke parity failed: {res}")
    return res
