"""faceposegenerator_b200 -- B200-native (sm_100a) implementation of the Stable Diffusion 2.1
denoising hot path that ID-Booth (`rangasaishreyas/FacePoseGenerator`) drives through
diffusers: `StableDiffusionPipeline(...)` + `load_lora_weights()` + `DDPMScheduler`.

Python host code -> thin C ABI (`include/idb.h`, `libidb_b200.so`) -> hand-written CUDA kernels
(tcgen05/TMEM implicit-GEMM convolutions and linears fed by TMA, flash-style attention, fused
norms, fused CFG + DDPM step).  There is no CPU / eager fallback: without the built library or a
B200 every compute call raises.
"""
from .pipeline import (AutoPipelineForText2Image, DPMSolverMultistepScheduler,  # noqa: F401
                       StableDiffusionPipeline, StableDiffusionPipelineOutput)
from .scheduler import DDPMScheduler  # noqa: F401
from .unet import UNet2DConditionModel  # noqa: F401
from .vae import AutoencoderKL  # noqa: F401

__all__ = ["StableDiffusionPipeline", "DDPMScheduler", "UNet2DConditionModel", "AutoencoderKL",
           "AutoPipelineForText2Image", "DPMSolverMultistepScheduler", "StableDiffusionPipelineOutput", "patch"]


def patch(pipe, lora=None, device=None):
    """Swap the B200 components into an existing diffusers-style pipeline object IN PLACE: `pipe.unet`,
    `pipe.vae` and `pipe.scheduler` are rebuilt from the pipeline's own `state_dict()`s / scheduler config (the key names
    and call surfaces are the diffusers ones), LoRA adapters (`lora`: a directory / file written by the reference trainer,
    `train_ID-Booth.py:696-720`, or a {path: (down, up, scale)} dict) are installed fused and unmerged instead of through
    peft injection.  Returns `pipe`."""
    import torch
    from .weights import UNET_CONFIG, VAE_CONFIG, load_lora_state
    dev = torch.device(device or getattr(pipe, "device", None) or "cuda:0")
    if dev.type != "cuda":
        raise RuntimeError("patch(): the pipeline must live on a CUDA device (sm_100a); there is no CPU path")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(dev):
        unet = UNet2DConditionModel({k: v.detach().float().cpu() for k, v in pipe.unet.state_dict().items()}, UNET_CONFIG, dev)
        vae = AutoencoderKL({k: v.detach().float().cpu() for k, v in pipe.vae.state_dict().items()}, VAE_CONFIG, dev)
        cfg = pipe.scheduler.config
        sched = DDPMScheduler.from_config(dict(cfg) if isinstance(cfg, dict) else cfg)
        if lora is not None:
            unet.set_lora(lora if isinstance(lora, dict) else load_lora_state(str(lora)))
    pipe.unet, pipe.vae, pipe.scheduler = unet, vae, sched
    return pipe
