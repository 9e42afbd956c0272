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
           "AutoPipelineForText2Image", "DPMSolverMultistepScheduler", "StableDiffusionPipelineOutput"]
