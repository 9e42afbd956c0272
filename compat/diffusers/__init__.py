"""Shim so `from diffusers import StableDiffusionPipeline, DDPMScheduler, ...` at
/root/reference/inference_ID-Booth.py:1,5,6,13 resolves to the B200-native implementation when
the real `diffusers` is not installed:  PYTHONPATH=<repo>/compat:<repo> python inference_ID-Booth.py
"""
from faceposegenerator_b200 import (AutoencoderKL, AutoPipelineForText2Image, DDPMScheduler,  # noqa: F401
                                    DPMSolverMultistepScheduler, StableDiffusionPipeline, UNet2DConditionModel)

__version__ = "0.32.2+idb_b200"
