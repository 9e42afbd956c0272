"""oracle/ -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU (fp32/fp64 torch) restatement of the algorithm on the ID-Booth hot path
(`/root/reference/inference_ID-Booth.py:103-138`): diffusers==0.32.2
`UNet2DConditionModel.forward`, `AutoencoderKL.decode`, `DDPMScheduler`,
peft LoRA `Linear.forward` and the `StableDiffusionPipeline.__call__` loop.

PARITY UNPINNED.  The arithmetic of the path lives in third-party packages
(`diffusers==0.32.2`, `peft`, `requirements.txt:4-6`) that are neither vendored
in the reference tree nor installed in this image, and the reference holds no
tests or golden vectors.  The restatement follows the
published diffusers 0.32.2 semantics and is pinned only by
self-made anchors: exact parameter counts (865,910,724 / 49,490,199), the
closed-form scheduler known answers (App. C), metamorphic identities (merged
LoRA == unmerged LoRA in fp64, ...).  That statement covers oracle/sd21.py (UNet, VAE,
scheduler, pipeline loop).  The pieces that CAN be checked against reference code run in
this container are PINNED by golden vectors under tests/golden/ (each with its generating
script): oracle/iresnet.py against `/root/reference/ArcFace_files/backbones/iresnet.py`,
oracle/arcface_glue.py against the glue functions of `/root/reference/train_ID-Booth.py:433-455`,
oracle/clip_text.py against transformers' own `CLIPTextModel`, and the caller-side host logic
(faceposegenerator_b200/sweep.py, extract_embeds.py) against logs of `/root/reference/inference_ID-Booth.py` and
`/root/reference/extract_ArcFace_embeds.py` themselves, executed under recording stubs.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package, and there only as the checker / the
reported CPU baseline, never as the thing shipped.
"""
