"""Drop-in boundary: the call sequence of the reference's `inference_ID-Booth.py:97-144` replayed line by
line through the `compat/` shims -- `from diffusers import StableDiffusionPipeline, DDPMScheduler`,
`from accelerate.utils import set_seed`, per-(identity, model) `from_pretrained(...).to(device)`, scheduler swap,
`load_lora_weights(<dir>)` from a `pytorch_lora_weights.safetensors` on disk, `torch.Generator(device).manual_seed(id)`
shared across the identity's prompts, `pipe(prompt=str, negative_prompt=str, output_type="np", ...)`,
`torch.Tensor(output.images)` -> `save_image`.  The reference script itself cannot travel to the GPU box
(`/root/reference` does not exist there), so the same statements are issued here with the script's own constants."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# constants of inference_ID-Booth.py:46-49,63,81-82
GUIDANCE_SCALE, NUM_INFERENCE_STEPS, WIDTH, HEIGHT = 5.0, 30, 512, 512
MODEL_ARCHITECTURE = "stabilityai/stable-diffusion-2-1-base"
CHECKPOINT = "checkpoint-31-6400"
NEGATIVE_PROMPT = ("cartoon, cgi, render, illustration, painting, drawing, black and white, bad body proportions, "
                   "landscape")
PROMPTS = ["face portrait photo of female sks person, forest background",
           "face side-portrait photo of female sks person, office background"]


@pytest.fixture(scope="module")
def shims():
    compat = os.path.join(ROOT, "compat")
    sys.path.insert(0, compat)
    try:
        import diffusers
        if not getattr(diffusers, "__version__", "").endswith("idb_b200"):
            pytest.skip("a real diffusers is installed; the shim is not on the import path")
        from accelerate.utils import set_seed
        from diffusers import DDPMScheduler, StableDiffusionPipeline
        yield StableDiffusionPipeline, DDPMScheduler, set_seed
    finally:
        sys.path.remove(compat)


def _make_lora_tree(root, models, which_id):
    from faceposegenerator_b200.weights import random_lora, save_lora_weights
    for k, name in enumerate(models):
        save_lora_weights(os.path.join(root, name, which_id, CHECKPOINT), random_lora(seed=10 + k, up_std=0.05))


def test_inference_script_call_sequence(tmp_path, cuda_dev, shims):
    StableDiffusionPipeline, DDPMScheduler, set_seed = shims
    from torchvision.utils import save_image
    device = "cuda:0"
    folder_of_models = str(tmp_path / "Trained_LoRA_Models")
    models_to_test = ["DreamBooth", "ID-Booth"]
    which_id, id_number = "1", 0
    _make_lora_tree(folder_of_models, models_to_test, which_id)
    set_seed(0)

    def run_identity(model_name, n_prompts, **extra):
        full_model_path = os.path.join(folder_of_models, model_name, which_id, CHECKPOINT)
        pipe = StableDiffusionPipeline.from_pretrained(MODEL_ARCHITECTURE, torch_dtype=torch.float16).to(device)
        pipe.scheduler = DDPMScheduler.from_pretrained(MODEL_ARCHITECTURE, subfolder="scheduler")
        pipe.load_lora_weights(full_model_path)
        pipe.set_progress_bar_config(disable=True)
        generator = torch.Generator(device=device).manual_seed(id_number)
        outs = []
        for i in range(n_prompts):
            output = pipe(prompt=PROMPTS[i], negative_prompt=NEGATIVE_PROMPT, output_type="np", generator=generator,
                          num_inference_steps=NUM_INFERENCE_STEPS, guidance_scale=GUIDANCE_SCALE, width=WIDTH,
                          height=HEIGHT, **extra)
            outs.append(output.images)
        return pipe, outs

    pipe, imgs = run_identity("DreamBooth", 2)
    for im in imgs:   # `.images` with output_type="np": float32 [n, 512, 512, 3] in [0, 1]
        assert isinstance(im, np.ndarray) and im.dtype == np.float32 and im.shape == (1, HEIGHT, WIDTH, 3)
        assert np.isfinite(im).all() and im.min() >= 0.0 and im.max() <= 1.0 and im.std() > 1e-3
    # the tail of the script's loop body: torch.Tensor(...), permute, save_image
    output = torch.permute(torch.Tensor(imgs[0]), (0, 3, 1, 2))
    path = str(tmp_path / f"0_0_{PROMPTS[0]}.png")
    save_image(output, fp=path)
    assert os.path.getsize(path) > 10_000

    # a pipeline rebuilt for the same (identity, model), as the script does for every pair, reproduces the images bit
    # for bit: cached base weights, adapters re-read from disk, generator re-seeded per identity
    _, imgs_again = run_identity("DreamBooth", 2)
    assert all(np.array_equal(a, b) for a, b in zip(imgs, imgs_again))
    # the identity's generator is shared by its prompts: the second prompt starts from the advanced state
    g2 = torch.Generator(device=device).manual_seed(id_number)
    second_alone = pipe(prompt=PROMPTS[1], negative_prompt=NEGATIVE_PROMPT, output_type="np", generator=g2,
                        num_inference_steps=NUM_INFERENCE_STEPS, guidance_scale=GUIDANCE_SCALE, width=WIDTH,
                        height=HEIGHT).images
    assert not np.array_equal(second_alone, imgs[1])
    # another model's adapters (hot-swapped onto the same cached base weights) change the image
    _, imgs_other = run_identity("ID-Booth", 1)
    assert np.abs(imgs_other[0] - imgs[0]).mean() > 1e-3

    # draw order / shape / dtype / device of diffusers: one fp16 draw of the initial latent in `prepare_latents`, then one
    # fp16 draw per `DDPMScheduler.step` with t > 0 (all 30 here: the last timestep is 1), all from the caller's generator
    g3 = torch.Generator(device=device).manual_seed(id_number)
    tape = torch.stack([torch.randn((1, 4, HEIGHT // 8, WIDTH // 8), generator=g3, device=device, dtype=torch.float16)
                        for _ in range(1 + NUM_INFERENCE_STEPS)]).float()
    pipe.load_lora_weights(os.path.join(folder_of_models, "DreamBooth", which_id, CHECKPOINT))   # the cached UNet carries ID-Booth's adapters now
    by_tape = pipe(prompt=PROMPTS[0], negative_prompt=NEGATIVE_PROMPT, output_type="np", noise_tape=tape,
                   num_inference_steps=NUM_INFERENCE_STEPS, guidance_scale=GUIDANCE_SCALE, width=WIDTH, height=HEIGHT).images
    assert np.array_equal(by_tape, imgs[0])
