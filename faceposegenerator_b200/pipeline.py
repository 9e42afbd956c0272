"""`StableDiffusionPipeline` with the diffusers 0.32.2 call surface that
`/root/reference/inference_ID-Booth.py:103-108,138` and `README.md:66-85` use:

    pipe = StableDiffusionPipeline.from_pretrained(id, torch_dtype=torch.float16).to("cuda:0")
    pipe.scheduler = DDPMScheduler.from_pretrained(id, subfolder="scheduler")
    pipe.load_lora_weights(dir); pipe.set_progress_bar_config(disable=True)
    pipe(prompt=..., negative_prompt=..., output_type="np", generator=g, num_inference_steps=30,
         guidance_scale=5.0, width=512, height=512).images

The denoising loop is B200-native: per step one CUDA-graph replay of
[duplicate latents -> UNet (hand-written sm_100a kernels) -> fused CFG + DDPMScheduler.step],
with the step's coefficients and timestep read from device tables (no host sync in the loop).
Cross-attention K/V projections of the text context are computed once per call.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import List, Optional, Union

import torch

from . import ops
from .scheduler import DDPMScheduler, randn_tensor
from .text import CLIPTextEncoder, load_tokenizer, text_manifest
from .unet import UNet2DConditionModel
from .vae import AutoencoderKL
from .weights import (LORA_FILE, UNET_CONFIG, VAE_CONFIG, load_lora_state, random_state_dict, unet_manifest,
                      vae_decoder_manifest)

f32 = torch.float32
_COMPONENT_CACHE = {}   # (model id, device) -> components with packed base weights (LoRA hot-swaps on top)


class StableDiffusionPipelineOutput:
    def __init__(self, images, nsfw_content_detected=None):
        self.images = images
        self.nsfw_content_detected = nsfw_content_detected

    def __getitem__(self, i):
        return (self.images, self.nsfw_content_detected)[i]


def _load_component_state(root: str, sub: str):
    from safetensors.torch import load_file
    for fn in ("diffusion_pytorch_model.safetensors", "model.safetensors",
               "diffusion_pytorch_model.fp16.safetensors", "model.fp16.safetensors"):
        p = os.path.join(root, sub, fn)
        if os.path.isfile(p):
            return load_file(p)
    return None


class StableDiffusionPipeline:
    def __init__(self, model_id: str, torch_dtype=None, seed: int = 0):
        self.model_id = model_id
        self.torch_dtype = torch_dtype or torch.float32
        self.weight_seed = seed
        self.device = torch.device("cpu")
        self.scheduler = DDPMScheduler.from_pretrained(model_id, subfolder="scheduler")
        self.unet: Optional[UNet2DConditionModel] = None
        self.vae: Optional[AutoencoderKL] = None
        self.text_encoder: Optional[CLIPTextEncoder] = None
        self.safety_checker = None
        self._lora = None
        self._progress = {}
        self._graphs = {}
        self.use_cuda_graph = os.environ.get("IDB_CUDA_GRAPH", "1") != "0"
        self.step_events = None    # set to a list to collect (start, end) CUDA events around every denoise step

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, torch_dtype=None, **kwargs):
        """Loads `<dir>/{unet,vae,text_encoder}/*.safetensors` when given a local snapshot;
        otherwise (offline, HF_HUB_OFFLINE=1) falls back to the built-in SD2.1-base configs with
        deterministic random-init weights (weights.random_state_dict)."""
        return cls(str(pretrained_model_name_or_path), torch_dtype=torch_dtype, seed=int(kwargs.get("weight_seed", 0)))

    def to(self, device=None, dtype=None):
        if device is None:
            return self
        device = torch.device(device)
        if device.type != "cuda":
            if self.unet is not None:
                raise RuntimeError("this pipeline has no CPU path")
            self.device = device
            return self
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA device required: the pipeline runs hand-written sm_100a kernels only "
                               "(no CPU / eager fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        key = (self.model_id, str(device), self.weight_seed)
        if key not in _COMPONENT_CACHE or os.environ.get("IDB_NO_WEIGHT_CACHE") == "1":
            with torch.cuda.device(device):
                local = self.model_id if os.path.isdir(self.model_id) else None
                usd = _load_component_state(local, "unet") if local else None
                vsd = _load_component_state(local, "vae") if local else None
                tsd = _load_component_state(local, "text_encoder") if local else None
                unet = UNet2DConditionModel(usd or random_state_dict(unet_manifest(UNET_CONFIG), self.weight_seed),
                                            UNET_CONFIG, device)
                vae = AutoencoderKL(vsd or random_state_dict(vae_decoder_manifest(VAE_CONFIG), self.weight_seed),
                                    VAE_CONFIG, device)
                text = CLIPTextEncoder(tsd or random_state_dict(text_manifest(), self.weight_seed), device,
                                       tokenizer=load_tokenizer(local))
            _COMPONENT_CACHE[key] = (unet, vae, text)
        self.unet, self.vae, self.text_encoder = _COMPONENT_CACHE[key]
        self.unet.set_lora(self._lora)   # a fresh pipeline starts without (or with its own) adapters
        self._graphs = {}
        return self

    # ------------------------------------------------------------------ LoRA lifecycle
    def load_lora_weights(self, pretrained_model_name_or_path_or_dict, weight_name: str = LORA_FILE, **kwargs):
        if isinstance(pretrained_model_name_or_path_or_dict, dict):
            lora = pretrained_model_name_or_path_or_dict
        else:
            lora = load_lora_state(str(pretrained_model_name_or_path_or_dict), weight_name)
        self._lora = lora
        if self.unet is not None:
            self.unet.set_lora(lora)
        self._graphs = {}

    def unload_lora_weights(self):
        self._lora = None
        if self.unet is not None:
            self.unet.set_lora(None)
        self._graphs = {}

    def set_progress_bar_config(self, **kwargs):
        self._progress = dict(kwargs)

    def progress_bar(self, iterable):
        if self._progress.get("disable", False):
            return iterable
        try:
            from tqdm.auto import tqdm
            return tqdm(iterable, **{k: v for k, v in self._progress.items() if k != "disable"})
        except Exception:
            return iterable

    # ------------------------------------------------------------------ prompt encoding
    def encode_prompt(self, prompt, negative_prompt, n: int, do_cfg: bool, prompt_embeds=None,
                      negative_prompt_embeds=None):
        if prompt_embeds is None:
            prompts = [prompt] if isinstance(prompt, str) else list(prompt)
            prompt_embeds = self.text_encoder.encode(prompts)
        prompt_embeds = prompt_embeds.to(self.device)
        if do_cfg and negative_prompt_embeds is None:
            if negative_prompt is None:
                neg = [""] * prompt_embeds.shape[0]
            elif isinstance(negative_prompt, str):
                neg = [negative_prompt] * prompt_embeds.shape[0]
            else:
                neg = list(negative_prompt)
            negative_prompt_embeds = self.text_encoder.encode(neg)
        if negative_prompt_embeds is not None:
            negative_prompt_embeds = negative_prompt_embeds.to(self.device)
        return prompt_embeds, negative_prompt_embeds

    # ------------------------------------------------------------------ one denoising step (graph-captured)
    def _step_eager(self, st):
        n = st.n
        if st.do_cfg:
            st.x2[:n].copy_(st.latents)
            st.x2[n:].copy_(st.latents)
        else:
            st.x2.copy_(st.latents)
        eps2 = self.unet.forward(st.x2, st.t_dev, context=st.context, temb=st.temb, return_dict=False)[0]
        ops.cfg_ddpm_step(eps2, st.latents, st.noise, st.coef, guidance_scale=st.guidance_scale,
                          use_cfg=st.do_cfg, v_prediction=st.vpred, x_prev=st.lat_next)
        st.latents.copy_(st.lat_next)

    def _make_step_state(self, n, h, w, do_cfg, guidance_scale, context):
        dev = self.device
        rows = 2 * n if do_cfg else n
        return SimpleNamespace(
            n=n, do_cfg=do_cfg, guidance_scale=float(guidance_scale), context=context,
            vpred=self.scheduler.config.prediction_type == "v_prediction",
            latents=torch.zeros((n, 4, h, w), dtype=f32, device=dev),
            lat_next=torch.zeros((n, 4, h, w), dtype=f32, device=dev),
            noise=torch.zeros((n, 4, h, w), dtype=f32, device=dev),
            x2=torch.zeros((rows, 4, h, w), dtype=f32, device=dev),
            t_dev=torch.zeros((rows,), dtype=f32, device=dev),
            temb=torch.zeros((rows, self.unet.t_w_all.shape[0]), dtype=f32, device=dev),
            coef=torch.zeros((5,), dtype=f32, device=dev), graph=None, launches_per_step=0)

    # ------------------------------------------------------------------ __call__
    @torch.no_grad()
    def __call__(self, prompt: Union[str, List[str], None] = None, height: Optional[int] = None,
                 width: Optional[int] = None, num_inference_steps: int = 50, guidance_scale: float = 7.5,
                 negative_prompt=None, num_images_per_prompt: int = 1, generator=None, latents=None,
                 prompt_embeds=None, negative_prompt_embeds=None, output_type: str = "pil", return_dict: bool = True,
                 **kwargs):
        if self.unet is None:
            raise RuntimeError("call .to('cuda:N') first: the pipeline runs on a B200 only (no CPU path)")
        dev = self.device
        with torch.cuda.device(dev):
            height = height or self.unet.config.sample_size * 8
            width = width or self.unet.config.sample_size * 8
            if height % 64 or width % 64:
                raise ValueError("height and width must be multiples of 64")
            do_cfg = guidance_scale > 1.0
            pe, ne = self.encode_prompt(prompt, negative_prompt, 0, do_cfg, prompt_embeds, negative_prompt_embeds)
            if num_images_per_prompt > 1:
                pe = pe.repeat_interleave(num_images_per_prompt, dim=0)
                ne = ne.repeat_interleave(num_images_per_prompt, dim=0) if ne is not None else None
            n = pe.shape[0]
            ctx = torch.cat([ne, pe], dim=0) if do_cfg else pe       # uncond first (A.1 step 2)
            self.scheduler.set_timesteps(num_inference_steps, device=dev)
            timesteps = self.scheduler._timesteps_list
            h, w = height // 8, width // 8
            draw_dtype = self.torch_dtype
            if latents is None:
                latents = randn_tensor((n, 4, h, w), generator=generator, device=dev, dtype=draw_dtype)
            # parity hooks (not part of the diffusers surface): a pre-drawn noise tape shared with the
            # oracle, teacher forcing of each step's input latent, and per-step latent collection
            noise_tape = kwargs.get("noise_tape")
            teacher = kwargs.get("teacher_latents")
            collected = [] if kwargs.get("collect_latents") else None
            if noise_tape is not None:
                latents = noise_tape[0]
            latents = latents.to(device=dev, dtype=f32) * self.scheduler.init_noise_sigma

            context = self.unet.encode_context(ctx)
            key = (n, h, w, do_cfg, float(guidance_scale), self.unet._lora_version)
            st = self._graphs.get(key) if self.use_cuda_graph else None
            if st is None:
                st = self._make_step_state(n, h, w, do_cfg, guidance_scale, context)
                if self.use_cuda_graph:
                    self._graphs = {key: st}     # keep one graph (its private pool holds all activations)
            # refresh the step-invariant context projections in place (graph reads these buffers)
            if st.context is not context:
                for dst, src in zip(st.context.kv, context.kv):
                    dst.copy_(src)
            st.latents.copy_(latents)
            t_table = torch.tensor(timesteps, dtype=f32, device=dev)
            # the time embedding + all 22 time_emb_proj layers depend on the timestep only: one batched call for
            # every step of this image batch instead of one per step
            temb_table = self.unet.time_embedding(t_table)

            for i in self.progress_bar(range(len(timesteps))):
                t = timesteps[i]
                st.t_dev.copy_(t_table[i].expand_as(st.t_dev))
                st.temb.copy_(temb_table[i].expand_as(st.temb))
                st.coef.copy_(self.scheduler.coef_row(i, t, dev))
                if teacher is not None:
                    st.latents.copy_(teacher[i])
                if noise_tape is not None:
                    st.noise.copy_(noise_tape[1 + i])
                elif t > 0:   # same draw order / shape / dtype / device as diffusers DDPMScheduler.step
                    st.noise.copy_(randn_tensor((n, 4, h, w), generator=generator, device=dev, dtype=draw_dtype))
                else:
                    st.noise.zero_()
                if self.use_cuda_graph:
                    if st.graph is None:
                        saved = st.latents.clone()
                        self._step_eager(st)               # warm-up (lazy kernel attribute setup)
                        st.latents.copy_(saved)
                        torch.cuda.synchronize(dev)
                        g = torch.cuda.CUDAGraph()
                        from . import _lib
                        n0 = _lib.launch_count
                        with torch.cuda.graph(g):
                            self._step_eager(st)
                        st.launches_per_step = _lib.launch_count - n0
                        st.graph = g
                        st.latents.copy_(saved)
                    if self.step_events is not None:
                        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                        ev[0].record()
                    st.graph.replay()
                    if self.step_events is not None:
                        ev[1].record()
                        self.step_events.append(ev)
                else:
                    self._step_eager(st)
                if collected is not None:
                    collected.append(st.latents.clone())
            latents = st.latents.clone()

            if output_type == "latent":
                images = latents
            else:
                img = self.vae.decode(latents / self.vae.config.scaling_factor, output_image=True)[0]   # NHWC [0,1]
                if output_type == "np":
                    images = img.cpu().numpy()
                elif output_type == "pt":
                    images = img.permute(0, 3, 1, 2)
                elif output_type == "pil":
                    from PIL import Image
                    arr = (img * 255).round().to(torch.uint8).cpu().numpy()
                    images = [Image.fromarray(a) for a in arr]
                else:
                    raise ValueError(f"unknown output_type {output_type!r}")
        if not return_dict:
            return (images, None)
        out = StableDiffusionPipelineOutput(images)
        if collected is not None:
            out.step_latents = torch.stack(collected)
        return out


class AutoPipelineForText2Image:
    """Imported (unused) by inference_ID-Booth.py:13."""
    from_pretrained = StableDiffusionPipeline.from_pretrained


class DPMSolverMultistepScheduler:
    """Imported (unused) by inference_ID-Booth.py:5; not on the hot path."""

    def __init__(self, *a, **k):
        raise NotImplementedError("DPMSolverMultistepScheduler is outside the accelerated path; use DDPMScheduler")

    @classmethod
    def from_pretrained(cls, *a, **k):
        return cls()
