#!/usr/bin/env python
"""bench.py -- 512x512, 30-step, CFG-5.0 DDPM image generation throughput (images/s) of the
ID-Booth SD2.1 hot path on N B200s (one process per GPU, data-parallel by image).

    python bench.py --gpus 1 --steps K --warmup W            # this framework, BASELINE.json configs[1]
    python bench.py --impl reference --steps K --warmup W    # the restated reference path on host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --config 4                               # configs[3]: 768x768, UNet batch 16 (8 images x CFG pair)
    python bench.py --config 3 [--gpus N under torchrun]     # configs[2]: the 1024-image sweep, strong scaling, PNGs written

A "step" is one pass of the hot path over one batch of synthetic input: one pipeline call that
generates 4 images (UNet batch 8 = 4 prompts x CFG pair, BASELINE.json configs[1]): 30 x
[UNet -> fused CFG + DDPMScheduler.step] + VAE decode + post-process.
  value : images/s with prompt embeddings and the noise tape already resident in HBM
  e2e   : same metric through the public pipe(...) call with HOST (pinned) prompt embeddings /
          initial latents copied H2D and the decoded images read back D2H inside the timed region
          (`e2e.text_included`: the same with prompt STRINGS through the CLIP tower, as `inference_ID-Booth.py:138` calls it;
           `e2e.script_mode`: one prompt per call = UNet batch 2, the reference script's own call pattern)
  roofline : the UNet step (the graph launch that dominates the loop): 6.4432 TFLOP per B=8 forward
           / its mean CUDA-event duration inside the timed region, vs the measured
          sustained bf16 peak of MEASURED_PEAKS.json
  cpu_baseline : the fp32 oracle (oracle/sd21.py = the torch CPU ops diffusers would dispatch to)
          timed on this box's host cores on a bounded sample (N=1, rank 0 only)
  library_baseline : the same oracle module graph on THIS GPU in bf16 through torch eager (cuDNN convs, cuBLAS
          linears, SDPA attention) on the same config -- the stack the reference would effectively run here
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

UNET_TFLOP_PER_ROW = 0.8054      # SURVEY.md 8(d) / BASELINE.md section 2 (64x64 latent, ctx 77, LoRA r=4)
UNET_TFLOP_PER_ROW_96 = 2.1516   # same at 96x96 latents (768x768 images)
VAE_TFLOP_PER_IMAGE = 2.5145
IMAGES_PER_CALL = 4              # 4 prompts x CFG pair -> UNet batch 8
NUM_STEPS = 30
GUIDANCE = 5.0
MODEL = "stabilityai/stable-diffusion-2-1-base"


def _step_traffic():
    """DRAM bytes of one UNet step from the newest committed ncu capture (profiles/r*_step_traffic.json), or None."""
    import glob
    for fn in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_step_traffic*.json")), reverse=True):
        try:
            with open(fn) as f:
                return float(json.load(f)["traffic_bytes"])
        except Exception:
            continue
    return None


def load_peaks():
    fn = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(fn):
        with open(fn) as f:
            p = json.load(f)
        return dict(bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU baseline / reference arm
def cpu_oracle_sample(n_steps: int, with_vae: bool, threads: int):
    """Times the restated reference path (fp32 torch CPU ops) on a bounded sample of configs[0]
    semantics: `n_steps` CFG denoise steps for ONE image (UNet batch 2) [+ one VAE decode]."""
    from oracle import sd21
    from faceposegenerator_b200.weights import (random_lora, random_state_dict, unet_manifest, vae_decoder_manifest)
    torch.set_num_threads(threads)
    sd = random_state_dict(unet_manifest(), 0)
    lora = random_lora(seed=0)
    g = torch.Generator().manual_seed(0)
    tape = torch.randn(1 + n_steps, 1, 4, 64, 64, generator=g)
    pe, ne = torch.randn(1, 77, 1024, generator=g), torch.randn(1, 77, 1024, generator=g)
    step_s = []
    with torch.no_grad():
        lat = tape[0]
        sch = sd21.DDPMSchedulerRef()
        sch.set_timesteps(NUM_STEPS)
        ctx = torch.cat([ne, pe])
        for i in range(n_steps):
            t0 = time.perf_counter()
            t = int(sch.timesteps[i])
            eps = sd21.unet_forward(sd, torch.cat([lat, lat]), t, ctx, lora)
            e = eps[:1] + GUIDANCE * (eps[1:] - eps[:1])
            lat, _ = sch.step(e, t, lat, tape[1 + i])
            step_s.append(time.perf_counter() - t0)
        vae_s = None
        if with_vae:
            vsd = random_state_dict(vae_decoder_manifest(), 0)
            t0 = time.perf_counter()
            sd21.postprocess_np(sd21.vae_decode(vsd, lat / 0.18215))
            vae_s = time.perf_counter() - t0
    return step_s, vae_s


def library_baseline_sample(dev, n_images: int, calls: int = 2):
    """The oracle's module graph (oracle/sd21.py) on the GPU in bf16 through torch eager: cuDNN convolutions
    (channels_last weights), cuBLAS linears, `F.scaled_dot_product_attention` (what diffusers' AttnProcessor2_0 calls)
    -- the library stack the reference would run on this box.  Same config as the product arm: `n_images` images x CFG
    pair per call, 30 steps, + VAE decode.  A reported comparator: nothing of it is on the product path."""
    from oracle import sd21
    from faceposegenerator_b200.weights import random_lora, random_state_dict, unet_manifest, vae_decoder_manifest
    bf = torch.bfloat16

    def cast(sd):
        out = {}
        for k, v in sd.items():
            v = v.to(device=dev, dtype=bf)
            out[k] = v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v
        return out
    sd, vsd = cast(random_state_dict(unet_manifest(), 0)), cast(random_state_dict(vae_decoder_manifest(), 0))
    lora = {k: (d.to(dev, bf), u.to(dev, bf), s) for k, (d, u, s) in random_lora(seed=0).items()}
    g = torch.Generator().manual_seed(0)
    pe = torch.randn(n_images, 77, 1024, generator=g).to(dev, bf)
    ne = torch.randn(n_images, 77, 1024, generator=g).to(dev, bf)
    tape = torch.randn(1 + NUM_STEPS, n_images, 4, 64, 64, generator=g).to(dev, bf)
    sch = sd21.DDPMSchedulerRef()
    sch.set_timesteps(NUM_STEPS)
    ctx = torch.cat([ne, pe])
    old = sd21.USE_SDPA
    sd21.USE_SDPA = True
    unet_ms = []
    try:
        def one_call(timed):
            lat = tape[0]
            for i, t in enumerate(sch.timesteps.tolist()):
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                x2 = torch.cat([lat, lat]).contiguous(memory_format=torch.channels_last)
                eps = sd21.unet_forward(sd, x2, t, ctx, lora)
                e1.record()
                if timed:
                    unet_ms.append((e0, e1))
                eu, ec = eps.chunk(2)
                lat, _ = sch.step((eu + GUIDANCE * (ec - eu)).to(bf), t, lat, tape[1 + i])
                lat = lat.to(bf)
            return sd21.vae_decode(vsd, (lat / 0.18215).contiguous(memory_format=torch.channels_last))
        with torch.no_grad():
            one_call(False)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(calls):
                img = one_call(True)
                (img * 0.5 + 0.5).clamp(0, 1)
            b.record()
            torch.cuda.synchronize(dev)
    finally:
        sd21.USE_SDPA = old
    ms = a.elapsed_time(b) / calls
    step_ms = sum(x.elapsed_time(y) for x, y in unet_ms) / max(len(unet_ms), 1)
    return {"value": n_images / (ms / 1e3), "unit": "images/s", "unet_step_ms": step_ms,
            "kind": "torch eager bf16 (cuDNN channels_last convs / cuBLAS / SDPA), oracle module graph, no CUDA graph",
            "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
            "sample": f"{calls} calls of {n_images} images x CFG pair, {NUM_STEPS} steps + VAE decode, inputs resident"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU path.  diffusers/peft are not installable here (no
    network, not vendored), so this times the oracle port with all host threads; each timed step
    is ONE CFG denoise step of one image, images/s = 1 / (30 * mean_step + vae_decode)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.warmup + args.steps
    step_s, vae_s = cpu_oracle_sample(n, True, threads)
    timed = step_s[args.warmup:]
    mean_step = sum(timed) / len(timed)
    value = 1.0 / (NUM_STEPS * mean_step + vae_s)
    sample = (f"{len(timed)} CFG denoise steps (UNet B=2, fp32) of one 512x512 image + 1 VAE decode, "
              f"extrapolated to {NUM_STEPS} steps; CPU throughput per image does not depend on the batch (compute bound), "
              "so one image stands for the 4-image call of the GPU arm")
    line = {"metric": "images_per_sec_512x512_30step_cfg5", "value": value, "unit": "images/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: SD2.1-base UNet random-init + rank-4 LoRA, 512x512, 30 DDPM steps, CFG 5.0",
                       "note": "bounded sample, one image"},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------ this framework
def _setup(args):
    import torch.distributed as dist
    from faceposegenerator_b200 import DDPMScheduler, StableDiffusionPipeline, _lib
    from faceposegenerator_b200.weights import random_lora
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().idb_device_check(), "idb_device_check")
    pipe = StableDiffusionPipeline.from_pretrained(MODEL, torch_dtype=torch.bfloat16, allow_random_weights=True).to(dev)
    pipe.scheduler = DDPMScheduler.from_pretrained(MODEL, subfolder="scheduler")
    pipe.load_lora_weights(random_lora(seed=0))          # synthetic "trained" rank-4 adapters, fused unmerged
    pipe.set_progress_bar_config(disable=True)
    return dist, pipe, world, rank, local, dev


def run_native(args):
    from faceposegenerator_b200 import _lib
    from faceposegenerator_b200.parallel import ImageGather, shard_units, unit_seed
    dist, pipe, world, rank, local, dev = _setup(args)

    big = args.config == 4
    n = 8 if big else IMAGES_PER_CALL          # images per call per GPU
    side, lat = (768, 96) if big else (512, 64)
    tflop_row = UNET_TFLOP_PER_ROW_96 if big else UNET_TFLOP_PER_ROW
    # one step = world * n image units; rank r generates units r, r + G, ... (weak scaling: n units per GPU per step)
    my_units = shard_units(world * n, rank, world)
    assert len(my_units) == n
    g = torch.Generator().manual_seed(unit_seed(my_units[0], 1000))
    # synthetic context (the resident / e2e legs feed embeddings; `text_included` feeds strings): N(0,1) embeddings
    pe_host = torch.randn(n, 77, 1024, generator=g).pin_memory()
    ne_host = torch.randn(n, 77, 1024, generator=g).pin_memory()
    lat_host = torch.randn(n, 4, lat, lat, generator=g).pin_memory()
    pe_dev, ne_dev = pe_host.to(dev), ne_host.to(dev)
    tape_dev = torch.randn(1 + NUM_STEPS, n, 4, lat, lat, generator=g).to(dev)
    gen = torch.Generator(device=dev).manual_seed(rank)
    gather = ImageGather(world * n, rank, world, (side, side, 3), dev) if world > 1 else None

    def call_resident():
        """inputs resident in HBM; result (uint8 images) stays on device; final NCCL gather when N > 1 (asynchronous:
        the all_gather of call i overlaps call i + 1 and is waited for one call later)"""
        out = pipe(prompt_embeds=pe_dev, negative_prompt_embeds=ne_dev, num_inference_steps=NUM_STEPS,
                   guidance_scale=GUIDANCE, height=side, width=side, output_type="pt", noise_tape=tape_dev)
        img = (out.images.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous()
        if gather is not None:
            gather.submit(img)
        return img

    def call_e2e():
        """public API with HOST buffers: H2D of embeddings + initial latents, D2H of the images"""
        out = pipe(prompt_embeds=pe_host.to(dev, non_blocking=True), negative_prompt_embeds=ne_host.to(dev, non_blocking=True),
                   latents=lat_host.to(dev, non_blocking=True), generator=gen, num_inference_steps=NUM_STEPS,
                   guidance_scale=GUIDANCE, height=side, width=side, output_type="np")
        return out.images

    text_calls = [0]

    def call_text(n_prompts=n):
        """the reference's own call: prompt STRINGS (new ones every call, so the prompt cache cannot help; the negative
        prompt is the sweep's constant one and is cached, `inference_ID-Booth.py:81`) through the CLIP tower"""
        text_calls[0] += 1
        prompts = [f"face portrait photo of sks person, variation {text_calls[0]}-{i}, rank {rank}" for i in range(n_prompts)]
        out = pipe(prompt=prompts if n_prompts > 1 else prompts[0],
                   negative_prompt="cartoon, cgi, render, illustration, painting, drawing, black and white, bad body proportions, landscape",
                   generator=gen, num_inference_steps=NUM_STEPS, guidance_scale=GUIDANCE, height=side, width=side,
                   output_type="np")
        return out.images

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, reps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        if gather is not None:
            gather.wait()
        b.record()
        barrier()
        return a.elapsed_time(b)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        call_resident()
    barrier()

    # ---- timed region 1: resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # CUDA events on the launching stream around the denoising launches: one pair per whole-loop graph (30 steps in one
    # launch, the default) or one pair per step graph (IDB_LOOP_GRAPH=0)
    step_events, loop_events = [], []
    if pipe.use_loop_graph and pipe.use_cuda_graph:
        pipe.loop_events = loop_events
    else:
        pipe.step_events = step_events
    launches0 = _lib.launch_count
    ms_total = timed(call_resident, args.steps)
    pipe.step_events = pipe.loop_events = None
    clocks = sampler.stop() if rank == 0 else None
    unet_ms = [a.elapsed_time(b) for a, b in step_events]
    for a, b, k in loop_events:
        unet_ms += [a.elapsed_time(b) / k] * k
    eager_launches = _lib.launch_count - launches0
    # graph replays do not pass through the ctypes counter: add their content (counted at capture time)
    gpu_launches = eager_launches + args.steps * pipe.launches_per_call(NUM_STEPS)

    # ---- timed region 2: end to end through the public API with host buffers
    for _ in range(2):
        call_e2e()
    ms_e2e = timed(call_e2e, args.steps)
    # ---- timed region 3: prompt strings through the CLIP tower (text-included e2e), and the script's B = 2 call pattern
    ms_text = ms_script = None
    script_reps = max(2, args.steps)
    if not big:
        for _ in range(2):
            call_text()
        ms_text = timed(call_text, args.steps)
        for _ in range(2):
            call_text(1)
        ms_script = timed(lambda: call_text(1), script_reps)

    vals = [ms_total, ms_e2e, ms_text or 0.0, ms_script or 0.0]
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_text, ms_script = (float(v) for v in t)

    if rank == 0:
        peaks = load_peaks()
        total_images = world * n * args.steps
        value = total_images / (ms_total / 1e3)
        e2e_value = total_images / (ms_e2e / 1e3)
        unet_step_ms = sum(unet_ms) / max(len(unet_ms), 1)
        rows = 2 * n
        # CFG pair: the layers in front of the first cross-attention see the same input in both halves and are evaluated
        # once (`UNet2DConditionModel.forward(cfg_pair=True)`): those FLOPs are not executed, so they are not claimed
        hw_lat = (side // 8) ** 2
        shared_tflop = (2 * hw_lat * 320 * 36 + 2 * (2 * hw_lat * 320 * 2880) + 6 * (2 * hw_lat * 320 * 320) + 4 * hw_lat * hw_lat * 320) / 1e12
        shared_on = bool(getattr(pipe, "cfg_shared_prefix", False)) and n >= 2
        step_tflop = rows * tflop_row - (n * shared_tflop if shared_on else 0.0)
        achieved = step_tflop / (unet_step_ms / 1e3) if unet_ms else None
        workload = ("configs[3]: SD2.1 768x768 (96x96 latents, 9216-token self-attention) UNet+VAE random-init + rank-4 LoRA "
                    "(fused, unmerged), 8 prompts x CFG pair = UNet batch 16, 30 DDPM steps, CFG 5.0, per GPU") if big else \
                   ("configs[1]: SD2.1-base UNet+VAE random-init + rank-4 LoRA (fused, unmerged), "
                    "4 prompts x CFG pair = UNet batch 8, 512x512, 30 DDPM steps, CFG 5.0, per GPU")
        e2e = {"value": e2e_value, "unit": "images/s",
               "h2d_bytes_per_step": int(pe_host.numel() * 4 + ne_host.numel() * 4 + lat_host.numel() * 4),
               "d2h_bytes_per_step": int(n * side * side * 3 * 4)}
        if ms_text:
            e2e["text_included"] = {"value": total_images / (ms_text / 1e3), "unit": "images/s",
                                    "what": "prompt strings (fresh every call) -> CLIP-H tower on the sm_100a kernels -> pipe; D2H of the images"}
            e2e["script_mode"] = {"value": world * script_reps / (ms_script / 1e3), "unit": "images/s",
                                  "what": "one prompt string per call (UNet batch 2), the call pattern of inference_ID-Booth.py:138"}
        line = {
            "metric": "images_per_sec_768x768_30step_cfg5" if big else "images_per_sec_512x512_30step_cfg5",
            "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload,
                       "images_per_step_per_gpu": n, "parallelism": f"dp{world} by image, asynchronous NCCL all_gather of uint8 images",
                       "l2": "inputs larger than L2: ~1.4 GB of activation traffic per UNet forward vs 126 MB L2",
                       "unet_step_ms": unet_step_ms, "unet_steps_timed": len(unet_ms),
                       "unet_frac_of_burst_peak": (achieved / peaks["bf16"]) if achieved else None},
            "e2e": e2e,
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["bf16_sustained"]) if achieved else None,
                         "traffic": None if big else _step_traffic(),
                         "kernel": "UNet step (CFG pair forward + fused CFG/DDPM update) inside the denoising graph: gemm_tc_kernel / "
                                   "attention_rs_kernel dominate; see profiles/",
                         "flops_per_launch": step_tflop * 1e12,
                         "flops_note": (f"{rows} rows x {tflop_row} TFLOP (SURVEY 8d) minus {n} x {shared_tflop:.4f} TFLOP: conv_in, the first "
                                        "ResnetBlock2D and the first transformer up to its cross-attention are identical in both halves of "
                                        "the CFG pair and run once") if shared_on else f"{rows} rows x {tflop_row} TFLOP (SURVEY 8d)",
                         "peak_source": peaks["source"] + ", sustained figure (timed inside a long step)"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_library_baseline and not big:
            try:
                line["library_baseline"] = library_baseline_sample(dev, n)
            except Exception as e:   # a comparator must never take the bench line down
                line["library_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        if world == 1 and not args.no_cpu_baseline and not big:
            threads = os.cpu_count() or 1
            step_s, vae_s = cpu_oracle_sample(6, True, threads)
            mean_step = sum(step_s[1:]) / len(step_s[1:])
            cpu_v = 1.0 / (NUM_STEPS * mean_step + vae_s)
            line["cpu_baseline"] = {"value": cpu_v, "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": "6 CFG denoise steps (UNet B=2, fp32 torch CPU) of one image + 1 VAE decode; "
                                              "mean of steps 2-6 extrapolated x30"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_sweep_config(args):
    """--config 3 (BASELINE.json configs[2]): the ID-Booth dataset-augmentation sweep -- 1024 images at 512x512, 30 steps,
    through the caller path (`faceposegenerator_b200.sweep`: prompt strings -> CLIP tower -> pipe -> PNG files), identities
    sharded over the ranks: STRONG scaling (total work fixed)."""
    import shutil
    import tempfile
    from faceposegenerator_b200 import _lib, sweep
    dist, pipe, world, rank, local, dev = _setup(args)
    root = tempfile.mkdtemp(prefix=f"idb_sweep_r{rank}_")
    try:
        cfg = sweep.synthetic_sweep(root, total_images=args.images, seed=0)
        # warm-up: one identity unit per rank (captures the graphs), not counted
        sweep.run_sweep(cfg, rank=rank, world_size=world, device=str(dev), limit_units=1, batch_prompts=args.batch_prompts)
        shutil.rmtree(os.path.join(root, "out"), ignore_errors=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        from faceposegenerator_b200 import pipeline as _pl
        n0 = _lib.launch_count + _pl.graph_launches
        t0 = time.perf_counter()
        stats = sweep.run_sweep(cfg, rank=rank, world_size=world, device=str(dev), batch_prompts=args.batch_prompts)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt, float(stats["images"])], dtype=torch.float64, device=dev)
        if world > 1:
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dt = float(mx[0])
        done = float(t[1])
        if rank == 0:
            emit({"metric": "images_per_sec_512x512_30step_cfg5", "value": done / dt, "unit": "images/s", "n_gpus": world,
                  "steps": 1, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
                  "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                  "config": {"workload": f"configs[2]: ID-Booth dataset augmentation sweep, {int(done)} images at 512x512, 30 steps, "
                                         "CFG 5.0, prompt strings -> CLIP tower -> pipe -> PNG files (async writer), identities sharded "
                                         f"over {world} GPU(s)", "batch_prompts": args.batch_prompts, "wall_s": dt,
                             "timing": "host wall clock around the whole sweep incl. file writes, max over ranks"},
                  "gpu_launches": int(_lib.launch_count + _pl.graph_launches - n0)})
    finally:
        shutil.rmtree(root, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


_JSON_OUT = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    chatter) is redirected to stderr."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4],
                    help="2 = BASELINE configs[1] (default, the headline), 3 = configs[2] 1024-image sweep, 4 = configs[3] 768x768 B=16")
    ap.add_argument("--images", type=int, default=1024, help="--config 3: images in the sweep")
    ap.add_argument("--batch-prompts", type=int, default=4, help="--config 3: prompts batched per pipe() call (1 = the script's own call pattern)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.config == 3:
        run_sweep_config(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
