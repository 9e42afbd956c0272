#!/usr/bin/env python
"""bench.py -- 512x512, 30-step, CFG-5.0 DDPM image generation throughput (images/s) of the
ID-Booth SD2.1 hot path on N B200s (one process per GPU, data-parallel by image).

    python bench.py --gpus 1 --steps K --warmup W            # this framework
    python bench.py --impl reference --steps K --warmup W    # the restated reference path on host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: one pipeline call that
generates 4 images (UNet batch 8 = 4 prompts x CFG pair, BASELINE.json configs[1]): 30 x
[UNet -> fused CFG + DDPMScheduler.step] + VAE decode + post-process.
  value : images/s with prompt embeddings and the noise tape already resident in HBM
  e2e   : same metric through the public pipe(...) call with HOST (pinned) prompt embeddings /
          initial latents copied H2D and the decoded images read back D2H inside the timed region
  roofline : the UNet step (the graph launch that dominates the loop): 6.4432 TFLOP per B=8 forward
           / its mean CUDA-event duration inside the timed region, vs the measured
          sustained bf16 peak of MEASURED_PEAKS.json
  cpu_baseline : the fp32 oracle (oracle/sd21.py = the torch CPU ops diffusers would dispatch to)
          timed on this box's host cores on a bounded sample (N=1, rank 0 only)
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
VAE_TFLOP_PER_IMAGE = 2.5145
IMAGES_PER_CALL = 4              # 4 prompts x CFG pair -> UNet batch 8
NUM_STEPS = 30
GUIDANCE = 5.0


def _step_traffic():
    """DRAM bytes of one UNet step from the committed ncu capture (profiles/r01_step_traffic.json), or None."""
    fn = os.path.join(ROOT, "profiles", "r01_step_traffic.json")
    try:
        with open(fn) as f:
            return float(json.load(f)["traffic_bytes"])
    except Exception:
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
              f"extrapolated to {NUM_STEPS} steps")
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
def run_native(args):
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

    model = "stabilityai/stable-diffusion-2-1-base"
    pipe = StableDiffusionPipeline.from_pretrained(model, torch_dtype=torch.bfloat16, allow_random_weights=True).to(dev)
    pipe.scheduler = DDPMScheduler.from_pretrained(model, subfolder="scheduler")
    pipe.load_lora_weights(random_lora(seed=0))          # synthetic "trained" rank-4 adapters, fused unmerged
    pipe.set_progress_bar_config(disable=True)

    from faceposegenerator_b200.parallel import gather_images, shard_units, unit_seed
    n = IMAGES_PER_CALL
    # one step = world * n image units; rank r generates units r, r + G, ... (weak scaling: n units per GPU per step)
    my_units = shard_units(world * n, rank, world)
    assert len(my_units) == n
    g = torch.Generator().manual_seed(unit_seed(my_units[0], 1000))
    # synthetic context (text encoder is outside the timed hot path): N(0,1) prompt / negative embeddings
    pe_host = torch.randn(n, 77, 1024, generator=g).pin_memory()
    ne_host = torch.randn(n, 77, 1024, generator=g).pin_memory()
    lat_host = torch.randn(n, 4, 64, 64, generator=g).pin_memory()
    pe_dev, ne_dev = pe_host.to(dev), ne_host.to(dev)
    tape_dev = torch.randn(1 + NUM_STEPS, n, 4, 64, 64, generator=g).to(dev)
    gen = torch.Generator(device=dev).manual_seed(rank)
    gathered = torch.empty((world * n, 512, 512, 3), dtype=torch.uint8, device=dev) if world > 1 else None

    def call_resident():
        """inputs resident in HBM; result (uint8 images) stays on device; final NCCL gather when N > 1"""
        out = pipe(prompt_embeds=pe_dev, negative_prompt_embeds=ne_dev, num_inference_steps=NUM_STEPS,
                   guidance_scale=GUIDANCE, height=512, width=512, output_type="pt", noise_tape=tape_dev)
        img = (out.images.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).contiguous()
        if world > 1:
            gather_images(img, world * n, rank, world, out=gathered)   # one all_gather_into_tensor over NCCL / NVLink
        return img

    def call_e2e():
        """public API with HOST buffers: H2D of embeddings + initial latents, D2H of the images"""
        out = pipe(prompt_embeds=pe_host.to(dev, non_blocking=True), negative_prompt_embeds=ne_host.to(dev, non_blocking=True),
                   latents=lat_host.to(dev, non_blocking=True), generator=gen, num_inference_steps=NUM_STEPS,
                   guidance_scale=GUIDANCE, height=512, width=512, output_type="np")
        return out.images

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        call_resident()
    barrier()

    # ---- timed region 1: resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_events = []
    pipe.step_events = step_events   # CUDA events on the launching stream around every denoise-step graph launch
    launches0 = _lib.launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        call_resident()
    e1.record()
    barrier()
    pipe.step_events = None
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    unet_ms = [a.elapsed_time(b) for a, b in step_events]
    eager_launches = _lib.launch_count - launches0
    # graph replays do not pass through the ctypes counter: add their content (counted at capture time)
    gpu_launches = eager_launches + args.steps * pipe.launches_per_call(NUM_STEPS)

    # ---- timed region 2: end to end through the public API with host buffers
    for _ in range(2):
        call_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        call_e2e()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks = load_peaks()
        total_images = world * n * args.steps
        value = total_images / (ms_total / 1e3)
        e2e_value = total_images / (ms_e2e / 1e3)
        unet_step_ms = sum(unet_ms) / max(len(unet_ms), 1)
        rows = 2 * n
        achieved = rows * UNET_TFLOP_PER_ROW / (unet_step_ms / 1e3) if unet_ms else None
        line = {
            "metric": "images_per_sec_512x512_30step_cfg5", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: SD2.1-base UNet+VAE random-init + rank-4 LoRA (fused, unmerged), "
                                   "4 prompts x CFG pair = UNet batch 8, 512x512, 30 DDPM steps, CFG 5.0, per GPU",
                       "images_per_step_per_gpu": n, "parallelism": f"dp{world} by image, final NCCL all_gather of uint8 images",
                       "l2": "inputs larger than L2: ~1.4 GB of activation traffic per UNet forward vs 126 MB L2",
                       "unet_step_ms": unet_step_ms, "unet_steps_timed": len(unet_ms),
                       "unet_frac_of_burst_peak": (achieved / peaks["bf16"]) if achieved else None},
            "e2e": {"value": e2e_value, "unit": "images/s",
                    "h2d_bytes_per_step": int(pe_host.numel() * 4 + ne_host.numel() * 4 + lat_host.numel() * 4),
                    "d2h_bytes_per_step": int(n * 512 * 512 * 3 * 4)},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["bf16_sustained"]) if achieved else None, "traffic": _step_traffic(),
                         "kernel": "UNet step graph (gemm_tc_kernel / attention_kernel dominate; see profiles/)",
                         "flops_per_launch": rows * UNET_TFLOP_PER_ROW * 1e12,
                         "peak_source": peaks["source"] + ", sustained figure (timed inside a long step)"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            step_s, vae_s = cpu_oracle_sample(2, True, threads)
            cpu_v = 1.0 / (NUM_STEPS * step_s[-1] + vae_s)
            line["cpu_baseline"] = {"value": cpu_v, "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": "2 CFG denoise steps (UNet B=2, fp32 torch CPU) of one image + 1 VAE decode, "
                                              "second step extrapolated x30"}
        emit(line)
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
