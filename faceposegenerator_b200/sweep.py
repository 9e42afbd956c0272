"""The caller side of the hot path: the identity x model x prompt sweep that the reference drives
from `inference_ID-Booth.py:17-156`, as a library + CLI that one process per GPU can run over a shard of the identities.

`plan()` is the host logic of the script restated as a pure function: identity order (`os.listdir` of the first model,
entries with ".json" dropped, natural sort, `:69-71`), the python-RNG prompt schedule (`set_seed(0)` `:67`, one
`random.sample` per identity `:94`, one `random.choice` per (model, prompt) for the pose `:126`), prompt assembly
(`:113-135`), LoRA paths (`:98`), per-identity generator seed (`:111`) and output file names (`:54-59,100,142-144,
151-156`).  It is pinned against a log of the reference script itself (tests/golden/inference_script_golden.json).
Because the schedule depends only on the seed and the directory listing, every rank replays it identically and takes
identities `rank, rank + world, ...` (`parallel.shard_units`): no communication inside the sweep, images are written
rank-locally (PNG per image, one comparison JPG per identity), one barrier at the end.

`run_identity()` issues, per (identity, model), exactly the statements of `:103-144`: `from_pretrained(...).to(device)`
(the base weights come from the pipeline's component cache, so only the 128 rank-4 adapters are swapped), scheduler
swap, `load_lora_weights`, a fresh `torch.Generator(device).manual_seed(id_number)`, one `pipe(...)` call per prompt
sample.  Additions that do not change any image: PNG / JPG encoding runs on writer threads (`AsyncImageWriter`) so it
overlaps the next image's denoising loop, and `skip_existing=True` restarts an interrupted sweep -- finished
(identity, model) runs are skipped outright, and inside a partly finished run the generator is advanced by the draws
the skipped images would have consumed (1 initial latent + one per scheduler step with t > 0), so the remaining images
are bit-identical to an uninterrupted run.

`batch_prompts=B > 1` is the fast mode: B consecutive prompts of one (identity, model) run go through ONE `pipe()` call
(UNet batch 2B instead of 2).  The generator draws of an image (1 initial latent + 1 per step with t > 0, each
`randn((1, 4, h, w))` in the pipeline dtype) do not depend on any image content, so they are pre-drawn in the script's
order into per-image tapes (`draw_tape`) and handed to the pipeline: every image starts from, and is perturbed by,
exactly the numbers the unbatched script would have used.  The last, partial batch of a run is padded to B rows so one
captured graph serves every call.  Images are bit-reproducible across runs, ranks and prompt groupings at a given B;
against `batch_prompts=1` they agree to the bf16 noise floor, not bit for bit (the GEMM tile / split-K schedule depends
on the batch), which is why B = 1 (the script's own call pattern) stays the default.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import re
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from itertools import product
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .parallel import shard_units

BACKGROUNDS = ["", "forest", "city street", "beach", "office", "bus", "laboratory", "factory", "construction site",
               "hospital", "night club"]
AGE_PHASES = ["", "young", "middle-aged", "old"]


@dataclass
class SweepConfig:
    """Constants of `inference_ID-Booth.py:17-82` (same names, same defaults)."""
    num_samples_per_prompt: int = 1
    num_prompts: int = 21
    add_gender: bool = True
    add_pose: bool = True
    add_age: bool = False
    add_background: bool = True
    do_not_use_negative_prompt: bool = False   # only renames the output folder in the reference (`:59`)
    use_non_finetuned: bool = False
    seed: int = 0
    guidance_scale: float = 5.0
    num_inference_steps: int = 30
    folder_of_models: str = "Trained_LoRA_Models"
    models_to_test: Tuple[str, ...] = ("DreamBooth", "PortraitBooth", "ID-Booth")
    checkpoint: str = "checkpoint-31-6400"
    model_architecture: str = "stabilityai/stable-diffusion-2-1-base"
    width: int = 512
    height: int = 512
    negative_prompt: str = ("cartoon, cgi, render, illustration, painting, drawing, black and white, "
                            "bad body proportions, landscape")
    original_prompt: str = "face portrait photo of sks person"
    gender_file: str = "tufts_gender_dict.json"
    output_root: str = ""      # prefix of `Generated_Samples/...` ("" = the script's cwd-relative tree)


@dataclass
class Job:
    i: int            # prompt index
    j: int            # sample index of the prompt
    prompt: str
    png_path: str


@dataclass
class ModelRun:
    model_name: str
    lora_path: str
    sample_dir: str
    jobs: List[Job] = field(default_factory=list)


@dataclass
class IdentityUnit:
    id_number: int    # position in the sorted identity list = seed of the identity's generator
    which_id: str
    runs: List[ModelRun]
    comparison_path: str
    comparison_nrow: int


def natural_keys(text: str):
    """Human sort key ("2" < "10"), the order `ids.sort(key=natural_keys)` produces at `:71`."""
    return [int(tok) if tok.isdigit() else tok for tok in re.split(r"(\d+)", text)]


def prompt_combinations(cfg: SweepConfig) -> list:
    bg = [f"{b} background" if b != "" else "" for b in BACKGROUNDS]
    if cfg.add_age and cfg.add_background:
        return list(product(AGE_PHASES, bg))
    if cfg.add_background:
        return list(bg[1:] * 10) if cfg.num_prompts == 100 else list([""] + bg[1:] * 2)
    if cfg.add_age:
        return list(AGE_PHASES * 6)
    return list([""] * cfg.num_prompts)


def folder_output(cfg: SweepConfig) -> str:
    out = os.path.join(cfg.output_root, "Generated_Samples/FacePortrait_Photo_21") if cfg.output_root else \
        "Generated_Samples/FacePortrait_Photo_21"
    for flag, suffix in ((cfg.add_gender, "_Gender"), (cfg.add_pose, "_Pose"), (cfg.add_age, "_Age"),
                         (cfg.add_background, "_Background"), (cfg.do_not_use_negative_prompt, "_NoNegPrompt")):
        if flag:
            out += suffix
    return out


def list_identities(cfg: SweepConfig) -> List[str]:
    ids = [i for i in os.listdir(os.path.join(cfg.folder_of_models, cfg.models_to_test[0])) if ".json" not in i]
    ids.sort(key=natural_keys)
    return ids


def load_gender_dict(cfg: SweepConfig) -> Dict[str, str]:
    if not cfg.add_gender:
        return {}
    with open(cfg.gender_file, "r") as fp:
        return json.load(fp)


def _assemble_prompt(cfg: SweepConfig, additions, gender: Optional[str], side_pose: bool) -> str:
    prompt = cfg.original_prompt
    if cfg.add_age:
        if isinstance(additions, str):
            age = additions
        else:
            age, additions = additions[0], additions[1:]
        if age != "":
            prompt = prompt.replace(" sks person", f" {age} sks person")
    if cfg.add_gender:
        prompt = prompt.replace(" sks person", f" {gender} sks person")
    if side_pose:
        prompt = prompt.replace("portrait", "side-portrait")
    if cfg.add_background:
        if isinstance(additions, str):
            prompt += f", {additions}"
        else:
            for extra in additions:
                if extra != "":
                    prompt += f", {extra}"
    return prompt


def plan(cfg: SweepConfig, ids: Sequence[str], gender_dict: Dict[str, str]) -> List[IdentityUnit]:
    """The whole sweep as data.  Pure function of (cfg, ids, gender_dict): replays the script's python RNG."""
    rng = random.Random(cfg.seed)          # `set_seed(seed)` seeds the module-level generator with the same int
    combos = prompt_combinations(cfg)
    out_root = folder_output(cfg)
    arch = cfg.model_architecture.split("/")[1]
    units = []
    for id_number, which_id in enumerate(ids):
        gender = None
        if cfg.add_gender:
            gender = {"M": "male", "F": "female"}.get(gender_dict[which_id], gender_dict[which_id])
        schedule = rng.sample(combos, cfg.num_prompts)
        runs = []
        for model_name in cfg.models_to_test:
            sample_dir = f"{os.path.join(out_root, model_name)}/{which_id}_{cfg.checkpoint}_{arch}"
            run = ModelRun(model_name, os.path.join(cfg.folder_of_models, model_name, which_id, cfg.checkpoint), sample_dir)
            for i in range(cfg.num_prompts):
                side = bool(cfg.add_pose and rng.choice([True, False]))
                prompt = _assemble_prompt(cfg, schedule[i], gender, side)
                for j in range(cfg.num_samples_per_prompt):
                    run.jobs.append(Job(i, j, prompt, f"{sample_dir}/{i}_{j}_{prompt}.png"))
            runs.append(run)
        comparison = f"{os.path.join(out_root, 'Comparison')}/{which_id}_{cfg.checkpoint}_{arch}_{cfg.guidance_scale}.jpg"
        units.append(IdentityUnit(id_number, which_id, runs, comparison, cfg.num_prompts * cfg.num_samples_per_prompt))
    return units


# ---------------------------------------------------------------------------------------------- image writer
def _atomic_save(save_fn, tensor, fp: str, kwargs) -> None:
    """Write to `<name>.tmp<ext>` and rename into place: a crash or a full disk never leaves a truncated file under the
    final name (which `skip_existing` would take for a finished image)."""
    os.makedirs(os.path.dirname(fp) or ".", exist_ok=True)
    base, ext = os.path.splitext(fp)
    tmp = f"{base}.tmp{ext}"
    try:
        save_fn(tensor, fp=tmp, **kwargs)
        os.replace(tmp, fp)
    except BaseException:
        try:
            os.remove(tmp)
        except OSError:
            pass
        raise


def _existing_image(path: str) -> bool:
    """`skip_existing`: a file counts as done only if it decodes (a leftover from an older, non-atomic writer or a
    damaged disk is regenerated instead of crashing the resume)."""
    if not os.path.isfile(path):
        return False
    try:
        from PIL import Image
        with Image.open(path) as im:
            im.verify()
        return True
    except Exception:
        return False


class AsyncImageWriter:
    """`torchvision.utils.save_image` (the call at `:144,156`) on worker threads: PNG / JPG encoding of image k overlaps the
    denoising loop of image k + 1.  At most `max_pending` images are held; errors surface in `close()`."""

    def __init__(self, workers: int = 4, max_pending: int = 64, save_fn: Optional[Callable] = None):
        if save_fn is None:
            from torchvision.utils import save_image as save_fn
        self._save = save_fn
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers), thread_name_prefix="idb-png")
        self._slots = threading.Semaphore(max_pending)
        self._futures = []
        self.written = 0

    def _job(self, tensor, fp, kwargs):
        try:
            _atomic_save(self._save, tensor, fp, kwargs)
        finally:
            self._slots.release()

    def _reap(self) -> None:
        """Raise the first error of a finished write NOW (a bad path must not surface hours later in close())."""
        keep = []
        for f in self._futures:
            if f.done():
                f.result()
            else:
                keep.append(f)
        self._futures = keep

    def save(self, tensor: torch.Tensor, fp: str, **kwargs) -> None:
        self._reap()
        self._slots.acquire()
        self._futures.append(self._pool.submit(self._job, tensor, fp, kwargs))
        self.written += 1

    def close(self) -> None:
        self._pool.shutdown(wait=True)
        futures, self._futures = self._futures, []
        for f in futures:
            f.result()


class _SyncWriter(AsyncImageWriter):
    def __init__(self, save_fn: Optional[Callable] = None):
        if save_fn is None:
            from torchvision.utils import save_image as save_fn
        self._save, self.written = save_fn, 0

    def save(self, tensor, fp, **kwargs):
        _atomic_save(self._save, tensor, fp, kwargs)
        self.written += 1

    def close(self):
        pass


# ---------------------------------------------------------------------------------------------- execution
def draws_per_image(scheduler, num_inference_steps: int) -> int:
    """Generator draws one `pipe()` call consumes: the initial latent (`prepare_latents`) + one variance-noise tensor per
    `DDPMScheduler.step` with t > 0."""
    scheduler.set_timesteps(num_inference_steps)
    return 1 + sum(1 for t in scheduler.timesteps.tolist() if t > 0)


def _advance_generator(generator, n_draws: int, cfg: SweepConfig, device, dtype) -> None:
    for _ in range(n_draws):   # same shape / dtype / device as the draws the skipped image would have made
        torch.randn((1, 4, cfg.height // 8, cfg.width // 8), generator=generator, device=device, dtype=dtype)


def _read_image(path: str) -> torch.Tensor:
    from PIL import Image
    with Image.open(path) as im:
        arr = np.asarray(im.convert("RGB"), dtype=np.float32) / 255.0
    return torch.from_numpy(arr)[None]


def draw_tape(generator, scheduler, cfg: SweepConfig, device, dtype) -> torch.Tensor:
    """The generator draws of ONE `pipe()` call of the script, in its order: [1 + steps, 1, 4, h, w] (index 0 = initial
    latent, 1 + i = variance noise of step i; a step with t == 0 draws nothing and gets zeros)."""
    shape = (1, 4, cfg.height // 8, cfg.width // 8)
    scheduler.set_timesteps(cfg.num_inference_steps)
    rows = [torch.randn(shape, generator=generator, device=device, dtype=dtype)]
    for t in scheduler.timesteps.tolist():
        rows.append(torch.randn(shape, generator=generator, device=device, dtype=dtype) if t > 0
                    else torch.zeros(shape, device=device, dtype=dtype))
    return torch.stack(rows)


def run_identity(unit: IdentityUnit, cfg: SweepConfig, device: str, writer: AsyncImageWriter, skip_existing: bool = False,
                 pipeline_cls=None, scheduler_cls=None, torch_dtype=torch.float16, batch_prompts: int = 1) -> Dict[str, int]:
    """`inference_ID-Booth.py:96-156` for one identity.  Returns counters {generated, skipped}."""
    if pipeline_cls is None or scheduler_cls is None:
        from . import DDPMScheduler, StableDiffusionPipeline
        pipeline_cls, scheduler_cls = pipeline_cls or StableDiffusionPipeline, scheduler_cls or DDPMScheduler
    stats = {"generated": 0, "skipped": 0}
    comparison: List[Optional[torch.Tensor]] = []
    for run in unit.runs:
        done = [skip_existing and _existing_image(job.png_path) for job in run.jobs]
        if all(done):   # the generator is re-created per (identity, model): a finished run leaves no state behind
            comparison += [_read_image(job.png_path) for job in run.jobs]
            stats["skipped"] += len(run.jobs)
            continue
        pipe = pipeline_cls.from_pretrained(cfg.model_architecture, torch_dtype=torch_dtype).to(device)
        pipe.scheduler = scheduler_cls.from_pretrained(cfg.model_architecture, subfolder="scheduler")
        if not cfg.use_non_finetuned:
            pipe.load_lora_weights(run.lora_path)
        pipe.set_progress_bar_config(disable=True)
        os.makedirs(os.path.dirname(run.sample_dir), exist_ok=True)
        generator = torch.Generator(device=device).manual_seed(unit.id_number)
        n_draws = draws_per_image(pipe.scheduler, cfg.num_inference_steps) if any(done) else 0
        call_kw = dict(negative_prompt=cfg.negative_prompt, output_type="np", num_inference_steps=cfg.num_inference_steps,
                       guidance_scale=cfg.guidance_scale, width=cfg.width, height=cfg.height)
        pending: List[Tuple[Job, int, torch.Tensor]] = []      # batched mode: (job, slot in `comparison`, tape)

        def flush():
            if not pending:
                return
            m = len(pending)
            prompts = [j.prompt for j, _, _ in pending] + [pending[-1][0].prompt] * (batch_prompts - m)
            tapes = [t for _, _, t in pending] + [torch.zeros_like(pending[0][2])] * (batch_prompts - m)
            out = pipe(prompt=prompts, generator=generator, noise_tape=torch.cat(tapes, dim=1), **call_kw)   # (the tape replaces every draw)
            images = torch.Tensor(out.images)
            for k, (job, slot, _) in enumerate(pending):
                comparison[slot] = images[k:k + 1]
                os.makedirs(run.sample_dir, exist_ok=True)
                writer.save(torch.permute(images[k:k + 1], (0, 3, 1, 2)), job.png_path)
                stats["generated"] += 1
            pending.clear()

        for job, have in zip(run.jobs, done):
            if have:
                _advance_generator(generator, n_draws, cfg, device, torch_dtype)
                comparison.append(_read_image(job.png_path))
                stats["skipped"] += 1
                continue
            if batch_prompts > 1:
                comparison.append(None)
                pending.append((job, len(comparison) - 1, draw_tape(generator, pipe.scheduler, cfg, device, torch_dtype)))
                if len(pending) == batch_prompts:
                    flush()
                continue
            output = pipe(prompt=job.prompt, generator=generator, **call_kw)
            output = torch.Tensor(output.images)
            comparison.append(output)
            os.makedirs(run.sample_dir, exist_ok=True)
            writer.save(torch.permute(output, (0, 3, 1, 2)), job.png_path)
            stats["generated"] += 1
        flush()
    if stats["generated"] or not (skip_existing and _existing_image(unit.comparison_path)):
        images = torch.permute(torch.cat(comparison), (0, 3, 1, 2))
        writer.save(images, unit.comparison_path, nrow=unit.comparison_nrow, padding=0)
    return stats


def run_sweep(cfg: SweepConfig, rank: int = 0, world_size: int = 1, device: Optional[str] = None,
              skip_existing: bool = False, writer: Optional[AsyncImageWriter] = None, max_identities: Optional[int] = None,
              pipeline_cls=None, scheduler_cls=None, set_seed_fn: Optional[Callable] = None,
              writer_threads: int = 4, batch_prompts: int = 1, limit_units: Optional[int] = None) -> Dict[str, float]:
    """One rank's share of the sweep: identities `rank, rank + world_size, ...` of the planned list
    (`limit_units`: only the first k of this rank's identities, e.g. a warm-up pass)."""
    if batch_prompts < 1:
        raise ValueError("batch_prompts must be >= 1")
    if set_seed_fn is None:
        def set_seed_fn(seed):   # accelerate.utils.set_seed (`:67`): python, numpy, torch (CPU + CUDA) generators
            random.seed(seed)
            np.random.seed(seed)
            torch.manual_seed(seed)
            if torch.cuda.is_available():
                torch.cuda.manual_seed_all(seed)
    set_seed_fn(cfg.seed)
    ids = list_identities(cfg)
    units = plan(cfg, ids, load_gender_dict(cfg))
    if max_identities is not None:
        units = units[:max_identities]
    if device is None:
        device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
    own_writer = writer is None
    writer = writer or AsyncImageWriter(writer_threads)
    t0 = time.time()
    totals = {"generated": 0, "skipped": 0, "identities": 0}
    mine = shard_units(len(units), rank, world_size)
    if limit_units is not None:
        mine = mine[:limit_units]
    try:
        for u in mine:
            st = run_identity(units[u], cfg, device, writer, skip_existing, pipeline_cls, scheduler_cls,
                              batch_prompts=batch_prompts)
            totals["generated"] += st["generated"]
            totals["skipped"] += st["skipped"]
            totals["identities"] += 1
    finally:
        if own_writer:
            writer.close()
    totals["images"] = totals["generated"]
    totals["seconds"] = time.time() - t0
    return totals


def synthetic_sweep(root: str, total_images: int = 1024, seed: int = 0, num_prompts: int = 16,
                    model_name: str = "ID-Booth") -> SweepConfig:
    """BASELINE.json configs[2] as a directory tree the sweep can run offline: `total_images / num_prompts` identities,
    each with its own synthetic rank-4 adapter checkpoint in the reference's on-disk layout
    (`<folder_of_models>/<model>/<id>/<checkpoint>/pytorch_lora_weights.safetensors`, `train_ID-Booth.py:696-720`), one
    model variant, a gender file; outputs go under `<root>/out`."""
    from .weights import random_lora, save_lora_weights
    n_ids = max(1, total_images // num_prompts)
    cfg = SweepConfig(num_prompts=num_prompts, models_to_test=(model_name,), seed=seed,
                      folder_of_models=os.path.join(root, "Trained_LoRA_Models"),
                      gender_file=os.path.join(root, "gender.json"), output_root=os.path.join(root, "out"))
    for i in range(n_ids):
        d = os.path.join(cfg.folder_of_models, model_name, str(i + 1), cfg.checkpoint)
        if not os.path.isfile(os.path.join(d, "pytorch_lora_weights.safetensors")):
            save_lora_weights(d, random_lora(seed=1000 + i))
    with open(cfg.gender_file, "w") as f:
        json.dump({str(i + 1): ("F" if i % 2 else "M") for i in range(n_ids)}, f)
    return cfg


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="ID-Booth image sweep (inference_ID-Booth.py) on B200, one process per GPU")
    ap.add_argument("--num-prompts", type=int, default=21)
    ap.add_argument("--num-samples-per-prompt", type=int, default=1)
    ap.add_argument("--folder-of-models", default="Trained_LoRA_Models")
    ap.add_argument("--models", nargs="+", default=["DreamBooth", "PortraitBooth", "ID-Booth"])
    ap.add_argument("--checkpoint", default="checkpoint-31-6400")
    ap.add_argument("--skip-existing", action="store_true")
    ap.add_argument("--max-identities", type=int, default=None)
    ap.add_argument("--writer-threads", type=int, default=4)
    ap.add_argument("--model-architecture", default=SweepConfig.model_architecture,
                    help="hub id (resolved in the local Hugging Face cache) or a local snapshot directory")
    ap.add_argument("--batch-prompts", type=int, default=1,
                    help="prompts per pipe() call; 1 = the script's own call pattern (bit-identical to it), 4-8 = fast mode")
    ap.add_argument("--allow-random-weights", action="store_true",
                    help="benchmarking only: random-init stand-ins when the checkpoint is not available (images are noise)")
    a = ap.parse_args(argv)
    if a.allow_random_weights:
        os.environ["IDB_ALLOW_RANDOM_WEIGHTS"] = "1"
    cfg = SweepConfig(num_prompts=a.num_prompts, num_samples_per_prompt=a.num_samples_per_prompt,
                      folder_of_models=a.folder_of_models, models_to_test=tuple(a.models), checkpoint=a.checkpoint,
                      model_architecture=a.model_architecture)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    try:
        totals = run_sweep(cfg, rank, world, skip_existing=a.skip_existing, writer_threads=a.writer_threads,
                           max_identities=a.max_identities, batch_prompts=a.batch_prompts)
    except BaseException:
        # a failing rank must not leave the others waiting in the final barrier: tear the whole job down
        if dist is not None:
            import traceback
            traceback.print_exc()
            os._exit(1)        # torchrun kills the remaining ranks when one exits non-zero
        raise
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    print(json.dumps({"rank": rank, "world_size": world, **totals}))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
