"""Caller side of the hot path: `faceposegenerator_b200.sweep` against a log of the reference's own
`inference_ID-Booth.py` (tests/golden/inference_script_golden.json, produced by executing that script under recording
stubs: tests/golden/make_inference_script_golden.py).  Host logic only -- no GPU, no compute calls."""
import json
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def gold():
    with open(os.path.join(ROOT, "tests", "golden", "inference_script_golden.json")) as f:
        return json.load(f)


@pytest.fixture()
def tree(tmp_path, monkeypatch, gold):
    """The directory layout the reference script walks, in a scratch cwd (the script uses relative paths)."""
    monkeypatch.chdir(tmp_path)
    for m in gold["models"]:
        for i in gold["ids_on_disk"]:
            os.makedirs(os.path.join("Trained_LoRA_Models", m, i, gold["checkpoint"]))
        with open(os.path.join("Trained_LoRA_Models", m, "training_args.json"), "w") as f:
            f.write("{}")
    with open("tufts_gender_dict.json", "w") as f:
        json.dump(gold["genders"], f)
    return tmp_path


def test_config_defaults_are_the_script_constants(gold):
    from faceposegenerator_b200.sweep import SweepConfig, folder_output, prompt_combinations
    cfg, c = SweepConfig(), gold["script_constants"]
    for k in ("num_samples_per_prompt", "num_prompts", "add_gender", "add_pose", "add_age", "add_background", "seed",
              "guidance_scale", "num_inference_steps", "folder_of_models", "model_architecture", "width", "height",
              "negative_prompt", "original_prompt"):
        assert getattr(cfg, k) == c[k], k
    assert list(cfg.models_to_test) == gold["models"] and cfg.checkpoint == gold["checkpoint"]
    assert folder_output(cfg) == c["folder_output"]
    assert prompt_combinations(cfg) == c["all_prompt_combinations"]


def test_plan_matches_reference_script_log(tree, gold):
    from faceposegenerator_b200.sweep import SweepConfig, list_identities, load_gender_dict, plan
    cfg = SweepConfig()
    ids = list_identities(cfg)
    assert ids == gold["script_constants"]["ids"] == ["1", "2", "3_b", "10"]   # natural order, the .json entry dropped
    units = plan(cfg, ids, load_gender_dict(cfg))
    # one pipeline per (identity, model), in the script's order, with its LoRA directory
    runs = [(u, r) for u in units for r in u.runs]
    assert [r.lora_path for _, r in runs] == [p["lora"] for p in gold["pipelines"]]
    # every pipe() call: which pipeline, the identity's generator seed, the prompt
    calls = [[k, u.id_number, job.prompt] for k, (u, r) in enumerate(runs) for job in r.jobs]
    assert calls == gold["calls"]
    # every file the script writes, in order: the identity's PNGs model by model, then its comparison JPG
    files = []
    for u in units:
        files += [job.png_path for r in u.runs for job in r.jobs] + [u.comparison_path]
    assert files == [s[0] for s in gold["saved"]]
    comp = [s for s in gold["saved"] if s[0].endswith(".jpg")]
    assert all(s[2] == {"nrow": u.comparison_nrow, "padding": 0} for s, u in zip(comp, units))


class _RecordingPipeline:
    """Same recording stub as the golden generator, plus a deterministic fake image that depends on every generator draw
    the real pipeline would make (1 initial latent + 30 step noises, fp16, [1, 4, 64, 64])."""
    log = None

    def __init__(self, arch, kwargs):
        self.rec = {"from_pretrained": arch, "torch_dtype": str(kwargs.get("torch_dtype")), "to": None, "scheduler": None,
                    "scheduler_args": None, "lora": None, "progress_bar": None}
        self.log["pipelines"].append(self.rec)

    @classmethod
    def from_pretrained(cls, arch, **kwargs):
        return cls(arch, kwargs)

    def to(self, device):
        self.rec["to"] = "cuda:0"   # the test runs the host logic on a CPU generator
        return self

    def __setattr__(self, name, value):
        if name == "scheduler":
            self.rec["scheduler"], self.rec["scheduler_args"] = type(value).__name__, getattr(value, "args", None)
        object.__setattr__(self, name, value)

    def load_lora_weights(self, path, **kwargs):
        self.rec["lora"] = path

    def set_progress_bar_config(self, **kwargs):
        self.rec["progress_bar"] = kwargs

    def __call__(self, **kwargs):
        gen = kwargs.pop("generator")
        tape = kwargs.pop("noise_tape", None)
        self.log["calls"].append([len(self.log["pipelines"]) - 1, gen.initial_seed(), kwargs["prompt"]])
        self.log["kwargs"].append({k: v for k, v in kwargs.items() if k != "prompt"})
        if tape is not None:     # batched mode: the caller pre-drew every image's 31 tensors ([31, n, 4, 64, 64])
            assert tape.shape[0] == 31 and tape.shape[1] == len(kwargs["prompt"]) and tape.dtype == torch.float16
            accs = [sum(float(tape[d, k].float().mean()) for d in range(31)) for k in range(tape.shape[1])]
        else:
            acc = 0.0
            for _ in range(31):
                acc += float(torch.randn((1, 4, 64, 64), generator=gen, device="cpu", dtype=torch.float16).float().mean())
            accs = [acc]
        img = np.stack([np.full((8, 8, 3), 0.5 + 0.4 * np.tanh(a * 10), dtype=np.float32) for a in accs])
        return type("Out", (), {"images": img})()


class DDPMScheduler:   # named like the class the script instantiates
    @classmethod
    def from_pretrained(cls, arch, **kwargs):
        s = cls()
        s.args = [arch, kwargs]
        return s

    def set_timesteps(self, n):
        self.timesteps = torch.arange(1 + 33 * (n - 1), 0, -33)   # leading spacing of SD2.1: the last timestep is 1

def _run(log, **kw):
    from faceposegenerator_b200.sweep import SweepConfig, _SyncWriter, run_sweep
    _RecordingPipeline.log = log
    saved = log.setdefault("saved", [])

    def save_fn(tensor, fp, **kwargs):
        base, ext = os.path.splitext(fp)                 # the writer saves to <name>.tmp<ext> and renames into place
        assert base.endswith(".tmp"), "images must be written atomically"
        saved.append([base[:-4] + ext, list(tensor.shape), kwargs])
        from torchvision.utils import save_image
        save_image(tensor, fp=fp, **kwargs)
    return run_sweep(SweepConfig(), device="cpu", writer=_SyncWriter(save_fn), pipeline_cls=_RecordingPipeline,
                     scheduler_cls=DDPMScheduler, **kw)


def test_run_sweep_issues_the_reference_scripts_statements(tree, gold):
    log = {"pipelines": [], "calls": [], "kwargs": []}
    totals = _run(log)
    assert totals["generated"] == len(gold["calls"]) == 252 and totals["skipped"] == 0 and totals["identities"] == 4
    assert log["pipelines"] == gold["pipelines"]
    assert log["calls"] == gold["calls"]
    const = {k: v for k, v in gold["call_constant_kwargs"].items() if k != "generator_device"}
    assert all(k == const for k in log["kwargs"])
    assert log["saved"] == gold["saved"]
    made = sorted(os.path.relpath(os.path.join(d, s)) for d, subs, _ in os.walk("Generated_Samples") for s in subs)
    assert made == gold["made_dirs"]


def test_rank_shards_partition_the_sweep(tree, gold):
    from faceposegenerator_b200.sweep import SweepConfig, list_identities, load_gender_dict, plan
    logs = []
    for rank in range(2):
        log = {"pipelines": [], "calls": [], "kwargs": []}
        totals = _run(log, rank=rank, world_size=2)
        assert totals["identities"] == 2
        logs.append(log)
    cfg = SweepConfig()
    units = plan(cfg, list_identities(cfg), load_gender_dict(cfg))
    for rank, log in enumerate(logs):   # rank r: identities r, r + 2 -- prompts and seeds as in the single-process sweep
        want = [[u.id_number, job.prompt] for u in units[rank::2] for r in u.runs for job in r.jobs]
        assert [[c[1], c[2]] for c in log["calls"]] == want
    files = sorted(s[0] for log in logs for s in log["saved"])
    assert files == sorted(s[0] for s in gold["saved"])


def test_skip_existing_resumes_with_identical_images(tree):
    def read(path):
        with open(path, "rb") as f:
            return f.read()
    first = {"pipelines": [], "calls": [], "kwargs": []}
    _run(first, max_identities=1)
    pngs = [s[0] for s in first["saved"] if s[0].endswith(".png")]
    before = {p: read(p) for p in pngs}
    assert len(set(before.values())) > 10     # the fake images depend on the generator state
    # an interrupted sweep: the second model's run lost its images from the 6th on, the third model's run is gone
    lost = pngs[21 + 5:21 * 2] + pngs[21 * 2:]
    for p in lost:
        os.remove(p)
    again = {"pipelines": [], "calls": [], "kwargs": []}
    totals = _run(again, max_identities=1, skip_existing=True)
    assert totals["generated"] == len(lost) and totals["skipped"] == 63 - len(lost)
    assert len(again["pipelines"]) == 2       # the finished first run builds no pipeline at all
    assert [c[2] for c in again["calls"]] == [c[2] for c in first["calls"]][21 + 5:]
    assert {p: read(p) for p in pngs} == before   # the generator was advanced past the skipped images
    # nothing left to do: no pipeline, no image, the comparison image stays
    idle = {"pipelines": [], "calls": [], "kwargs": []}
    totals = _run(idle, max_identities=1, skip_existing=True)
    assert totals["generated"] == 0 and idle["pipelines"] == [] and idle["saved"] == []


def test_batched_sweep_feeds_every_image_the_scripts_own_generator_draws(tree):
    """`batch_prompts=4`: prompts go through the pipeline four at a time (last call of a run padded to four rows), and
    every image still gets exactly the 31 generator draws the one-prompt-per-call script gives it -- the fake image is
    a function of those draws, so the files must be byte-identical to the unbatched sweep, also after a resume."""
    def read(path):
        with open(path, "rb") as f:
            return f.read()
    first = {"pipelines": [], "calls": [], "kwargs": []}
    _run(first, max_identities=1)
    files = [s[0] for s in first["saved"]]
    before = {p: read(p) for p in files}
    for p in files:
        os.remove(p)
    b = {"pipelines": [], "calls": [], "kwargs": []}
    totals = _run(b, max_identities=1, batch_prompts=4)
    assert totals["generated"] == 63
    assert [s[0] for s in b["saved"]] == files                      # same files, same order
    assert {p: read(p) for p in files} == before
    assert len(b["calls"]) == 3 * 6 and all(len(c[2]) == 4 for c in b["calls"])      # 21 prompts = 5 x 4 + 1 (padded)
    flat = [p for c in b["calls"][:6] for p in c[2]]
    assert flat[:21] == [c[2] for c in first["calls"][:21]] and flat[21:] == [flat[20]] * 3
    # resume in batched mode: images 3..9 of the second model's run are lost
    pngs = [p for p in files if p.endswith(".png")]
    lost = pngs[21 + 3:21 + 10]
    for p in lost:
        os.remove(p)
    again = {"pipelines": [], "calls": [], "kwargs": []}
    totals = _run(again, max_identities=1, batch_prompts=4, skip_existing=True)
    assert totals["generated"] == 7 and {p: read(p) for p in files} == before


def test_async_writer_writes_everything_and_reports_errors(tmp_path):
    from faceposegenerator_b200.sweep import AsyncImageWriter
    w = AsyncImageWriter(workers=3, max_pending=4)
    for k in range(12):
        w.save(torch.full((1, 3, 16, 16), k / 12.0), str(tmp_path / "a" / f"{k}.png"))
    w.close()
    assert sorted(os.listdir(tmp_path / "a")) == sorted(f"{k}.png" for k in range(12))

    def boom(tensor, fp, **kw):
        raise OSError("disk full")
    w = AsyncImageWriter(workers=1, save_fn=boom)
    w.save(torch.zeros(1, 3, 4, 4), str(tmp_path / "b.png"))
    with pytest.raises(OSError):
        w.close()
    # an error surfaces at the NEXT save, not only at the end of the sweep; nothing is left under the final name
    w = AsyncImageWriter(workers=1, save_fn=boom)
    w.save(torch.zeros(1, 3, 4, 4), str(tmp_path / "c.png"))
    import time
    time.sleep(0.3)
    with pytest.raises(OSError):
        w.save(torch.zeros(1, 3, 4, 4), str(tmp_path / "d.png"))
    assert not os.path.exists(tmp_path / "c.png") and not os.path.exists(tmp_path / "c.tmp.png")

    def half(tensor, fp, **kw):          # a writer that dies mid-file
        with open(fp, "wb") as f:
            f.write(b"\x89PNG truncated")
        raise KeyboardInterrupt
    from faceposegenerator_b200.sweep import _atomic_save, _existing_image
    with pytest.raises(KeyboardInterrupt):
        _atomic_save(half, torch.zeros(1), str(tmp_path / "e.png"), {})
    assert not os.path.exists(tmp_path / "e.png")
    with open(tmp_path / "f.png", "wb") as f:
        f.write(b"\x89PNG truncated")
    assert not _existing_image(str(tmp_path / "f.png"))      # skip_existing regenerates an undecodable file


# ---------------------------------------------------------------------------------------------- N > 1 over gloo
def _gloo_worker(rank, world, port, cwd, q):
    import torch.distributed as dist
    os.chdir(cwd)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        log = {"pipelines": [], "calls": [], "kwargs": []}
        totals = _run(log, rank=rank, world_size=world)
        dist.barrier()                                   # the sweep's only synchronisation point (sweep.main)
        files = [None] * world
        dist.all_gather_object(files, [s[0] for s in log["saved"]])
        q.put((rank, totals["identities"], totals["generated"], files))
    finally:
        dist.destroy_process_group()


def test_two_process_gloo_sweep_writes_the_single_process_tree(tree, gold):
    """World-size-2 run of the sweep over real processes (gloo): every rank replays the plan, takes its identities, writes its
    own files; together they are exactly the files of the reference script's single-process run."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, str(tree), q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r, n_id, n_img) for r, n_id, n_img, _ in results] == [(0, 2, 126), (1, 2, 126)]
    gathered = results[0][3]
    assert results[1][3] == gathered                     # both ranks see the same gathered lists
    assert sorted(f for part in gathered for f in part) == sorted(s[0] for s in gold["saved"])
    assert all(os.path.isfile(f) for part in gathered for f in part)
