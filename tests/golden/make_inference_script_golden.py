"""Generates tests/golden/inference_script_golden.json by EXECUTING the reference's own driver script
(/root/reference/inference_ID-Booth.py, unmodified, via runpy) in this container, with recording stubs in place of the
packages it calls into (diffusers / accelerate are not installed here and there is no GPU): every
`StableDiffusionPipeline.from_pretrained / load_lora_weights / __call__` and every `save_image` the script issues is
logged together with the seed of the `torch.Generator` it was given.  The log pins the caller side of the hot path --
identity order (natural sort, `.json` entries dropped), the python-RNG prompt schedule (`random.sample` per identity,
`random.choice` per prompt for the pose), LoRA paths, per-identity generator seeds, pipeline kwargs and output file
names -- which `faceposegenerator_b200.sweep.plan` must reproduce (tests/test_sweep_cpu.py).
    python tests/golden/make_inference_script_golden.py
"""
import json
import os
import random
import runpy
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
IDS = ["1", "2", "10", "3_b"]                      # exercises the natural sort ("10" after "2")
GENDERS = {"1": "M", "2": "F", "10": "F", "3_b": "M"}
MODELS = ["DreamBooth", "PortraitBooth", "ID-Booth"]
CHECKPOINT = "checkpoint-31-6400"

LOG = {"pipelines": [], "calls": [], "saved": []}


class _Generator:   # stands in for torch.Generator(device="cuda:0") on a CPU-only torch
    def __init__(self, device=None):
        self.device, self.seed = str(device), None

    def manual_seed(self, seed):
        self.seed = int(seed)
        return self


class _Output:
    def __init__(self):
        self.images = np.zeros((1, 8, 8, 3), dtype=np.float32)


class _Pipeline:
    def __init__(self, arch, kwargs):
        self.rec = {"from_pretrained": arch, "torch_dtype": str(kwargs.get("torch_dtype")), "to": None, "scheduler": None,
                    "scheduler_args": None, "lora": None, "progress_bar": None}
        LOG["pipelines"].append(self.rec)

    @classmethod
    def from_pretrained(cls, arch, **kwargs):
        return cls(arch, kwargs)

    def to(self, device):
        self.rec["to"] = str(device)
        return self

    def __setattr__(self, name, value):
        if name == "scheduler":
            self.rec["scheduler"] = type(value).__name__
            self.rec["scheduler_args"] = getattr(value, "args", None)
        object.__setattr__(self, name, value)

    def load_lora_weights(self, path, **kwargs):
        self.rec["lora"] = path

    def set_progress_bar_config(self, **kwargs):
        self.rec["progress_bar"] = kwargs

    def __call__(self, **kwargs):
        gen = kwargs.pop("generator")
        LOG["calls"].append({"pipeline": len(LOG["pipelines"]) - 1, "generator_seed": gen.seed, "generator_device": gen.device,
                             **kwargs})
        return _Output()


class _Scheduler:
    @classmethod
    def from_pretrained(cls, arch, **kwargs):
        s = cls()
        s.args = [arch, kwargs]
        return s


class DDPMScheduler(_Scheduler):
    pass


class DPMSolverMultistepScheduler(_Scheduler):
    pass


def _set_seed(seed, *a, **k):   # accelerate.utils.set_seed: python, numpy and torch generators
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def _save_image(tensor, fp, **kwargs):
    LOG["saved"].append({"fp": fp, "shape": list(tensor.shape), **kwargs})


def main():
    diffusers = types.ModuleType("diffusers")
    diffusers.StableDiffusionPipeline = _Pipeline
    diffusers.AutoPipelineForText2Image = _Pipeline
    diffusers.DDPMScheduler = DDPMScheduler
    diffusers.DPMSolverMultistepScheduler = DPMSolverMultistepScheduler
    accelerate = types.ModuleType("accelerate")
    accelerate_utils = types.ModuleType("accelerate.utils")
    accelerate_utils.set_seed = _set_seed
    accelerate.utils = accelerate_utils
    sys.modules.update({"diffusers": diffusers, "accelerate": accelerate, "accelerate.utils": accelerate_utils})
    import torchvision.utils
    torchvision.utils.save_image = _save_image
    torch.Generator = _Generator
    sys.path.insert(0, REFERENCE)   # `from utils.sorting_utils import natural_keys`

    with tempfile.TemporaryDirectory() as cwd:
        os.chdir(cwd)
        for m in MODELS:
            for i in IDS:
                os.makedirs(os.path.join("Trained_LoRA_Models", m, i, CHECKPOINT))
            with open(os.path.join("Trained_LoRA_Models", m, "training_args.json"), "w") as f:
                f.write("{}")
        with open("tufts_gender_dict.json", "w") as f:
            json.dump(GENDERS, f)
        ns = runpy.run_path(os.path.join(REFERENCE, "inference_ID-Booth.py"), run_name="__main__")
        made_dirs = sorted(os.path.relpath(os.path.join(d, s), cwd) for d, subs, _ in os.walk("Generated_Samples") for s in subs)

    gold = {"how": "runpy of /root/reference/inference_ID-Booth.py under recording stubs (tests/golden/make_inference_script_golden.py)",
            "ids_on_disk": IDS, "genders": GENDERS, "models": MODELS, "checkpoint": CHECKPOINT,
            "script_constants": {k: ns[k] for k in ("num_samples_per_prompt", "num_prompts", "add_gender", "add_pose", "add_age",
                                                    "add_background", "seed", "guidance_scale", "num_inference_steps",
                                                    "folder_of_models", "folder_output", "model_architecture", "width", "height",
                                                    "negative_prompt", "original_prompt", "all_prompt_combinations", "ids")},
            "pipelines": LOG["pipelines"], "made_dirs": made_dirs}
    # compact form: the kwargs every call shares are stored once (and checked to be shared)
    const = {k: v for k, v in LOG["calls"][0].items() if k not in ("pipeline", "generator_seed", "prompt")}
    assert all({k: c[k] for k in const} == const and set(c) == set(const) | {"pipeline", "generator_seed", "prompt"} for c in LOG["calls"])
    gold["call_constant_kwargs"] = const
    gold["calls"] = [[c["pipeline"], c["generator_seed"], c["prompt"]] for c in LOG["calls"]]
    gold["saved"] = [[s_["fp"], s_["shape"], {k: v for k, v in s_.items() if k not in ("fp", "shape")}] for s_ in LOG["saved"]]
    path = os.path.join(HERE, "inference_script_golden.json")
    with open(path, "w") as f:
        json.dump(gold, f, indent=0)
    print(len(LOG["pipelines"]), "pipelines,", len(LOG["calls"]), "calls,", len(LOG["saved"]), "saved;", os.path.getsize(path), "bytes")
    print(LOG["calls"][0])
    print(LOG["calls"][-1]["prompt"], "|", LOG["saved"][-1]["fp"])


if __name__ == "__main__":
    main()
