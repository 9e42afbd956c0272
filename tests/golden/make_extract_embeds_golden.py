"""Generates tests/golden/extract_embeds_golden.pt by EXECUTING the reference's own
/root/reference/extract_ArcFace_embeds.py (unmodified, via runpy) on a small seeded image tree in this container, with
stubs for what it cannot have here: `facenet_pytorch.MTCNN` (third party, absent: `.detect` returns preset boxes, one
image without a face), `Arcface_files.ArcFace_functions.prepare_locked_ArcFace_model` (a recording model -- the real
backbone is pinned separately by iresnet100_golden.pt), and `.to("cuda:0")` mapped to the CPU (no GPU here).  The log pins
what `faceposegenerator_b200.extract_embeds` must reproduce: which crop reaches the backbone for every identity folder (the
script embeds the LAST crop of a folder, `:68`), its preprocessing (`prepare_for_arcface_model_torch`, `:12-18`:
torchvision `resize` to 112x112, `(x/255 - 0.5)/0.5`), the saved file names and the `files_without_faces.json` content.
    python tests/golden/make_extract_embeds_golden.py
"""
import json
import os
import runpy
import sys
import tempfile
import types

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
SCRIPT = "/root/reference/extract_ArcFace_embeds.py"
# folder -> [(file name, (H = W) size, bbox x0 y0 x1 y1 or None)]; `os.listdir` order is made deterministic by sorting
TREE = {"idA": [("0.png", 160, (20.3, 31.9, 131.2, 140.7)), ("1.png", 160, (-4.0, 10.0, 170.0, 150.0))],
        "idB": [("0.png", 128, (30.0, 20.0, 100.0, 117.0)), ("1.png", 128, None), ("2.png", 128, (8.6, 9.4, 99.9, 120.1))]}
LOG = {"model_inputs": [], "saved": []}


def seeded_image(folder, name, size):
    seed = sum(map(ord, folder + name))
    rng = np.random.RandomState(seed)
    base = rng.randint(0, 256, size=(size // 8, size // 8, 3)).astype(np.uint8)     # blocky: resizing has structure to average
    return np.kron(base, np.ones((8, 8, 1), dtype=np.uint8))


class _Model:
    def to(self, *a, **k):
        return self

    def __call__(self, x):
        LOG["model_inputs"].append(x.detach().clone())
        return x.mean(dim=(2, 3)).repeat(1, 171)[:, :512]      # a stand-in "embedding" that depends on the input


class _MTCNN:
    def __init__(self, **kwargs):
        LOG["mtcnn_kwargs"] = {k: str(v) for k, v in kwargs.items()}

    def detect(self, images, landmarks=False):
        folder = _MTCNN.current
        boxes = [None if b is None else np.array([b], dtype=np.float32) for _, _, b in TREE[folder]]
        return np.array(boxes, dtype=object), [None] * len(boxes)


def main():
    fn_mod = types.ModuleType("Arcface_files.ArcFace_functions")
    fn_mod.prepare_locked_ArcFace_model = lambda: _Model()
    pkg = types.ModuleType("Arcface_files")
    pkg.ArcFace_functions = fn_mod
    facenet = types.ModuleType("facenet_pytorch")
    facenet.MTCNN = _MTCNN
    sys.modules.update({"Arcface_files": pkg, "Arcface_files.ArcFace_functions": fn_mod, "facenet_pytorch": facenet})

    real_to = torch.Tensor.to

    def to_cpu(self, *args, **kwargs):      # `.to("cuda:0")` / `.to(device="cuda:0")` -> stay on the CPU
        args = tuple("cpu" if isinstance(a, str) and a.startswith("cuda") else a for a in args)
        if isinstance(kwargs.get("device"), str) and kwargs["device"].startswith("cuda"):
            kwargs["device"] = "cpu"
        return real_to(self, *args, **kwargs)
    torch.Tensor.to = to_cpu

    real_listdir, real_save = os.listdir, torch.save

    def listdir(path):
        names = sorted(real_listdir(path))
        if os.path.basename(path) in TREE:
            _MTCNN.current = os.path.basename(path)
        return names
    os.listdir = listdir

    def save(obj, path, *a, **k):
        LOG["saved"].append((path, obj.detach().clone()))
        return real_save(obj, path, *a, **k)
    torch.save = save

    with tempfile.TemporaryDirectory() as cwd:
        os.chdir(cwd)
        for folder, files in TREE.items():
            os.makedirs(os.path.join("FACE_DATASET", "images", folder))
            for name, size, _ in files:
                Image.fromarray(seeded_image(folder, name, size)).save(os.path.join("FACE_DATASET", "images", folder, name))
        runpy.run_path(SCRIPT, run_name="__main__")
        with open("FACE_DATASET/files_without_faces.json") as f:
            without = json.load(f)
        made = sorted(os.path.relpath(os.path.join(d, n), cwd) for d, _, names in os.walk("FACE_DATASET") for n in names
                      if not n.endswith(".png"))
    torch.Tensor.to, os.listdir, torch.save = real_to, real_listdir, real_save
    LOG["model_inputs"][1:] = [t[..., ::4, ::4].clone() for t in LOG["model_inputs"][1:]]   # keep the fixture small
    gold = {"tree": TREE, "model_inputs": LOG["model_inputs"], "saved": LOG["saved"], "files_without_faces": without,
            "made_files": made, "mtcnn_kwargs": LOG["mtcnn_kwargs"]}
    path = os.path.join(HERE, "extract_embeds_golden.pt")
    real_save(gold, path)
    print(os.path.getsize(path), [tuple(t.shape) for t in gold["model_inputs"]], [p for p, _ in gold["saved"]], without, made)


if __name__ == "__main__":
    main()
