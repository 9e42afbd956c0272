"""`faceposegenerator_b200.extract_embeds` host logic against a log of the reference's own `extract_ArcFace_embeds.py`
(tests/golden/extract_embeds_golden.pt, made by executing that script under stubs): same crop reaches the backbone, same
preprocessing, same files, same `files_without_faces.json`.  The backbone and the detector are injected; no GPU."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden_helpers():
    spec = importlib.util.spec_from_file_location("make_extract_embeds_golden",
                                                  os.path.join(ROOT, "tests", "golden", "make_extract_embeds_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)      # constants + the seeded image generator; the reference is only touched by main()
    return mod


class _Model:
    def __init__(self):
        self.inputs = []

    def to(self, *a, **k):
        return self

    def __call__(self, x):
        self.inputs.append(x.detach().clone())
        return x.mean(dim=(2, 3)).repeat(1, 171)[:, :512]


@pytest.fixture()
def dataset(tmp_path, monkeypatch):
    h = _golden_helpers()
    monkeypatch.chdir(tmp_path)
    for folder, files in h.TREE.items():
        os.makedirs(os.path.join("FACE_DATASET", "images", folder))
        for name, size, _ in files:
            Image.fromarray(h.seeded_image(folder, name, size)).save(os.path.join("FACE_DATASET", "images", folder, name))
    return h


def test_extraction_matches_reference_script_log(dataset):
    from faceposegenerator_b200.extract_embeds import run
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "extract_embeds_golden.pt"), weights_only=False)
    tree = dataset.TREE
    order = []

    def listdir(path):
        names = sorted(os.listdir(path))
        if os.path.basename(path) in tree:
            order.append(os.path.basename(path))
        return names

    def detector(images):   # the boxes the golden run's MTCNN stub returned
        return [None if b is None else np.array([b], dtype=np.float32) for _, _, b in tree[order[-1]]], None
    model = _Model()
    without = run("FACE_DATASET", device="cpu", model=model, detector=detector, listdir=listdir)
    assert without == gold["files_without_faces"]
    with open("FACE_DATASET/files_without_faces.json") as f:
        assert json.load(f) == gold["files_without_faces"]
    assert len(model.inputs) == len(gold["model_inputs"]) == 2          # one backbone call per folder: its last crop
    for got, want in zip(model.inputs, gold["model_inputs"]):
        got = got if want.shape[-1] == 112 else got[..., ::4, ::4]
        assert torch.equal(got, want)
    made = sorted(os.path.relpath(os.path.join(d, n)) for d, _, names in os.walk("FACE_DATASET") for n in names
                  if not n.endswith(".png"))
    assert made == gold["made_files"]
    for path, emb in gold["saved"]:
        assert torch.equal(torch.load(path), emb)


def test_embed_all_and_bbox_file(dataset):
    from faceposegenerator_b200.extract_embeds import run
    boxes = {os.path.join("images", folder, name): (None if b is None else list(b))
             for folder, files in dataset.TREE.items() for name, _, b in files}
    with open("boxes.json", "w") as f:
        json.dump(boxes, f)
    model = _Model()
    without = run("FACE_DATASET", device="cpu", model=model, bbox_file="boxes.json", embed="all",
                  listdir=lambda p: sorted(os.listdir(p)))
    assert without == {"files_without_faces": ["FACE_DATASET/images/idB/1.png"]}      # the image that really has no face
    assert [tuple(x.shape) for x in model.inputs] == [(2, 3, 112, 112), (2, 3, 112, 112)]
    assert torch.load("FACE_DATASET/ArcFace_embeds/idB/idB.pt").shape == (2, 512)


def test_missing_detector_is_an_error(dataset):
    from faceposegenerator_b200.extract_embeds import run
    with pytest.raises(RuntimeError, match="no face detector"):
        run("FACE_DATASET", device="cpu", model=_Model())
