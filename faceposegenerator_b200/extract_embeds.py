"""ArcFace ground-truth embedding extraction: the reference's `extract_ArcFace_embeds.py` with the
IResNet-100 backbone on the hand-written sm_100a kernels (`iresnet.IResNet`).  For every identity folder under
`<origin>/images/` it loads the images (`:40-46`), asks a face detector for one box per image (`:50`), crops
(`:58-62`, both axes clipped with the image HEIGHT like the reference), resizes to 112 x 112 with torchvision's `resize`
and normalises to [-1, 1] (`prepare_for_arcface_model_torch`, `:12-18`), embeds and saves `<origin>/ArcFace_embeds/
<folder>/<folder>.pt` (`:70-71`), then writes `<origin>/files_without_faces.json` (`:73-78`).

Reference behaviour kept bit for bit (pinned by tests/golden/extract_embeds_golden.pt, a log of the reference script itself):
the script embeds `img_cropped`, i.e. only the LAST crop of a folder (`:68`), and records the folder's LAST file name for
every image without a face (`:55`).  `embed="all"` returns / saves the embeddings of every detected face instead.

The detector is the caller's: MTCNN (`facenet_pytorch`, third party, not in the reference tree and not installed here) is
used when importable, otherwise pass `detector(images) -> (boxes, probs)` with `boxes[i]` = `None` or an array
`[[x0, y0, x1, y1], ...]`, or a `bbox_file` (JSON: relative image path -> [x0, y0, x1, y1] or null).  There is no
full-frame fallback: a missing detector is an error.
"""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch


def prepare_for_arcface(img: torch.Tensor) -> torch.Tensor:
    """[h, w, 3] (uint8, any device) -> [1, 3, 112, 112] float in [-1, 1]; torchvision's `resize` is the reference's own call."""
    from torchvision import transforms
    img = img.permute(2, 0, 1)
    img = transforms.functional.resize(img, (112, 112))
    img = img.float()
    img = ((img / 255) - 0.5) / 0.5
    return img[None, :, :, :]


def crop_to_bbox(image: torch.Tensor, bbox) -> torch.Tensor:
    x0, y0, x1, y1 = (int(v) for v in np.asarray(bbox).astype(int)[:4])
    lim = image.shape[0]
    return image[max(0, y0):min(y1, lim), max(0, x0):min(x1, lim)]


def load_arcface(weights: Optional[str] = "ArcFace_files/ArcFace_r100_ms1mv3_backbone.pth", device="cuda:0",
                 allow_random_weights: Optional[bool] = None):
    """`prepare_locked_ArcFace_model` (ArcFace_functions.py:27-37): r100, frozen, eval.  A missing checkpoint RAISES:
    "ground-truth" embeddings from a random backbone would be silently useless.  Deterministic random weights keyed by
    parameter name are an explicit opt-in for benchmarking / tests (`allow_random_weights=True` or
    IDB_ALLOW_RANDOM_WEIGHTS=1)."""
    from .iresnet import IResNet
    from .pipeline import random_weights_allowed
    from .weights import random_iresnet_state_dict
    if weights and os.path.isfile(weights):
        sd = torch.load(weights, map_location="cpu", weights_only=True)
    elif random_weights_allowed(allow_random_weights):
        import warnings
        warnings.warn(f"ArcFace checkpoint {weights!r} not found: RANDOM-INIT IResNet-100 (benchmarking / testing only; the "
                      "embeddings carry no identity information)", stacklevel=2)
        sd = random_iresnet_state_dict("r100", 0)
    else:
        raise FileNotFoundError(f"ArcFace checkpoint not found: {weights!r} (ArcFace_functions.py:29-30 loads "
                                "ArcFace_r100_ms1mv3_backbone.pth); pass its path, or opt in to random weights for benchmarking")
    return IResNet(sd, "r100", device)


def _json_detector(bbox_file: str, origin_path: str) -> Callable:
    with open(bbox_file) as f:
        table = json.load(f)

    def detect(images, paths):
        boxes = []
        for p in paths:
            b = table.get(os.path.relpath(p, origin_path), table.get(p))
            boxes.append(None if b is None else np.asarray([b], dtype=np.float32))
        return boxes, [None] * len(boxes)
    return detect


def embed_folder(model, images: Sequence[torch.Tensor], boxes: Sequence, embed: str = "last"):
    """-> (embeddings or None, indices of the images without a face)."""
    crops, missing = [], []
    for k, (image, bbox) in enumerate(zip(images, boxes)):
        if bbox is None:
            missing.append(k)
            continue
        crops.append(prepare_for_arcface(crop_to_bbox(image, bbox[0])))
    if not crops:
        return None, missing
    if embed == "last":
        return model(crops[-1]), missing
    if embed == "all":
        return model(torch.cat(crops, 0)), missing
    raise ValueError("embed must be 'last' (reference behaviour) or 'all'")


def run(origin_path: str = "FACE_DATASET", device: str = "cuda:0", model=None, detector: Optional[Callable] = None,
        bbox_file: Optional[str] = None, embed: str = "last", weights: Optional[str] = "ArcFace_files/ArcFace_r100_ms1mv3_backbone.pth",
        listdir: Callable = os.listdir) -> Dict[str, List[str]]:
    from PIL import Image
    if model is None:
        model = load_arcface(weights, device)
    model = model.to(device=device) or model
    takes_paths = False
    if detector is None and bbox_file is not None:
        detector, takes_paths = _json_detector(bbox_file, origin_path), True
    if detector is None:
        try:
            from facenet_pytorch import MTCNN
        except ImportError as e:
            raise RuntimeError("no face detector: install facenet_pytorch or pass detector= / bbox_file=") from e
        mtcnn = MTCNN(image_size=112, device=device, margin=0)
        detector = lambda images: mtcnn.detect(images, landmarks=False)   # noqa: E731
    without = {"files_without_faces": []}
    for folder in listdir(os.path.join(origin_path, "images")):
        folder_path = os.path.join(origin_path, "images", folder)
        output_path = folder_path.replace("images", "ArcFace_embeds")
        os.makedirs(output_path, exist_ok=True)
        paths = [os.path.join(folder_path, n) for n in listdir(folder_path)]
        images = [torch.from_numpy(np.array(Image.open(p))).to(device) for p in paths]
        stacked = torch.stack(images, dim=0)
        boxes, _ = detector(stacked, paths) if takes_paths else detector(stacked)
        emb, missing = embed_folder(model, images, boxes, embed)
        # the reference appends `img_path`, which after the loading loop is the folder's last file (`:55`)
        without["files_without_faces"] += [paths[-1] if embed == "last" else paths[k] for k in missing]
        if emb is not None:
            torch.save(emb, os.path.join(output_path, folder + ".pt"))
    with open(f"{origin_path}/files_without_faces.json", "w") as fp:
        json.dump(without, fp)
    return without


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="ArcFace embedding extraction (extract_ArcFace_embeds.py) on B200")
    ap.add_argument("--origin-path", default="FACE_DATASET")
    ap.add_argument("--bbox-file", default=None)
    ap.add_argument("--embed", default="last", choices=["last", "all"])
    ap.add_argument("--device", default="cuda:0")
    a = ap.parse_args()
    print(json.dumps(run(a.origin_path, a.device, bbox_file=a.bbox_file, embed=a.embed)))
