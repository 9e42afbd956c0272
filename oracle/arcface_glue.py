"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the glue between the decoded x0 image and the ArcFace backbone
(SURVEY 8a row a16): `latents_to_image_for_mtcnn` (/root/reference/train_ID-Booth.py:433-443), the bbox crop at its call
sites (`:1090,1123`) and `cropped_image_to_arcface_input` (`:445-455`, torchvision `resize(..., antialias=None)` on a
tensor = bilinear, align_corners=False, no antialiasing).
PINNED: checked against outputs of those reference functions themselves (tests/golden/arcface_glue_golden.pt, made by
tests/golden/make_arcface_glue_golden.py)."""
import torch
import torch.nn.functional as F


def decoded_to_mtcnn_image(image: torch.Tensor) -> torch.Tensor:
    """`vae.decode(z).sample` [1, 3, H, W] in about [-1, 1] -> [H, W, 3] in 0..255 (float, not rounded)."""
    return ((image / 2 + 0.5).clamp(0, 1) * 255)[0].permute(1, 2, 0)


def crop_to_arcface_input(img_hwc: torch.Tensor, bbox, size: int = 112) -> torch.Tensor:
    """[H, W, 3] in 0..255 + (x0, y0, x1, y1) -> [1, 3, size, size] in [-1, 1].  Both axes are clipped with the image HEIGHT,
    as the reference does (`initial_size = img.shape[0]`)."""
    lim = img_hwc.shape[0]
    x0, y0, x1, y1 = (int(v) for v in bbox)
    crop = img_hwc[max(0, y0):min(y1, lim), max(0, x0):min(x1, lim)]
    t = F.interpolate(crop.permute(2, 0, 1)[None], size=(size, size), mode="bilinear", align_corners=False, antialias=False)
    return ((t / 255) - 0.5) / 0.5
