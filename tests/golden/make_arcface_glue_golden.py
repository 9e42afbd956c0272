"""Generates tests/golden/arcface_glue_golden.pt by running the reference's OWN glue functions between the decoded x0 image
and the ArcFace backbone -- `latents_to_image_for_mtcnn` and `cropped_image_to_arcface_input`
(/root/reference/train_ID-Booth.py:433-455), and its in-tree twin of the pipeline's image post-process,
`latents_to_pil_images` (`:408-417`), plus the bbox crop expression of its call sites (`:1090,1123`) -- in this
container.  `train_ID-Booth.py` cannot be imported (it imports diffusers / accelerate / facenet_pytorch at the top), so
the function definitions are taken from its syntax tree and executed as they are, with `torch` and
`torchvision.transforms` in scope; nothing of them is stored in this repo.  Inputs are regenerated from seeds.
    python tests/golden/make_arcface_glue_golden.py
"""
import ast
import os

import torch
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/train_ID-Booth.py"
WANTED = ("latents_to_image_for_mtcnn", "cropped_image_to_arcface_input", "latents_to_pil_images")
# (seed, image size, bbox x0 y0 x1 y1) -- the fixed cfg-5 box, a box hanging over two edges, a small off-centre box
CASES = [(0, 512, (96, 96, 416, 416)), (1, 512, (-20, 30, 540, 470)), (2, 256, (100, 90, 171, 200))]


def reference_functions():
    tree = ast.parse(open(SRC).read())
    from PIL import Image
    ns = {"torch": torch, "transforms": transforms, "Image": Image}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            exec(compile(ast.Module([node], []), SRC, "exec"), ns)
    return [ns[n] for n in WANTED]


def decoded_image(seed, size):
    """What `vae.decode(latents).sample` hands over: NCHW, roughly [-1.2, 1.2] so that the clamp is exercised."""
    return (torch.rand(1, 3, size, size, generator=torch.Generator().manual_seed(seed)) * 2.4 - 1.2)


class _Vae:   # `vae.decode(z).sample`: the decoder itself is pinned elsewhere; here it returns the seeded image
    def __init__(self, image):
        self.image = image

    def decode(self, z):
        return type("Out", (), {"sample": self.image})()


if __name__ == "__main__":
    to_mtcnn, to_arcface, to_pil = reference_functions()
    gold = {"cases": CASES, "mtcnn_image_slices": [], "arcface_inputs": []}
    # row a14: `latents_to_pil_images` (:408-417) = decode, /2 + 0.5, clamp, NHWC, * 255, round, uint8 -- two images, subsampled
    import numpy as np
    batch = torch.cat([decoded_image(10, 256), decoded_image(11, 256)])
    pils = to_pil(torch.zeros(2, 4, 32, 32), _Vae(batch))
    gold["pil_uint8_slices"] = torch.from_numpy(np.stack([np.asarray(im) for im in pils]))[:, ::4, ::4].clone()
    for seed, size, bbox in CASES:
        img = to_mtcnn(torch.zeros(1, 4, size // 8, size // 8), _Vae(decoded_image(seed, size)))     # [H, W, 3] in 0..255
        assert img.shape == (size, size, 3)
        initial_size = img.shape[0]
        img_cropped = img[max(0, bbox[1]): min(bbox[3], initial_size), max(0, bbox[0]): min(bbox[2], initial_size)]   # call site :1090
        x = to_arcface(img_cropped)
        assert x.shape == (1, 3, 112, 112)
        gold["mtcnn_image_slices"].append(img[::16, ::16].clone())
        gold["arcface_inputs"].append(x[0].clone() if seed == 0 else x[0, :, ::4, ::4].clone())
    path = os.path.join(HERE, "arcface_glue_golden.pt")
    torch.save(gold, path)
    print(os.path.getsize(path), [tuple(t.shape) for t in gold["arcface_inputs"]], float(gold["arcface_inputs"][0].abs().mean()))
