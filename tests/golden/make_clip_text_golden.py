"""Generates tests/golden/clip_text_golden.pt by running transformers' own `CLIPTextModel` (the class behind
`pipe.text_encoder` / `encode_prompt` at `inference_ID-Booth.py:138`, in-tree twin `train_ID-Booth.py:457-491`) in this
container: SD2.1-base text config (hidden 1024, 23 layers, 16 heads, intermediate 4096, 77 positions,
vocab 49408, gelu), weights regenerated deterministically from key names (`weights.random_state_dict(text_manifest())`),
token ids from the repo's hashed tokenizer, the encoder driven through the reference's OWN in-tree `encode_prompt`
(`train_ID-Booth.py:476-491`).  Only outputs are stored.  The installed transformers is 5.5.0 (the
reference pins 4.34.1, `requirements.txt`): same module, same state-dict keys for everything the manifest names.
    python tests/golden/make_clip_text_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import transformers  # noqa: E402
from transformers import CLIPTextConfig, CLIPTextModel  # noqa: E402

from faceposegenerator_b200.text import TEXT_CONFIG, HashTokenizer, text_manifest  # noqa: E402
from faceposegenerator_b200.weights import random_state_dict  # noqa: E402

PROMPTS = ["face portrait photo of a 34 y.o. woman, neutral expression, studio lighting",
           "blurry, cartoon, low resolution"]


def build_model():
    c = TEXT_CONFIG
    cfg = CLIPTextConfig(vocab_size=c["vocab"], hidden_size=c["hidden"], intermediate_size=c["intermediate"],
                         num_hidden_layers=c["layers"], num_attention_heads=c["heads"],
                         max_position_embeddings=c["max_pos"], hidden_act="gelu", layer_norm_eps=c["eps"],
                         projection_dim=512, pad_token_id=1, bos_token_id=0, eos_token_id=2)
    model = CLIPTextModel(cfg).eval()
    sd = random_state_dict(text_manifest(), seed=0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    # everything the manifest names must land in the transformers module; the only keys it may not cover are buffers
    assert not unexpected, unexpected
    assert all("position_ids" in k for k in missing), missing
    return model, sd


if __name__ == "__main__":
    torch.manual_seed(0)
    model, sd = build_model()
    ids = HashTokenizer()(PROMPTS)
    with torch.no_grad():
        out = model(input_ids=ids, output_hidden_states=True)
        # the reference's own in-tree `encode_prompt` (train_ID-Booth.py:476-491, taken from its syntax tree because the
        # module imports diffusers at the top) driving the same text encoder: this is the tensor the UNet receives
        import ast
        src = "/root/reference/train_ID-Booth.py"
        ns = {"torch": torch}
        for node in ast.parse(open(src).read()).body:
            if isinstance(node, ast.FunctionDef) and node.name == "encode_prompt":
                exec(compile(ast.Module([node], []), src, "exec"), ns)
        via_reference = ns["encode_prompt"](model, ids, None)
        assert torch.equal(via_reference, out.last_hidden_state)
    hs = out.hidden_states   # embeddings + one per layer (before the final LayerNorm)
    gold = {"transformers_version": transformers.__version__, "ids": ids, "prompts": PROMPTS,
            "last_hidden_state": via_reference.clone(),   # == out.last_hidden_state (asserted above)
            "hidden_1_slice": hs[1][:, :, :16].clone(), "hidden_12_slice": hs[12][:, :, :16].clone(),
            "hidden_23_slice": hs[23][:, :, :16].clone(),
            "n_params": sum(p.numel() for p in model.parameters())}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "clip_text_golden.pt")
    torch.save(gold, path)
    print({k: (tuple(v.shape) if hasattr(v, "shape") else v) for k, v in gold.items()},
          float(out.last_hidden_state.abs().mean()), os.path.getsize(path))
