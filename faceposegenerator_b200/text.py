"""Prompt -> [n, 77, 1024] context for the UNet.  PLUMBING, not part of the accelerated path:
SURVEY.md section 8(f)-1 lists the CLIP-H text encoder as the next row; until it is moved onto
the sm_100a kernels this module runs the 23-layer OpenCLIP-H text tower (A.0 config) with plain
torch ops on the GPU, random-initialised deterministically when no snapshot exists offline, and
tokenises with a hashed word-piece stand-in because the CLIP BPE vocabulary is not available
without network access.  Results are cached per prompt string (the negative prompt of the
reference sweep is constant, `inference_ID-Booth.py:81`).
"""
from __future__ import annotations

import hashlib
import re
from typing import Dict, List

import torch
import torch.nn.functional as F

TEXT_CONFIG = dict(hidden=1024, intermediate=4096, heads=16, layers=23, max_pos=77, vocab=49408, eps=1e-5)
BOS, EOS = 49406, 49407


class HashTokenizer:
    """Deterministic stand-in for the CLIP BPE tokenizer (vocab files are not shipped offline):
    lower-cased word / punctuation pieces hashed into [0, 49405], BOS ... EOS, padded with EOS
    to 77 like `tokenizer(..., padding="max_length", max_length=77)` (`train_ID-Booth.py:463-469`)."""
    model_max_length = 77

    def __call__(self, texts: List[str]) -> torch.Tensor:
        rows = []
        for t in texts:
            pieces = re.findall(r"[a-z0-9]+|[^\sa-z0-9]", t.lower())
            ids = [int.from_bytes(hashlib.sha256(p.encode()).digest()[:4], "little") % 49406 for p in pieces]
            ids = [BOS] + ids[: self.model_max_length - 2] + [EOS]
            ids += [EOS] * (self.model_max_length - len(ids))
            rows.append(ids)
        return torch.tensor(rows, dtype=torch.long)


def text_manifest(cfg: dict = TEXT_CONFIG):
    h, inter = cfg["hidden"], cfg["intermediate"]
    m = [("text_model.embeddings.token_embedding.weight", (cfg["vocab"], h)),
         ("text_model.embeddings.position_embedding.weight", (cfg["max_pos"], h))]
    for i in range(cfg["layers"]):
        p = f"text_model.encoder.layers.{i}"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            m += [(f"{p}.self_attn.{n}.weight", (h, h)), (f"{p}.self_attn.{n}.bias", (h,))]
        m += [(f"{p}.layer_norm1.weight", (h,)), (f"{p}.layer_norm1.bias", (h,)),
              (f"{p}.mlp.fc1.weight", (inter, h)), (f"{p}.mlp.fc1.bias", (inter,)),
              (f"{p}.mlp.fc2.weight", (h, inter)), (f"{p}.mlp.fc2.bias", (h,)),
              (f"{p}.layer_norm2.weight", (h,)), (f"{p}.layer_norm2.bias", (h,))]
    m += [("text_model.final_layer_norm.weight", (h,)), ("text_model.final_layer_norm.bias", (h,))]
    return m


class CLIPTextEncoder:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device, dtype=torch.bfloat16, cfg: dict = TEXT_CONFIG):
        self.cfg, self.device, self.dtype = cfg, torch.device(device), dtype
        self.sd = {k: v.to(device=self.device, dtype=dtype) for k, v in state_dict.items()}
        self.tokenizer = HashTokenizer()
        self._cache: Dict[str, torch.Tensor] = {}

    @torch.no_grad()
    def forward_ids(self, ids: torch.Tensor) -> torch.Tensor:
        sd, cfg = self.sd, self.cfg
        ids = ids.to(self.device)
        n, S = ids.shape
        x = sd["text_model.embeddings.token_embedding.weight"][ids] + \
            sd["text_model.embeddings.position_embedding.weight"][:S][None]
        heads, h = cfg["heads"], cfg["hidden"]
        for i in range(cfg["layers"]):
            p = f"text_model.encoder.layers.{i}"
            r = x
            y = F.layer_norm(x, (h,), sd[p + ".layer_norm1.weight"], sd[p + ".layer_norm1.bias"], cfg["eps"])
            q, k, v = (F.linear(y, sd[f"{p}.self_attn.{n_}.weight"], sd[f"{p}.self_attn.{n_}.bias"])
                       .view(n, S, heads, h // heads).transpose(1, 2) for n_ in ("q_proj", "k_proj", "v_proj"))
            a = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(n, S, h)
            x = r + F.linear(a, sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"])
            r = x
            y = F.layer_norm(x, (h,), sd[p + ".layer_norm2.weight"], sd[p + ".layer_norm2.bias"], cfg["eps"])
            y = F.gelu(F.linear(y, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"]))
            x = r + F.linear(y, sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
        return F.layer_norm(x, (h,), sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"],
                            cfg["eps"])

    def encode(self, prompts: List[str]) -> torch.Tensor:
        missing = [p for p in dict.fromkeys(prompts) if p not in self._cache]
        if missing:
            out = self.forward_ids(self.tokenizer(missing))
            for p, e in zip(missing, out):
                self._cache[p] = e
        return torch.stack([self._cache[p] for p in prompts])
