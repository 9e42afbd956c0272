"""Prompt -> [n, 77, 1024] context for the UNet: the 23-layer OpenCLIP-H text tower
(A.0 config; `encode_prompt` behind `inference_ID-Booth.py:138`, in-tree twin `train_ID-Booth.py:457-491`) on the
sm_100a kernels: LayerNorm -> fused q/k/v GEMM -> causal attention (`idb_attention`, short-context kernel) ->
out_proj GEMM (+residual) -> LayerNorm -> fc1 GEMM with an erf-GELU epilogue -> fc2 GEMM (+residual); fp32 residual
stream, bf16 operands.  Only the embedding gather is a torch indexing op.  Weights are random-initialised
deterministically when no snapshot exists offline, and the tokenizer is a hashed word-piece stand-in because the CLIP
BPE vocabulary is not available without network access.  Results are cached per prompt string (the negative prompt
of the reference sweep is constant, `inference_ID-Booth.py:81`).  The plain-torch fp32 formulation the parity test
compares against lives in `oracle/clip_text.py`.
"""
from __future__ import annotations

import hashlib
import re
from typing import Dict, List

import torch

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


class SnapshotTokenizer:
    """The snapshot's own CLIP BPE tokenizer (`<dir>/tokenizer/{vocab.json, merges.txt, ...}`) through transformers'
    `CLIPTokenizer`, called the way diffusers' `encode_prompt` / `train_ID-Booth.py:463-469` call it:
    `padding="max_length"`, `max_length=model_max_length` (77), `truncation=True`."""

    def __init__(self, tokenizer_dir: str):
        from transformers import CLIPTokenizer
        self.tok = CLIPTokenizer.from_pretrained(tokenizer_dir, local_files_only=True)
        self.model_max_length = min(int(self.tok.model_max_length), TEXT_CONFIG["max_pos"])

    def __call__(self, texts: List[str]) -> torch.Tensor:
        out = self.tok(list(texts), padding="max_length", max_length=self.model_max_length, truncation=True,
                       return_tensors="pt")
        return out.input_ids.to(torch.long)


def load_tokenizer(snapshot_dir=None, allow_hash: bool = True):
    """`SnapshotTokenizer` when a local model snapshot ships its tokenizer files, else the hashed stand-in (there is no
    CLIP vocabulary offline, and with random-init weights the token ids only need to be deterministic).  With real
    weights (`allow_hash=False`) a missing tokenizer is an error: hashed ids would silently produce wrong embeddings."""
    if snapshot_dir:
        import os
        tok_dir = os.path.join(str(snapshot_dir), "tokenizer")
        if os.path.isfile(os.path.join(tok_dir, "vocab.json")) and os.path.isfile(os.path.join(tok_dir, "merges.txt")):
            return SnapshotTokenizer(tok_dir)
    if not allow_hash:
        raise FileNotFoundError(f"no tokenizer/vocab.json + merges.txt under {snapshot_dir!r}: the CLIP BPE vocabulary is "
                                "required with real weights (the hashed stand-in is for random-init benchmarking only)")
    return HashTokenizer()


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
    def __init__(self, state_dict: Dict[str, torch.Tensor], device, dtype=torch.bfloat16, cfg: dict = TEXT_CONFIG,
                 tokenizer=None):
        self.cfg, self.device, self.dtype = cfg, torch.device(device), dtype
        if self.device.type != "cuda":
            raise RuntimeError("CLIPTextEncoder runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.tokenizer = tokenizer if tokenizer is not None else HashTokenizer()
        self._cache: Dict[str, torch.Tensor] = {}
        bf, f32 = torch.bfloat16, torch.float32
        dev = self.device
        g = lambda k, dt: state_dict[k].to(device=dev, dtype=dt).contiguous()
        self.tok = g("text_model.embeddings.token_embedding.weight", f32)
        self.pos = g("text_model.embeddings.position_embedding.weight", f32)
        self.layers = []
        for i in range(cfg["layers"]):
            p = f"text_model.encoder.layers.{i}"
            cat = lambda suf, dt: torch.cat([state_dict[f"{p}.self_attn.{n_}.{suf}"] for n_ in ("q_proj", "k_proj", "v_proj")], 0) \
                .to(device=dev, dtype=dt).contiguous()
            self.layers.append(dict(
                ln1=(g(p + ".layer_norm1.weight", f32), g(p + ".layer_norm1.bias", f32)),
                w_qkv=cat("weight", bf), b_qkv=cat("bias", f32),
                w_o=g(p + ".self_attn.out_proj.weight", bf), b_o=g(p + ".self_attn.out_proj.bias", f32),
                ln2=(g(p + ".layer_norm2.weight", f32), g(p + ".layer_norm2.bias", f32)),
                w_fc1=g(p + ".mlp.fc1.weight", bf), b_fc1=g(p + ".mlp.fc1.bias", f32),
                w_fc2=g(p + ".mlp.fc2.weight", bf), b_fc2=g(p + ".mlp.fc2.bias", f32)))
        self.ln_f = (g("text_model.final_layer_norm.weight", f32), g("text_model.final_layer_norm.bias", f32))

    @torch.no_grad()
    def forward_ids(self, ids: torch.Tensor) -> torch.Tensor:
        """ids [n, S<=77] -> bf16 [n, S, hidden] on the hand-written kernels."""
        from . import ops
        cfg = self.cfg
        ids = ids.to(self.device)
        n, S = ids.shape
        h, heads = cfg["hidden"], cfg["heads"]
        x = (self.tok[ids] + self.pos[:S][None]).reshape(n * S, h).contiguous()       # fp32 residual stream
        for L in self.layers:
            y = ops.layernorm(x, *L["ln1"], eps=cfg["eps"])
            _, qkv = ops.gemm_conv(y, L["w_qkv"], bias=L["b_qkv"], want_bf16=True)
            a = ops.attention(qkv, qkv, qkv, batch=n, heads=heads, t_q=S, t_kv=S, scale=(h // heads) ** -0.5,
                              col0_q=0, col0_k=h, col0_v=2 * h, causal=True)
            x, _ = ops.gemm_conv(a, L["w_o"], bias=L["b_o"], residual=x, want_f32=True)
            y = ops.layernorm(x, *L["ln2"], eps=cfg["eps"])
            _, y = ops.gemm_conv(y, L["w_fc1"], bias=L["b_fc1"], gelu=True, want_bf16=True)
            x, _ = ops.gemm_conv(y, L["w_fc2"], bias=L["b_fc2"], residual=x, want_f32=True)
        return ops.layernorm(x, *self.ln_f, eps=cfg["eps"]).view(n, S, h)

    def _forward_ids_graphed(self, ids: torch.Tensor) -> torch.Tensor:
        """`forward_ids` as one CUDA-graph launch per batch size (164 kernels of a few microseconds each: eager, the host
        launch rate is the bottleneck).  IDB_CUDA_GRAPH=0: eager."""
        import os
        if os.environ.get("IDB_CUDA_GRAPH", "1") == "0" or torch.cuda.is_current_stream_capturing():
            return self.forward_ids(ids)
        ids = ids.to(self.device)
        key = tuple(ids.shape)
        graphs = self.__dict__.setdefault("_graphs", {})
        st = graphs.get(key)
        if st is None:
            if len(graphs) >= 8:
                graphs.pop(next(iter(graphs)))
            from . import _lib
            buf = ids.clone()
            with torch.cuda.device(self.device):
                self.forward_ids(buf)                       # warm-up outside capture (lazy kernel attributes)
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                n0 = _lib.launch_count
                with torch.cuda.graph(g):
                    out = self.forward_ids(buf)
            st = graphs[key] = (g, buf, out, _lib.launch_count - n0)
        g, buf, out, launches = st
        buf.copy_(ids)
        g.replay()
        from . import pipeline as _pl
        _pl.graph_launches += launches                      # (replays do not pass through the library's launch counter)
        return out.clone()

    def encode(self, prompts: List[str]) -> torch.Tensor:
        missing = [p for p in dict.fromkeys(prompts) if p not in self._cache]
        if missing:
            out = self._forward_ids_graphed(self.tokenizer(missing))
            for p, e in zip(missing, out):
                self._cache[p] = e
        return torch.stack([self._cache[p] for p in prompts])
