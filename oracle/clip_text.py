"""TEST INFRASTRUCTURE ONLY -- fp32 restatement of the CLIP text tower of SD2.1 (OpenCLIP ViT-H text encoder as exposed by
transformers `CLIPTextModel`: pre-LN blocks, causal self-attention, erf-GELU MLP, final LayerNorm; SURVEY App. A.0;
reference call site `encode_prompt` behind `inference_ID-Booth.py:138`, in-tree twin `train_ID-Booth.py:457-491`).
PINNED: checked against outputs of transformers' own `CLIPTextModel` (5.5.0 installed here; the reference pins 4.34.1) with
the SD2.1-base text config on the repo's keyed random-init weights (tests/golden/clip_text_golden.pt, made by
tests/golden/make_clip_text_golden.py; the SD2.1 text weights themselves are not available offline)."""
import torch
import torch.nn.functional as F

TEXT_CONFIG = dict(hidden=1024, intermediate=4096, heads=16, layers=23, max_pos=77, vocab=49408, eps=1e-5)


@torch.no_grad()
def clip_text_forward(sd, ids, cfg=TEXT_CONFIG):
    """sd: state dict (any device, fp32); ids: [n, S] long -> [n, S, hidden] fp32."""
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
    return F.layer_norm(x, (h,), sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"], cfg["eps"])
