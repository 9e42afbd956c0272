"""ctypes binding of libidb_b200.so (C ABI declared in include/idb.h).

The library is the only compute path: if it is missing or the device is not sm_100
every op raises -- there is no eager / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libidb_b200.so")

A_1X1, A_3X3, A_3X3_S2, A_3X3_S2_ASYM, A_2X2 = 0, 1, 2, 3, 4
EPI_GEGLU = 1
EPI_F16 = 2
EPI_GELU = 4
EPI_PHASES4 = 8

c_void_p, c_int32, c_int64, c_float, c_size_t = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t


class GemmConvArgs(C.Structure):
    _fields_ = [
        ("a0", c_void_p), ("a0_mode", c_int32), ("c0", c_int32),
        ("a1", c_void_p), ("c1", c_int32),
        ("batch", c_int32), ("height", c_int32), ("width", c_int32),
        ("w", c_void_p), ("n", c_int32),
        ("bias", c_void_p), ("rowvec", c_void_p), ("rowvec_ld", c_int64), ("residual", c_void_p),
        ("lora_down", c_void_p), ("lora_up", c_void_p), ("lora_rank_pad", c_int32), ("lora_seg_n", c_int32),
        ("flags", c_int32),
        ("out_f32", c_void_p), ("out_bf16", c_void_p),
        ("k_splits", c_int32), ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("stats_partials", c_void_p),
        ("stats_image_sums", c_void_p), ("stats_hw", c_int32), ("stats_gran", c_int32),
        ("prelu", c_void_p),
        ("tap_off_x", c_int32), ("tap_off_y", c_int32),
        ("out_scale", c_int32), ("out_phase_x", c_int32), ("out_phase_y", c_int32),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("ld_q", c_int64), ("col0_q", c_int32),
        ("k", c_void_p), ("ld_k", c_int64), ("col0_k", c_int32),
        ("v", c_void_p), ("ld_v", c_int64), ("col0_v", c_int32),
        ("out", c_void_p), ("ld_out", c_int64),
        ("batch", c_int32), ("heads", c_int32), ("t_q", c_int32), ("t_kv", c_int32),
        ("scale", c_float), ("causal", c_int32),
        ("lse", c_void_p),
    ]


class AttentionBwdArgs(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("ld_q", c_int64), ("col0_q", c_int32),
        ("k", c_void_p), ("ld_k", c_int64), ("col0_k", c_int32),
        ("v", c_void_p), ("ld_v", c_int64), ("col0_v", c_int32),
        ("o", c_void_p), ("ld_o", c_int64), ("col0_o", c_int32),
        ("d_o", c_void_p), ("ld_do", c_int64), ("col0_do", c_int32),
        ("lse", c_void_p), ("dsum", c_void_p),
        ("dq", c_void_p), ("ld_dq", c_int64), ("col0_dq", c_int32),
        ("dk", c_void_p), ("ld_dk", c_int64), ("col0_dk", c_int32),
        ("dv", c_void_p), ("ld_dv", c_int64), ("col0_dv", c_int32),
        ("batch", c_int32), ("heads", c_int32), ("t_q", c_int32), ("t_kv", c_int32),
        ("scale", c_float),
    ]


class GroupNormBwdArgs(C.Structure):
    _fields_ = [
        ("dy", c_void_p),
        ("x0", c_void_p), ("c0", c_int32),
        ("x1", c_void_p), ("c1", c_int32),
        ("batch", c_int32), ("hw", c_int32), ("groups", c_int32), ("silu", c_int32),
        ("stats", c_void_p),
        ("gamma", c_void_p), ("beta", c_void_p),
        ("scratch", c_void_p),
        ("dx0", c_void_p), ("dx1", c_void_p),
        ("add0", c_int32), ("add1", c_int32),
    ]


class GroupNormArgs(C.Structure):
    _fields_ = [
        ("x0", c_void_p), ("c0", c_int32),
        ("x1", c_void_p), ("c1", c_int32),
        ("batch", c_int32), ("hw", c_int32), ("groups", c_int32),
        ("eps", c_float),
        ("gamma", c_void_p), ("beta", c_void_p),
        ("silu", c_int32),
        ("out_norm", c_void_p), ("out_raw", c_void_p),
        ("partials", c_void_p),
        ("x0_stats", c_void_p), ("x1_stats", c_void_p), ("x0_stats_phases", c_int32),
        ("x0_sums", c_void_p), ("x1_sums", c_void_p), ("sums_gran", c_int32), ("x1_batch", c_int32),
    ]


class TimeEmbedArgs(C.Structure):
    _fields_ = [
        ("timesteps", c_void_p),
        ("batch", c_int32), ("dim_sin", c_int32), ("dim_emb", c_int32),
        ("w1", c_void_p), ("b1", c_void_p),
        ("w2", c_void_p), ("b2", c_void_p),
        ("w_all", c_void_p), ("b_all", c_void_p), ("n_all", c_int32),
        ("proj_out", c_void_p),
        ("scratch", c_void_p),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "idb_version": (c_int32, []),
    "idb_last_error": (c_int32, [C.c_char_p, c_size_t]),
    "idb_device_check": (c_int32, []),
    "idb_num_sms": (c_int32, []),
    "idb_launch_count": (C.c_uint64, []),
    "idb_stream_k_launch_count": (C.c_uint64, []),
    "idb_stream_k_mode": (c_int32, []),
    "idb_gemm_conv": (c_int32, [C.POINTER(GemmConvArgs), c_void_p]),
    "idb_gemm_conv_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "idb_sizeof_args": (c_size_t, [c_int32]),
    "idb_attention": (c_int32, [C.POINTER(AttentionArgs), c_void_p]),
    "idb_groupnorm": (c_int32, [C.POINTER(GroupNormArgs), c_void_p]),
    "idb_attention_backward": (c_int32, [C.POINTER(AttentionBwdArgs), c_void_p]),
    "idb_layernorm_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_float, c_void_p]),
    "idb_groupnorm_backward": (c_int32, [C.POINTER(GroupNormBwdArgs), c_void_p]),
    "idb_geglu_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "idb_lora_wgrad_workspace_bytes": (c_size_t, [c_int32]),
    "idb_lora_wgrad": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_float, c_int32,
                                 c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "idb_zero_insert2x": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "idb_sumpool2x": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "idb_groupnorm_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "idb_layernorm": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p]),
    "idb_softmax_rows": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p]),
    "idb_time_embed": (c_int32, [C.POINTER(TimeEmbedArgs), c_void_p]),
    "idb_conv3x3_small_cin": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "idb_conv3x3_small_cout": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                         c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "idb_upsample2x": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "idb_cast_bf16": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "idb_vae_latent_prep": (c_int32, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int32, c_int32, c_void_p]),
    "idb_channel_affine": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                     c_int32, c_void_p]),
    "idb_crop_resize_norm": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                       c_void_p]),
    "idb_latent_operand": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "idb_cfg_ddpm_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int32, c_int32,
                                    c_void_p, c_void_p, c_int64, c_void_p]),
}

_lib: Optional[C.CDLL] = None
trace = None      # profiling only: set to a list to record (entry point, description) per call
launch_count = 0  # kernels launched by the library so far (exact: the library counts every launch; bench.py's gpu_launches)   # entry points that always launch more than one kernel


def load() -> C.CDLL:
    """dlopen the in-tree library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m faceposegenerator_b200.csrc.build` "
                "(this package has no CPU / eager fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        for which, struct in enumerate((GemmConvArgs, AttentionArgs, GroupNormArgs, TimeEmbedArgs, AttentionBwdArgs, GroupNormBwdArgs)):
            if lib.idb_sizeof_args(which) != C.sizeof(struct):   # a stale binding would make the library read past the struct
                raise RuntimeError(f"{struct.__name__}: ctypes layout ({C.sizeof(struct)} B) does not match include/idb.h "
                                   f"({lib.idb_sizeof_args(which)} B); rebuild the library / update _lib.py")
        _lib = lib
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    load().idb_last_error(buf, 1024)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def call(name: str, *args, desc=None, launches: int = 0) -> None:
    """Calls an entry point; `launch_count` follows the library's own exact kernel-launch counter (`idb_launch_count`).
    (`launches` is accepted for callers that know the count; the library's counter is authoritative.)"""
    global launch_count
    if trace is not None:
        trace.append((name, desc))
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {last_error()}")
    launch_count = int(lib.idb_launch_count())
