"""ctypes binding of libtimegan_b200.so (C ABI declared in include/timegan_b200.h).

There is NO fallback: if the shared library is missing the import fails loudly, and every compute entry
point needs a CUDA device (SURVEY.md section 8b "Must NOT exist: ... CPU fallback").
"""
import ctypes as C
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libtimegan_b200.so"

# flags (keep in sync with include/timegan_b200.h)
GRU_SAVE, GRU_NO_BULK, GRU_DY_LAST = 1, 2, 4
PROJ_FP32, PROJ_TF32, PROJ_TF32X3 = 0, 1, 2

if not LIB_PATH.exists():
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA extension first "
        f"(`python -c 'import __graft_entry__ as g; g.build()'` or eeg-gan-timegan-cgan_b200/csrc/build.sh). "
        "There is no CPU/PyTorch fallback for the TimeGAN hot path.")

lib = C.CDLL(str(LIB_PATH))

_vp, _i, _ll, _ull, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_ulonglong, C.c_float, C.c_size_t

# name -> (restype, argtypes).  Every symbol of include/timegan_b200.h must be listed (tests check this).
SIGNATURES = {
    "tg_version": (_i, []),
    "tg_last_error": (C.c_char_p, []),
    "tg_device_sm_count": (_i, []),
    "tg_set_option": (_i, [C.c_char_p, _i]),
    "tg_cluster_capacity": (_i, [_i, _i, _i]),
    "tg_launch_count": (_ll, []),
    "tg_prof_kinds": (_i, []),
    "tg_prof_kind_name": (C.c_char_p, [_i]),
    "tg_prof_enable": (None, [_i]),
    "tg_prof_reset": (None, []),
    "tg_prof_read": (_i, [_i, C.POINTER(C.c_double), C.POINTER(_ll), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "tg_proj": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i]),
    "tg_bf16_gi_supported": (_i, [_i, _i, _i, _i]),
    "tg_proj_bf16": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i]),
    "tg_dgrad": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i]),
    "tg_wgrad_workspace_bytes": (_sz, [_i, _i, _i]),
    "tg_wgrad": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _sz, _i]),
    "tg_wgrad_gru_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "tg_wgrad_gru": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _i]),
    "tg_gru_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "tg_gru_fwd_bf16gi": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "tg_gru_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "tg_gru_jvp_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "tg_gru_jvp_bwd": (_i, [_vp] * 14 + [_i, _i, _i, _i, _vp]),
    "tg_reduce_workspace_bytes": (_sz, []),
    "tg_sqdiff_sum": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _sz]),
    "tg_scaled_diff": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i]),
    "tg_diff1_sum": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _sz]),
    "tg_diff1_grad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "tg_center_scale": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i]),
    "tg_colsum_workspace_bytes": (_sz, [_i]),
    "tg_colsum": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _sz]),
    "tg_acf_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "tg_acf_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "tg_acf_bwd_final": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i]),
    "tg_acf_score": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "tg_sumsq_workspace_bytes": (_sz, [_i, C.POINTER(_ll)]),
    "tg_sumsq": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_ll), _vp, _vp, _sz]),
    "tg_adam": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_ll), _vp,
                     _f, _f, _f, _f, _f, _i, _f, _vp]),
    "tg_head_fwd": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tg_head_seed": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _f]),
    "tg_head_bwd": (_i, [_vp, _vp, _ll, _vp, _ll] + [_vp] * 13 + [_i, _i, _f, _f]),
    "tg_head_adv_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _f]),
    "tg_snapshot_if_better": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_ll), _vp, _vp, _vp, _f]),
    "tg_rng_uniform": (_i, [_vp, _vp, _ll, _ull, _ull, _f, _f, _vp]),
    "tg_rng_add_normal": (_i, [_vp, _vp, _vp, _ll, _f, _ull, _ull, _vp]),
    "tg_rng_add_normal_dev": (_i, [_vp, _vp, _vp, _ll, _vp, _ull, _ull, _vp]),
    "tg_peer_chunk_floats": (_i, []),
    "tg_peer_alloc": (_i, [C.POINTER(_vp), _sz]),
    "tg_peer_free": (_i, [_vp]),
    "tg_peer_export": (_i, [_vp, C.c_char_p]),
    "tg_peer_open": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "tg_peer_close": (_i, [_vp]),
    "tg_peer_site_bytes": (_sz, [_i, C.POINTER(_ll), _i, C.POINTER(_sz)]),
    "tg_peer_allreduce": (_i, [_vp, _i, _i, C.POINTER(_vp), _sz, _sz, _vp, _vp, _i, C.POINTER(_vp), C.POINTER(_ll)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = library older than the header: rebuild
    _fn.restype = _res
    _fn.argtypes = _args

ABI_VERSION = lib.tg_version()


def last_error() -> str:
    return (lib.tg_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str = ""):
    if rc != 0:
        kind = "argument error" if rc < 0 else "CUDA error"
        raise RuntimeError(f"libtimegan_b200 {what}: {kind} {rc}: {last_error()}")


def launch_count() -> int:
    return int(lib.tg_launch_count())


def prof_enable(on: bool):
    lib.tg_prof_enable(1 if on else 0)


def prof_reset():
    lib.tg_prof_reset()


def prof_read() -> dict:
    """{family: {"ms", "calls", "bytes", "flops"}} since the last prof_reset (synchronises the recorded events)."""
    out = {}
    for k in range(lib.tg_prof_kinds()):
        ms, calls, by, fl = C.c_double(), _ll(), C.c_double(), C.c_double()
        check(lib.tg_prof_read(k, C.byref(ms), C.byref(calls), C.byref(by), C.byref(fl)), "tg_prof_read")
        out[lib.tg_prof_kind_name(k).decode()] = dict(ms=ms.value, calls=calls.value, bytes=by.value, flops=fl.value)
    return out


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, name: str = "tensor") -> torch.Tensor:
    """The hot path has no CPU implementation: refuse anything that is not a CUDA fp32 tensor."""
    if not t.is_cuda:
        raise RuntimeError(
            f"timegan_b200: {name} is on {t.device}; the TimeGAN hot path only exists as sm_100a CUDA kernels "
            "(no CPU fallback). Move the model and data to a CUDA device.")
    if t.dtype != torch.float32:
        raise RuntimeError(f"timegan_b200: {name} must be float32, got {t.dtype}")
    return t


def ptr(t):
    return None if t is None else t.data_ptr()
