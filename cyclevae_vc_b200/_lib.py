"""ctypes binding of libcyclevae_b200.so (include/cyclevae_b200.h).

The CUDA library is the only execution path: importing this module without the built
library raises, and every call that fails raises RuntimeError(cvb_last_error()).
There is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcyclevae_b200.so")
ABI_VERSION = 13

c_float_p = C.POINTER(C.c_float)


class CvbNet(C.Structure):
    _fields_ = [
        ("in_dim", C.c_int32), ("out_dim", C.c_int32), ("hidden", C.c_int32), ("kernel_size", C.c_int32),
        ("n_conv", C.c_int32), ("has_scale_in", C.c_int32), ("has_scale_out", C.c_int32), ("reserved", C.c_int32),
        ("scale_in_w", C.c_void_p), ("scale_in_b", C.c_void_p),
        ("conv_w", C.c_void_p * 4), ("conv_b", C.c_void_p * 4),
        ("w_ih", C.c_void_p), ("w_hh", C.c_void_p), ("b_ih", C.c_void_p), ("b_hh", C.c_void_p),
        ("out_w", C.c_void_p), ("out_b", C.c_void_p), ("scale_out_w", C.c_void_p), ("scale_out_b", C.c_void_p),
    ]


class CvbNetGrads(C.Structure):
    _fields_ = [
        ("scale_in_w", C.c_void_p), ("scale_in_b", C.c_void_p),
        ("conv_w", C.c_void_p * 4), ("conv_b", C.c_void_p * 4),
        ("w_ih", C.c_void_p), ("w_hh", C.c_void_p), ("b_ih", C.c_void_p), ("b_hh", C.c_void_p),
        ("out_w", C.c_void_p), ("out_b", C.c_void_p), ("scale_out_w", C.c_void_p), ("scale_out_b", C.c_void_p),
        ("accumulate", C.c_int32), ("reserved", C.c_int32),
    ]


HEAD_NONE, HEAD_CLAMP, HEAD_SCALE_OUT = 0, 1, 2

_vp, _i, _sz, _f, _u64 = C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_uint64
_netp, _gradp = C.POINTER(CvbNet), C.POINTER(CvbNetGrads)

# name -> (restype, argtypes); every symbol declared in include/cyclevae_b200.h
PROTOTYPES = {
    "cvb_last_error": (C.c_char_p, []),
    "cvb_abi_version": (_i, []),
    "cvb_device_info": (_i, [C.POINTER(_i)] * 4),
    "cvb_frontend_ws_floats": (_sz, [_netp, _i, _i]),
    "cvb_recurrent_ws_floats": (_sz, [_netp, _i, _i, _i, _i]),
    "cvb_scratch_floats": (_sz, [_netp, _i, _i, _i]),
    "cvb_recurrence_max_rows": (_i, [_netp, _i]),
    "cvb_last_recurrence_path": (_i, [_i]),
    "cvb_last_recurrence_hops": (_i, [_i]),
    "cvb_gru_rnn_forward": (_i, [_netp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cvb_gru_rnn_backward": (_i, [_netp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _gradp, _vp]),
    "cvb_frontend_fwd": (_i, [_netp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "cvb_frontend_bwd_ws_floats": (_sz, [_netp, _i, _i]),
    "cvb_frontend_bwd": (_i, [_netp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _gradp, _vp]),
    "cvb_reparam_concat_fwd": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp]),
    "cvb_reparam_concat_bwd": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "cvb_comm_unique_id": (_i, [_vp]),
    "cvb_comm_init": (_i, [C.POINTER(_vp), _vp, _i, _i]),
    "cvb_allreduce_sum": (_i, [_vp, _vp, _sz, _vp]),
    "cvb_comm_destroy": (_i, [_vp]),
    "cvb_concat2_fwd": (_i, [_i, _i, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "cvb_kl_fwd": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "cvb_kl_bwd": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "cvb_mcd_l1_fwd": (_i, [_i, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "cvb_mcd_l1_bwd": (_i, [_i, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "cvb_dtw_ws_bytes": (_sz, [_i, _i]),
    "cvb_dtw_mcd": (_i, [_i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "cvb_mcd_aligned": (_i, [_i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "cvb_dropout_mask": (_i, [_sz, _f, _u64, _u64, _vp, _vp, _vp]),
    "cvb_state_advance": (_i, [_vp, _u64, _u64, _vp]),
    "cvb_adam_step": (_i, [_sz, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _vp, _f, _vp]),
    "cvb_weights_changed": (_i, []),
    "cvb_reserve_workspace": (_i, [_sz]),
    "cvb_profile_enable": (_i, [_i]),
    "cvb_profile_reset": (_i, []),
    "cvb_profile_summary": (_i, [_i, C.POINTER(C.c_float), C.POINTER(_i)]),
    "cvb_launch_count": (C.c_longlong, []),
    "cvb_gemm_tc": (_i, [_i, _i, _i, _i, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _vp]),
    "cvb_gemm": (_i, [_i, _i, _i, _i, _i, _f, _vp, _i, _vp, _i, _f, _vp, _i, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C cyclevae_vc_b200/csrc`). cyclevae_vc_b200 has no CPU / eager fallback.")
    # the library links libcublas / libcudart; torch has usually loaded them already, otherwise fall back to the toolkit's
    try:
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    except OSError:
        for name in ("libcudart.so.12", "libcublasLt.so.12", "libcublas.so.12"):
            for d in ("/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib"):
                p = os.path.join(d, name)
                if os.path.exists(p):
                    C.CDLL(p, mode=C.RTLD_GLOBAL)
                    break
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    v = lib.cvb_abi_version()
    if v != ABI_VERSION:
        raise ImportError(f"libcyclevae_b200.so ABI {v} != binding ABI {ABI_VERSION}: rebuild the library")
    return lib


lib = _load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.cvb_last_error()
        raise RuntimeError(f"libcyclevae_b200 {what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a contiguous fp32/int32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("cyclevae_vc_b200 runs on CUDA only (no CPU fallback): got a CPU tensor")
    if not t.is_contiguous():
        raise RuntimeError("internal error: non-contiguous tensor passed to the C ABI")
    return t.data_ptr()
