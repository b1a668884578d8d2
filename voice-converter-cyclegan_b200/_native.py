"""ctypes binding of libcgvc.so (the C ABI declared in include/cgvc.h).

There is deliberately no fallback: if the shared library is missing or fails to load, importing the
package's compute classes raises.  PyTorch is used by callers for device storage only.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

ARENA_PARAM, ARENA_GRAD, ARENA_ADAM_M, ARENA_ADAM_V, ARENA_WORK = range(5)
PREC_FP32_SIMT, PREC_BF16X3, PREC_BF16, PREC_F16F8 = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32_SIMT, "fp32_simt": PREC_FP32_SIMT, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16, "f16f8": PREC_F16F8}
ERR_DIRECTION = -4
ERR_UNSUPPORTED = -5

LOSS_NAMES = ("cycle_loss", "identity_loss", "generator_loss_A2B", "generator_loss_B2A", "generator_loss",
              "discriminator_loss_A", "discriminator_loss_B", "discriminator_loss")


class Config(C.Structure):
    _fields_ = [("num_features", C.c_int), ("max_batch", C.c_int), ("max_frames", C.c_int),
                ("precision", C.c_int), ("device", C.c_int), ("train", C.c_int)]


class CgvcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libcgvc error %d: %s" % (code, msg))
        self.code = code


def _declare(lib):
    vp, ci, cf, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    P = C.POINTER
    sig = {
        "cgvc_abi_version": (ci, []),
        "cgvc_create": (ci, [P(Config), P(vp)]),
        "cgvc_destroy": (ci, [vp]),
        "cgvc_last_error": (C.c_char_p, [vp]),
        "cgvc_arena_bytes": (ci, [vp, ci, P(sz)]),
        "cgvc_bind_arena": (ci, [vp, ci, vp, sz]),
        "cgvc_param_count": (ci, [vp, P(ci), P(sz)]),
        "cgvc_param_info": (ci, [vp, ci, P(C.c_char_p), P(sz), P(ci), P(ci * 4)]),
        "cgvc_params_updated": (ci, [vp, vp]),
        "cgvc_set_adam_step": (ci, [vp, C.c_longlong]),
        "cgvc_get_adam_step": (ci, [vp, P(C.c_longlong)]),
        "cgvc_train_step": (ci, [vp, vp, vp, ci, ci, cf, cf, cf, cf, vp, vp, vp, vp]),
        "cgvc_compute_gradients": (ci, [vp, vp, vp, ci, ci, cf, cf, vp, vp, vp, vp]),
        "cgvc_adam_step": (ci, [vp, cf, cf, cf, vp]),
        "cgvc_generator_forward": (ci, [vp, ci, vp, vp, ci, ci, vp]),
        "cgvc_discriminator_forward": (ci, [vp, ci, vp, vp, ci, ci, vp]),
        "cgvc_debug_activation": (ci, [vp, C.c_char_p, vp, sz, P(sz), vp]),
        "cgvc_comm_unique_id": (ci, [vp, vp]),
        "cgvc_comm_init": (ci, [vp, vp, ci, ci]),
        "cgvc_comm_destroy": (ci, [vp]),
        "cgvc_allreduce_grads": (ci, [vp, vp]),
        "cgvc_sample_plan": (ci, [vp, vp, ci, vp, ci, C.c_ulonglong, C.c_longlong, ci, vp, vp, vp]),
        "cgvc_gather_minibatch": (ci, [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp, vp, vp]),
        "cgvc_kernel_launches": (ci, [P(C.c_ulonglong)]),
        "cgvc_set_option": (ci, [vp, C.c_char_p, ci]),
        "cgvc_profile_enable": (ci, [ci]),
        "cgvc_profile_collect": (ci, [P(C.c_double), P(C.c_double), P(C.c_longlong)]),
        "cgvc_profile_launches": (ci, [P(C.c_double), P(C.c_double), P(C.c_longlong), ci, P(ci)]),
        "cgvc_conv_forward": (ci, [vp, ci, vp, vp, vp, vp] + [ci] * 9 + [vp]),
        "cgvc_conv_backward": (ci, [vp, ci, vp, vp, vp, vp, vp, vp] + [ci] * 9 + [vp]),
        "cgvc_in_glu_forward": (ci, [vp] * 8 + [ci] * 4 + [vp]),
        "cgvc_in_glu_backward": (ci, [vp] * 13 + [ci] * 4 + [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTED_SYMBOLS = None


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load libcgvc.so, building it in-tree with nvcc first if it is missing or stale."""
    global _LIB, EXPORTED_SYMBOLS
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if build_if_missing and _build.is_stale():
        _build.build()
    if not os.path.exists(path):
        raise RuntimeError("libcgvc.so not found at %s (run __graft_entry__.build()); there is no CPU fallback" % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    EXPORTED_SYMBOLS = sorted(_declare(lib).keys())
    if lib.cgvc_abi_version() != 1:
        raise RuntimeError("libcgvc.so ABI version mismatch")
    _LIB = lib
    return lib


def check(handle, code):
    if code != 0:
        msg = _LIB.cgvc_last_error(handle)
        raise CgvcError(code, msg.decode() if msg else "?")
