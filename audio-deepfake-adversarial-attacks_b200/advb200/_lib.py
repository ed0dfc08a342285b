"""ctypes binding of libadvb200.so (the C ABI declared in include/advb200.h).

The library is mandatory: there is no Python/PyTorch fallback for any arithmetic.  ``load()`` raises if the shared
object is missing (build it with ``python audio-deepfake-adversarial-attacks_b200/build.py`` or
``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadvb200.so")
LIB_PATH = os.environ.get("ADVB_LIB", LIB_PATH)  # tuning experiments: an alternative build of the same library

MODEL_LCNN, MODEL_SPECRNET, MODEL_RAWNET3, MODEL_FRONTEND_ONLY = 1, 2, 3, 4
FRONTEND_NONE, FRONTEND_LFCC, FRONTEND_MFCC = 0, 1, 2
ATTACK_FGSM, ATTACK_PGD, ATTACK_PGDL2, ATTACK_FAB, ATTACK_CW = 1, 2, 3, 4, 5
NORM_LINF, NORM_L2 = 0, 1
GRAD_CE, GRAD_LOGIT = 0, 1


class TensorRef(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ptr", C.c_void_p), ("numel", C.c_int64)]


class ModelDesc(C.Structure):
    _fields_ = [
        ("model_kind", C.c_int), ("frontend_kind", C.c_int), ("device", C.c_int), ("max_batch", C.c_int),
        ("n_samples", C.c_int), ("n_tensors", C.c_int), ("tensors", C.POINTER(TensorRef)),
    ]


class AttackDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("eps", C.c_float), ("alpha", C.c_float), ("steps", C.c_int), ("eps_div", C.c_float),
        ("alpha_max", C.c_float), ("eta", C.c_float), ("beta", C.c_float), ("c", C.c_float), ("kappa", C.c_float),
        ("lr", C.c_float), ("n_global_batch", C.c_int), ("targeted", C.c_int), ("norm", C.c_int), ("target_labels", C.c_void_p),
    ]


# name -> (restype, argtypes): every symbol include/advb200.h declares
SIGNATURES = {
    "advb_version": (C.c_int, []),
    "advb_last_error": (C.c_char_p, []),
    "advb_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(ModelDesc)]),
    "advb_destroy": (None, [C.c_void_p]),
    "advb_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "advb_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "advb_rebind": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(TensorRef)]),
    "advb_invalidate_weights": (C.c_int, [C.c_void_p]),
    "advb_attack": (C.c_int, [C.c_void_p, C.POINTER(AttackDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_int, C.c_int, C.c_void_p]),
    "advb_attack_minmax": (C.c_int, [C.c_void_p, C.POINTER(AttackDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_int, C.c_void_p]),
    "advb_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "advb_grad": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                            C.c_int, C.c_void_p]),
    "advb_frontend_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "advb_frontend_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "advb_minmax": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "advb_revert_minmax": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "advb_projection_linf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "advb_projection_l2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "advb_mel_spec_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "advb_row_diff_norms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "advb_debug_stage": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]),
    "advb_launch_count": (C.c_int64, [C.c_void_p]),
    "advb_profile_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "advb_profile_end": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int64]),
    "advb_xrank_export": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "advb_xrank_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
    "advb_xrank_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
}
XRANK_HANDLE_BYTES = 64

_lib = None


def load():
    """Load the shared library once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the advb200 CUDA extension is not built and there is no CPU fallback "
            "(run `python audio-deepfake-adversarial-attacks_b200/build.py`)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError("libadvb200: " + load().advb_last_error().decode("utf-8", "replace"))
