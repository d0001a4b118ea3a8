"""ctypes binding of libmimic_b200.so (include/mimic_b200.h).  There is NO fallback: if the
library cannot be loaded (and cannot be built because nvcc is absent) importing this module
raises, and every op raises `MimicError` on a non-zero status."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmimic_b200.so")

BF16, F32 = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2
IMPL_BY_NAME = {"auto": IMPL_AUTO, "simt": IMPL_SIMT, "tcgen05": IMPL_TCGEN05}


class MimicError(RuntimeError):
    pass


class AttnDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "H", "Hkv", "T", "S", "d")] + \
               [(n, C.c_int64) for n in ("q_sb", "q_sh", "q_st", "k_sb", "k_sh", "k_ss",
                                         "v_sb", "v_sh", "v_ss", "o_sb", "o_st", "o_sh")] + \
               [("scale", C.c_float), ("causal", C.c_int32), ("active", C.c_int32),
                ("dtype", C.c_int32), ("impl", C.c_int32)]


ABI_VERSION = 2          # MIMIC_ABI_VERSION of include/mimic_b200.h this front-end was written against

EXPORTS = {
    # name: (restype, argtypes)
    "mimic_abi_version": (C.c_int, []),
    "mimic_adamw_workspace": (C.c_size_t, []),
    "mimic_last_error": (C.c_char_p, []),
    "mimic_last_attn_impl": (C.c_char_p, []),
    "mimic_debug_set_trace": (None, [C.c_void_p, C.c_size_t]),
    "mimic_debug_force_fwd_kernel": (None, [C.c_int]),
    "mimic_shifted_attn_fwd": (C.c_int, [C.POINTER(AttnDesc)] + [C.c_void_p] * 12),
    "mimic_shifted_attn_bwd_workspace": (C.c_size_t, [C.POINTER(AttnDesc)]),
    "mimic_shifted_attn_bwd": (C.c_int, [C.POINTER(AttnDesc)] + [C.c_void_p] * 9 + [C.c_int64] * 3 +
                               [C.c_void_p] * 10 + [C.c_size_t, C.c_void_p]),
    "mimic_shifted_attn_bwd_lse": (C.c_int, [C.POINTER(AttnDesc)] + [C.c_void_p] * 9 + [C.c_int64] * 3 +
                                   [C.c_void_p] * 11 + [C.c_size_t, C.c_void_p]),
    "mimic_norm_shift_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_void_p]),
    "mimic_norm_shift_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "mimic_record_gather": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_void_p]),
    "mimic_row_plan": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mimic_mse_workspace": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "mimic_mse_fwd_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mimic_cos_fwd_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mimic_kl_fwd_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                   C.c_void_p]),
    "mimic_rope_qk": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32] +
                      [C.c_int32] * 6 + [C.c_void_p]),
    "mimic_ce_workspace": (C.c_size_t, [C.c_int32]),
    "mimic_ce_fwd_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_float,
                                   C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mimic_adamw_step_groups": (C.c_int, [C.c_void_p] * 6 + [C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                                            C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "mimic_adamw_step": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                                     C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
}


def _bind(path: str) -> C.CDLL:
    lib = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)       # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


def _load() -> C.CDLL:
    import fcntl
    from . import build as _build
    try:
        if not _build.up_to_date():   # content hash of csrc/ + include/ against the one recorded at build time
            with open(LIB_PATH + ".lock", "w") as lock:  # ranks of one job import at the same time: one builds, the rest wait
                fcntl.flock(lock, fcntl.LOCK_EX)
                if not _build.up_to_date():
                    _build.build(force=True)             # needs nvcc
    except RuntimeError as e:
        if not os.path.exists(LIB_PATH):
            raise                     # no library and no way to build it: fail loudly, there is no fallback path
        # A library exists but it was NOT built from the sources in the tree and could not be rebuilt: kernels and
        # checked-in code may disagree.  Refuse, unless the caller explicitly accepts the stale binary.
        if not os.environ.get("MIMIC_ALLOW_STALE_LIB"):
            raise ImportError(f"libmimic_b200.so is stale (sources changed since it was built) and the rebuild failed: {e}\n"
                              "fix the build, or set MIMIC_ALLOW_STALE_LIB=1 to bind the old binary anyway") from e
        import warnings
        warnings.warn(f"binding a STALE libmimic_b200.so (MIMIC_ALLOW_STALE_LIB set); rebuild failed: {e}", RuntimeWarning)
    lib = _bind(LIB_PATH)             # AttributeError when a prebuilt library is older than the header
    if lib.mimic_abi_version() != ABI_VERSION:
        raise ImportError(f"libmimic_b200.so ABI {lib.mimic_abi_version()} != {ABI_VERSION}; rebuild (python -m mimic_b200.build --force)")
    return lib


lib = _load()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise MimicError(f"{what} failed (status {rc}): {lib.mimic_last_error().decode()}")


def last_attn_impl() -> str:
    return lib.mimic_last_attn_impl().decode()
