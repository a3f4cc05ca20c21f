"""ctypes binding of include/bcmpc.h.  No fallback: a missing library is an ImportError-class
failure at first use, a missing GPU is a RuntimeError from bcmpc_create."""
import ctypes as C
import os

import numpy as np

from . import _build

MAX_ACT = 8
ABI_VERSION = 4      # BCMPC_ABI_VERSION of include/bcmpc.h this binding was written against


class Config(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("n_hidden", C.c_int32),
                ("hidden_size", C.c_int32), ("activation", C.c_int32), ("num_paths", C.c_int32),
                ("horizon", C.c_int32), ("max_states", C.c_int32), ("precision", C.c_int32),
                ("proposal", C.c_int32), ("device", C.c_int32), ("first_proposal", C.c_int32),
                ("path_offset", C.c_int64), ("total_paths", C.c_int64), ("explore", C.c_double)]


class Record(C.Structure):
    _fields_ = [("cost", C.c_double), ("index", C.c_int64), ("action", C.c_double * MAX_ACT)]


RECORD_DTYPE = np.dtype([("cost", "<f8"), ("index", "<i8"), ("action", "<f8", (MAX_ACT,))])
assert RECORD_DTYPE.itemsize == C.sizeof(Record) == 80

ACT_RELU, ACT_TANH = 0, 1
PREC_FP32_VERIFY, PREC_BF16_FAST = 0, 1
PROP_UNIFORM, PROP_POLICY_GAUSSIAN, PROP_POLICY_MIX, PROP_NET_PLUS_UNIFORM, PROP_POLICY_MEAN = 0, 1, 2, 3, 4

_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
# name -> (restype, argtypes); every symbol include/bcmpc.h declares
SIGNATURES = {
    "bcmpc_abi_version": (C.c_int, []),
    "bcmpc_last_error": (C.c_char_p, [_P]),
    "bcmpc_create": (C.c_int, [C.POINTER(Config), _PP]),
    "bcmpc_destroy": (C.c_int, [_P]),
    "bcmpc_set_dynamics": (C.c_int, [_P, _PP, _PP, C.c_int32]),
    "bcmpc_set_dynamics_device": (C.c_int, [_P, _PP, _PP, C.c_int32, C.c_int32, _P]),
    "bcmpc_set_policy_device": (C.c_int, [_P, C.c_int32, _P, _PP, _PP, C.c_int32, _P, _P, _P, C.c_float, _P]),
    "bcmpc_set_layer_norm": (C.c_int, [_P, _PP, _PP, C.c_int32, C.c_float]),
    "bcmpc_set_reward_head": (C.c_int, [_P, _PP, _PP, C.c_double, C.c_double, C.c_double]),
    "bcmpc_set_reward_layer_norm": (C.c_int, [_P, _P, _P, C.c_float]),
    "bcmpc_set_layer_norm_device": (C.c_int, [_P, _PP, _PP, C.c_int32, C.c_float, _P]),
    "bcmpc_set_reward_head_device": (C.c_int, [_P, _PP, _PP, C.c_int32, C.c_double, C.c_double, C.c_double, _P]),
    "bcmpc_set_reward_layer_norm_device": (C.c_int, [_P, _P, _P, C.c_float, _P]),
    "bcmpc_predict_reward": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P]),
    "bcmpc_set_normalization": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "bcmpc_set_action_bounds": (C.c_int, [_P, _P, _P]),
    "bcmpc_set_policy": (C.c_int, [_P, C.c_int32, _P, _PP, _PP, _P, _P, _P, C.c_float]),
    "bcmpc_plan": (C.c_int, [_P, _P, C.c_int32, _P, C.c_uint64, _P]),
    "bcmpc_plan_device": (C.c_int, [_P, _P, C.c_int32, _P, C.c_uint64, _P, _P]),
    "bcmpc_plan_device_seeded": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P]),
    "bcmpc_rollout_debug": (C.c_int, [_P, _P, _P, C.c_uint64, _P, _P, _P]),
    "bcmpc_predict": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "bcmpc_policy_act": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "bcmpc_cheetah_cost": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "bcmpc_sample_injection": (C.c_int, [_P, C.c_uint64, C.c_int32, _P]),
    "bcmpc_reduce_records": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P]),
    "bcmpc_current_device": (C.c_int, [C.POINTER(C.c_int32)]),
    "bcmpc_peer_exchange_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "bcmpc_set_peer_exchange": (C.c_int, [_P, C.c_int32, C.c_int32, _PP]),
    "bcmpc_kernel_launches": (C.c_int64, [_P]),
    "bcmpc_enable_timing": (C.c_int, [_P, C.c_int32]),
    "bcmpc_last_kernel_ms": (C.c_int, [_P, _P]),
}

_lib = None


def lib():
    """Loads (building first if the sources are newer) libbcmpc.so."""
    global _lib
    if _lib is None:
        path = os.environ.get("BCMPC_LIB") or _build.LIB        # (BCMPC_LIB: a developer A/B build, tools/ab_deep.sh)
        if path == _build.LIB and _build.stale():
            try:
                _build.build()
            except Exception as e:  # no nvcc on the box: use the shipped .so if there is one, but say so
                if not os.path.exists(path):
                    raise ImportError(f"libbcmpc.so is missing and could not be built: {e}") from e
                import warnings
                warnings.warn(f"bc_mpc_b200: the CUDA sources are newer than libbcmpc.so and the rebuild failed ({e}); "
                              "loading the existing library", RuntimeWarning)
        if not os.path.exists(path):
            raise ImportError("libbcmpc.so is missing: run `python -m bc_mpc_b200._build`")
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.bcmpc_abi_version() != ABI_VERSION:
            raise ImportError(f"libbcmpc.so ABI version {L.bcmpc_abi_version()} != {ABI_VERSION} expected by this package: "
                              "rebuild with `python -m bc_mpc_b200._build`")
        _lib = L
    return _lib


def current_device():
    """The calling thread's current CUDA device (what `device=None` resolves to in the helper planners)."""
    dev = C.c_int32(0)
    check(lib().bcmpc_current_device(C.byref(dev)))
    return int(dev.value)


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def check(rc, handle=None):
    if rc != 0:
        msg = lib().bcmpc_last_error(handle)
        raise RuntimeError(f"bcmpc error {rc}: {msg.decode() if msg else ''}")
