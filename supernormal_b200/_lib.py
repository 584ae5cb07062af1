"""ctypes binding of libsnb200.so (the C ABI declared in include/snb200.h).

There is deliberately NO fallback: if the CUDA library is missing or an entry point returns an
error the call raises.  Prototypes are parsed from the header itself so the Python side cannot
drift from the ABI.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "snb200.h")
LIB_PATH = os.path.join(_HERE, "lib", "libsnb200.so")

SNB_MAX_LEVELS = 16


class HashGridMeta(C.Structure):
    _fields_ = [("n_levels", C.c_uint32), ("offsets", C.c_uint32 * (SNB_MAX_LEVELS + 1)),
                ("scales", C.c_float * SNB_MAX_LEVELS), ("resolutions", C.c_uint32 * SNB_MAX_LEVELS)]


_CTYPE = {"int32_t": C.c_int32, "int64_t": C.c_int64, "uint32_t": C.c_uint32, "uint64_t": C.c_uint64, "float": C.c_float,
          "snb_stream_t": C.c_void_p}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """{symbol: (restype, [argtypes])} for every `snb_*` function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(const char \*|int32_t|uint32_t|int64_t)\s*(snb_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        argtypes = []
        for a in [s.strip() for s in args.split(",")]:
            if a in ("void", ""):
                continue
            if "*" in a:
                argtypes.append(C.c_void_p)
            else:
                argtypes.append(_CTYPE[a.split()[-2] if a.split()[0] != "const" else a.split()[1]])
        restype = C.c_char_p if "char" in ret else _CTYPE[ret]
        protos[name] = (restype, argtypes)
    return protos


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m supernormal_b200.build` "
                "(nvcc, sm_100a). supernormal_b200 has no CPU or PyTorch fallback.")
        _lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in parse_header().items():
            fn = getattr(_lib, name)  # AttributeError if the header and the .so disagree
            fn.restype, fn.argtypes = restype, argtypes
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"libsnb200 error {rc}: {lib().snb_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL). The tensor must be CUDA + contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")  # nerfacc's own message, NA/ray_marching.py:131
    assert t.is_contiguous(), "libsnb200 needs contiguous tensors"
    return t.data_ptr() or None


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def stream() -> int:
    """cudaStream_t of torch's current stream on the current device.  The raw C accessors cost ~0.5 us; torch.cuda.current_stream()
    builds a Stream object (~12 us) and a training step asks several times (scripts/host_profile_e2e.py)."""
    if _raw_stream is not None and _cur_device is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


LAUNCH_COUNT = 0
# kernels launched per entry point (memsets not counted)
KERNELS = {"snb_occgrid_binarize": 2, "snb_compact_samples": 2, "snb_max_i64": 2, "snb_train_fwd_bwd": 8,
           "snb_train_optim": 1, "snb_occgrid_update_fused": 3, "snb_train_fwd_bwd_lean": 5, "snb_train_tail": 1}


PROFILE = None  # set to a list to record (name, start_event, end_event) around every call


def call(name: str, *args) -> None:
    global LAUNCH_COUNT
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(lib(), name)(*args, stream()))
        e1.record()
        PROFILE.append((name, e0, e1))
    else:
        check(getattr(lib(), name)(*args, stream()))
    LAUNCH_COUNT += KERNELS.get(name, 1)


def make_meta(n_levels, log2_hashmap_size, base_resolution, per_level_scale) -> Tuple[HashGridMeta, int]:
    m = HashGridMeta()
    total = lib().snb_hashgrid_make_meta(n_levels, log2_hashmap_size, base_resolution, per_level_scale, C.byref(m))
    if total == 0:
        raise RuntimeError(lib().snb_last_error().decode())
    return m, int(total)
