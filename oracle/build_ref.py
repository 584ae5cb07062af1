"""Build recipe for oracle/_ref: the UNMODIFIED reference nerfacc 0.3.5 CUDA extension.

TEST INFRASTRUCTURE ONLY.  This compiles the reference's own sources *where they lie*
under /root/reference (third_parties/nerfacc-0.3.5/nerfacc-0.3.5/nerfacc/cuda/csrc/*.cu,
registered by csrc/pybind.cu:162-206) for sm_100a with the reference's own flag (-O3,
nerfacc/cuda/_backend.py:43-44) through torch.utils.cpp_extension's include/link
settings.  Nothing is copied into the repo; the only output is the pybind module
oracle/_ref/nerfacc_ref_C*.so (git-ignored, travels to the GPU box with the snapshot).

The GPU parity tests load it to compare our kernels bit-for-bit with the reference
kernels on the same B200 (ray marching, patch weights fwd/bwd, CUB transmittance).
The product never imports it.

Usage:  python oracle/build_ref.py            (no-op if sources are absent or .so fresh)
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
import sysconfig

REF_CSRC = "/root/reference/third_parties/nerfacc-0.3.5/nerfacc-0.3.5/nerfacc/cuda/csrc"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
MOD = "nerfacc_ref_C"


def so_path() -> str:
    return os.path.join(OUT_DIR, MOD + ".so")


def build(verbose: bool = False) -> str | None:
    if not os.path.isdir(REF_CSRC):
        return so_path() if os.path.exists(so_path()) else None
    srcs = sorted(glob.glob(os.path.join(REF_CSRC, "*.cu")))
    os.makedirs(OUT_DIR, exist_ok=True)
    out = so_path()
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    import torch
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths("cuda") + [sysconfig.get_paths()["include"]]
    libdirs = ce.library_paths("cuda")
    common = [
        "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
        "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
        # torch's default defines (needed by the two off-path files)
        "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
        "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
        f"-DTORCH_EXTENSION_NAME={MOD}", "-DTORCH_API_INCLUDE_EXTENSION_H",
        f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
    ] + [f"-I{p}" for p in inc]
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(OUT_DIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        procs.append(subprocess.Popen(["nvcc", *common, "-c", s, "-o", o]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference nerfacc compile failed")
    link = ["nvcc", "-shared", "-o", out, *objs] + [f"-L{p}" for p in libdirs] + [
        "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
    ]
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    if verbose:
        print("built", out)
    return out


REF_MODELS = "/root/reference/models"
MODELS_OUT = os.path.join(OUT_DIR, "models")
MODEL_FILES = ("renderer", "fields")


def build_models(verbose: bool = False) -> str | None:
    """Byte-compile the reference's OWN models/renderer.py and models/fields.py, where they lie, into sourceless
    oracle/_ref/models/*.pyc (a compiled artefact like the .so above: git-ignored, travels to the GPU box, no source copied into
    the repo).  The `-m gpu` drop-in test executes this literal reference code on top of the product's nerfacc / tinycudann shims
    (tests/test_reference_literal.py); oracle/ref_models.py loads it."""
    if not os.path.isdir(REF_MODELS):
        return MODELS_OUT if all(os.path.exists(os.path.join(MODELS_OUT, f + ".pyc")) for f in MODEL_FILES) else None
    import py_compile
    os.makedirs(MODELS_OUT, exist_ok=True)
    for f in MODEL_FILES:
        src, out = os.path.join(REF_MODELS, f + ".py"), os.path.join(MODELS_OUT, f + ".pyc")
        if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            py_compile.compile(src, cfile=out, dfile=f"<reference>/models/{f}.py", doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            if verbose:
                print("compiled", out)
    return MODELS_OUT


def load():
    """Import the built module (requires torch; returns None when unavailable)."""
    p = so_path()
    if not os.path.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)

    spec = importlib.util.spec_from_file_location(MOD, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose=True))
    print(build_models(verbose=True))
