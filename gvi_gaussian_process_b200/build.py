"""In-tree build of libgvigp.so (sm_100a only) with nvcc.  The .so is git-ignored but travels with gpurun."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["gram.cu", "gram_dmma.cu", "chol.cu", "predict.cu", "grad.cu", "risk.cu", "select.cu", "classify.cu", "exact.cu", "optim.cu", "grad_wide.cu", "comm.cu", "prng.cu"]
LIB_PATH = os.path.join(HERE, "libgvigp.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "predict.cuh"), os.path.join(CSRC, "comm.cuh"),
                                                        os.path.join(HERE, "..", "include", "gvigp.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    flags += os.environ.get("GVI_NVCC_FLAGS", "").split()  # e.g. -DGVI_PHASE_CLOCKS (profiling builds)
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        cmd = [_nvcc(), *flags, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{out}")
    cmd = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared", "-lcublas", "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


FFI_SOURCE = os.path.join(CSRC, "ffi", "gvigp_xla_ffi.cc")
FFI_LIB_PATH = os.path.join(HERE, "libgvigp_ffi.so")


def xla_ffi_include_dir():
    """jaxlib ships the XLA FFI headers (xla/ffi/api/ffi.h) under jaxlib/include; None when jaxlib is absent."""
    try:
        import jaxlib
    except ImportError:
        return None
    inc = os.path.join(os.path.dirname(jaxlib.__file__), "include")
    return inc if os.path.exists(os.path.join(inc, "xla", "ffi", "api", "ffi.h")) else None


def build_ffi() -> str:
    """libgvigp_ffi.so: the XLA-FFI handlers over the C ABI (csrc/ffi/gvigp_xla_ffi.cc).  Needs a jaxlib that ships the
    FFI headers; this repository's image has none, so the adapter has never been compiled here (build() does not call
    this)."""
    inc = xla_ffi_include_dir()
    if inc is None:
        raise RuntimeError("no jaxlib with xla/ffi/api/ffi.h importable: the XLA-FFI adapter cannot be built in this "
                           "environment (the C ABI in libgvigp.so is the tested surface)")
    build()
    cmd = [_nvcc(), "-std=c++17", "-O2", "-shared", "-Xcompiler", "-fPIC", "-I", inc, FFI_SOURCE, "-o", FFI_LIB_PATH,
           "-L", HERE, "-l:libgvigp.so", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN", "-cudart", "shared"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"building the XLA-FFI adapter failed:\n{r.stdout}")
    return FFI_LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
