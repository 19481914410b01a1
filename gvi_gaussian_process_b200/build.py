"""In-tree build of libgvigp.so (sm_100a only) with nvcc.  The .so is git-ignored but travels with gpurun."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["gram.cu", "gram_mma.cu", "gram_dmma.cu", "chol.cu", "chol_coop.cu", "dgemm.cu", "predict.cu", "grad.cu", "risk.cu", "select.cu", "classify.cu", "exact.cu", "optim.cu", "grad_wide.cu", "comm.cu", "prng.cu", "diag.cu"]
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


TEST_LIB_PATH = os.path.join(HERE, "libgvigp_test.so")  # same sources + -DGVI_TEST_HOOKS (include/gvigp_test.h)


def _deps():
    return [os.path.join(CSRC, s) for s in SOURCES] + [
        os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "predict.cuh"), os.path.join(CSRC, "comm.cuh"), os.path.join(CSRC, "dgemm.cuh"),
        os.path.join(HERE, "..", "include", "gvigp.h"), os.path.join(HERE, "..", "include", "gvigp_test.h")]


def needs_build(lib_path: str = LIB_PATH) -> bool:
    if not os.path.exists(lib_path):
        return True
    t = os.path.getmtime(lib_path)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile_and_link(lib_path: str, obj_dir: str, extra_flags, verbose: bool):
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + list(extra_flags)
    flags += os.environ.get("GVI_NVCC_FLAGS", "").split()  # e.g. -DGVI_PHASE_CLOCKS (profiling builds)
    os.makedirs(obj_dir, exist_ok=True)
    objs, procs = [], []
    for s in SOURCES:
        obj = os.path.join(obj_dir, s.replace(".cu", ".o"))
        cmd = [_nvcc(), *flags, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)

    def finish():
        for s, p in procs:
            out, _ = p.communicate()
            if verbose and out:
                print(out)
            if p.returncode != 0:
                raise RuntimeError(f"nvcc failed for {s}:\n{out}")
        cmd = [_nvcc(), "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared", "-ldl"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
        return lib_path

    return finish


def build(force: bool = False, verbose: bool = False) -> str:
    """libgvigp.so (the product) and libgvigp_test.so (tests / tools: the same sources with -DGVI_TEST_HOOKS)."""
    jobs = []
    if force or needs_build(LIB_PATH):
        jobs.append(_compile_and_link(LIB_PATH, os.path.join(HERE, "build"), [], verbose))
    if force or needs_build(TEST_LIB_PATH):
        jobs.append(_compile_and_link(TEST_LIB_PATH, os.path.join(HERE, "build", "test"), ["-DGVI_TEST_HOOKS"], verbose))
    for finish in jobs:  # both compilations run concurrently; wait and link
        finish()
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
