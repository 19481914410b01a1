"""A/B builds for kernel experiments: python tools/ab_build.py NAME SOURCE.cu [-DFLAG ...]
Recompiles ONE source of the library with extra flags and links it with the objects of the regular build into
ab/libgvigp_NAME.so (git-ignored, travels with gpurun); run a tool against it with GVI_LIB_PATH=ab/libgvigp_NAME.so."""
import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gvi_gaussian_process_b200 import build as B

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
root = os.path.join(B.HERE, "..", "ab")
os.makedirs(root, exist_ok=True)
obj = os.path.join(root, f"{name}_{src.replace('.cu', '.o')}")
base = [f for f in B.NVCC_FLAGS if not f.startswith("--use_fast_math")]
subprocess.run([B._nvcc(), *base, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj], check=True)
objs = [obj if s == src else os.path.join(B.HERE, "build", s.replace(".cu", ".o")) for s in B.SOURCES]
out = os.path.join(root, f"libgvigp_{name}.so")
subprocess.run([B._nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared", "-ldl"], check=True)
print(out)
