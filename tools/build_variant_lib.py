"""Experiment build: recompile ONE source of the library with extra -D definitions and link it with the objects of
the regular build -> kokorox_b200/lib/libkkx_<tag>.so (use with KKX_LIB=...).
usage: python tools/build_variant_lib.py <tag> <source.cu> -DNAME=VALUE ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokorox_b200 import build as b  # noqa: E402

tag, src, defs = sys.argv[1], sys.argv[2], sys.argv[3:]
b.build()
obj = f"/tmp/kkx_variant_{tag}.o"
subprocess.run([b.NVCC, *b.FLAGS, *defs, "-c", os.path.join(b.CSRC, src), "-o", obj], check=True)
objs = [obj if os.path.basename(s) == src else os.path.join(b.OUT_DIR, os.path.basename(s)[:-3] + ".o") for s in b._sources()]
out = os.path.join(b.OUT_DIR, f"libkkx_{tag}.so")
subprocess.run([b.NVCC, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                "-cudart", "static"], check=True)
print(out)
