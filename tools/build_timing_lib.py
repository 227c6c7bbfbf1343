"""Diagnostic build: libkkx with per-role cycle counters in the tcgen05 kernels (-DKKX_TC_TIMING -DKKX_ARB_TIMING
-DKKX_EXPERIMENTS) -> kokorox_b200/lib/libkkx_timing.so.  Run a workload with KKX_LIB=<that file> KKX_ARB_TIMING=1
and the library prints the phase cycles of one CTA per launch family to stderr after each run."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokorox_b200 import build as b  # noqa: E402

OUT = os.path.join(b.OUT_DIR, "libkkx_timing.so")
OBJ = "/tmp/kkx_timing_objs"
os.makedirs(OBJ, exist_ok=True)
flags = b.FLAGS + ["-DKKX_EXPERIMENTS"] + (["-DKKX_TC_TIMING", "-DKKX_ARB_TIMING"] if os.environ.get("KKX_NO_COUNTERS") != "1" else [])
if os.environ.get("KKX_NO_COUNTERS") == "1":
    OUT = os.path.join(b.OUT_DIR, "libkkx_exp.so")          # experiment switches only (no cycle counters in the kernels)
    OBJ = "/tmp/kkx_exp_objs"
    extra = os.environ.get("KKX_EXP_DEFS", "").split()        # e.g. KKX_EXP_DEFS="-DKKX_ARB_GP6" -> libkkx_exp2.so
    if extra:
        flags += extra
        OUT = os.path.join(b.OUT_DIR, "libkkx_exp2.so")
        OBJ = "/tmp/kkx_exp2_objs"
    os.makedirs(OBJ, exist_ok=True)


def cc(src):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    r = subprocess.run([b.NVCC, *flags, "-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    return obj


with ThreadPoolExecutor(max_workers=8) as ex:
    objs = list(ex.map(cc, b._sources()))
r = subprocess.run([b.NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                    "-cudart", "static"], capture_output=True, text=True)
if r.returncode != 0:
    raise RuntimeError(r.stderr)
print(OUT)
