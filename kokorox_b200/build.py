"""Build libkkx.so (sm_100a only) in-tree with nvcc.

This is what ``kokorox/build.rs`` would do in the reference tree (today that file
only links libsonic/libpcaudio, /root/reference/kokorox/build.rs:113-118): compile
``csrc/*.cu`` with ``-gencode arch=compute_100a,code=sm_100a`` and link one shared library that
exports the C ABI of ``include/kkx.h``.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libkkx.so")
LOADGEN = os.path.join(OUT_DIR, "kkx_loadgen")   # native load generator (csrc/loadgen.cpp), a client of the C ABI
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]
FLAGS += os.environ.get("KKX_NVCC_EXTRA", "").split()   # e.g. -DKKX_ARB_TIMING for the diagnostic build


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".h", ".cuh", ".cpp")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "libkkx.stamp")
    dig = _digest()

    def fresh():
        return os.path.exists(LIB) and os.path.exists(LOADGEN) and os.path.exists(stamp) and open(stamp).read() == dig
    if not force and fresh():
        return LIB
    # several ranks of one torchrun launch may get here at once on a box without a prebuilt library: one builds,
    # the others wait on the lock and then find the result
    import fcntl
    with open(os.path.join(OUT_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():
                return LIB
            return _build_locked(stamp, dig, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(stamp: str, dig: str, verbose: bool) -> str:
    objs = []

    def cc(src):
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(cc, _sources()))
    tmp = LIB + ".tmp%d" % os.getpid()
    cmd = [NVCC, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)          # a process that already mapped the old library keeps its copy
    # the native load generator links the shared library like any other client of include/kkx.h
    tmpg = LOADGEN + ".tmp%d" % os.getpid()
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-pthread", os.path.join(CSRC, "loadgen.cpp"), "-o", tmpg,
           "-L" + OUT_DIR, "-lkkx", "-Wl,-rpath,$ORIGIN", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"loadgen link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmpg, LOADGEN)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
