"""Builds genetic_b200/libgenetic_b200.so (CUDA kernels + C-ABI) for sm_100a with nvcc, in-tree.

    python -m genetic_b200.build [--force]

nvcc cross-compiles without a GPU. Objects are cached under genetic_b200/csrc/_build/ (git-ignored).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libgenetic_b200.so")
NVCC = os.environ.get("GX_NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr", "-diag-suppress", "128,20013"]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "genetic_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs if os.path.exists(h))


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    # (fft_variants.cu is a development sweep harness, not part of the C ABI: only built with GX_DEV=1)
    dev = os.environ.get("GX_DEV", "0") == "1"
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and (dev or f != "fft_variants.cu"))
    hdr_m = _deps()
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_m):
            jobs.append([NVCC] + FLAGS + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose and r.stderr:
            sys.stderr.write(r.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        run([NVCC, "-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart", "-lpthread"])
    build_host(force or bool(jobs))
    return LIB


HOST_DIR = os.path.join(os.path.dirname(HERE), "host")
HOST_PROGRAMS = ("genetic_b200_ic", "genetic_b200_chi2", "genetic_b200_zoom", "genetic_b200_modify")
HOST_BIN = os.path.join(HOST_DIR, "genetic_b200_ic")


def build_host(force=False, program="genetic_b200_ic"):
    """The C++ host layer (include/genetic_b200.hpp) and its drivers: plain g++, linked against the C ABI."""
    inc = os.path.join(os.path.dirname(HERE), "include")
    out = None
    for prog in HOST_PROGRAMS:
        src, exe = os.path.join(HOST_DIR, prog + ".cpp"), os.path.join(HOST_DIR, prog)
        newest = max(os.path.getmtime(p) for p in (src, os.path.join(inc, "genetic_b200.hpp"), os.path.join(inc, "genetic_b200.h"),
                                                   os.path.join(inc, "genetic_b200_zoom.hpp"),
                                                   os.path.join(inc, "genetic_b200_modif.hpp")))
        if force or not os.path.exists(exe) or os.path.getmtime(exe) < newest:
            cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-I", inc, src, "-L", HERE, "-lgenetic_b200",
                   "-Wl,-rpath,$ORIGIN/../genetic_b200", "-o", exe]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("g++ failed: %s\n%s" % (" ".join(cmd), r.stderr))
        if prog == program:
            out = exe
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
