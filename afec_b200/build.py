"""Builds afec_b200/csrc/libafec_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m afec_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libafec_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# FP64 descriptors must round like the reference: no fast-math; FMA contraction stays on for the
# FP64 stages (differences are ~1 ulp, far inside tolerance) -- the float32 stages that must be
# bit-exact (downmix, resampler, onset functions) use explicit __fadd_rn / __fmul_rn intrinsics.
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--prec-div=true", "--prec-sqrt=true",
         "--ftz=false", "-Xptxas", "-v" if os.environ.get("AFX_PTXAS_V") else "-O3"]

SOURCES = ["afx_api.cu", "afx_condition.cu", "afx_spectrum.cu", "afx_peaks.cu", "afx_bands.cu",
           "afx_pitch.cu", "afx_autocorr.cu", "afx_rhythm.cu", "afx_stats.cu", "afx_debug.cu", "afx_part.cu", "afx_highlevel.cu", "afx_pack.cu", "afx_ext.cu"]
DEFINES = {"afx_peaks.cu": "AFX_HAVE_PEAKS", "afx_bands.cu": "AFX_HAVE_BANDS", "afx_pitch.cu": "AFX_HAVE_PITCH",
           "afx_autocorr.cu": "AFX_HAVE_AUTOCORR", "afx_rhythm.cu": "AFX_HAVE_RHYTHM", "afx_stats.cu": "AFX_HAVE_STATS"}


def _stale(target: str, deps: list) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "afec_b200.h"))
    headers.append(os.path.abspath(__file__))
    defs = ["-D" + DEFINES[s] for s in srcs if s in DEFINES]
    objs, jobs = [], []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, s[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC] + ARCH + FLAGS + defs + ["-c", src, "-o", obj]
        stamp = obj + ".cmd"
        old = open(stamp).read() if os.path.exists(stamp) else ""
        if force or _stale(obj, [src] + headers) or old != " ".join(cmd):
            jobs.append(cmd)
            with open(stamp, "w") as f:
                f.write(" ".join(cmd))

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + cmd[-3])
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        # link under a temporary name and rename: a gpurun snapshot taken meanwhile never sees a half-written library
        run([NVCC] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB + ".tmp"] + objs + ["-cudart", "static"])
        os.replace(LIB + ".tmp", LIB)
    build_host(force=force or bool(jobs), verbose=verbose)
    return LIB


HOST = os.path.join(HERE, "host")
HOST_LIB = os.path.join(HOST, "libafec_b200_host.so")
CRAWLER = os.path.join(HOST, "afec-b200-crawler")
HOST_SOURCES = ["descriptors.cpp", "sqlite_pool.cpp", "direct_db_writer.cpp", "audio_reader.cpp", "gpu_analyser.cpp", "capi.cpp"]
SQLITE = "/usr/lib/x86_64-linux-gnu/libsqlite3.so.0"


def build_host(force: bool = False, verbose: bool = False):
    """C++ host adapter (reference extractor interface + afec-ll.db sink) and the crawler executable."""
    srcs = [os.path.join(HOST, s) for s in HOST_SOURCES]
    deps = srcs + [os.path.join(HOST, h) for h in os.listdir(HOST) if h.endswith(".h")] + [
        os.path.join(os.path.dirname(HERE), "include", "afec_b200.h"), LIB, os.path.abspath(__file__)]
    cxx = os.environ.get("CXX", "g++")
    common = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-pthread"]
    rpath = ["-Wl,-rpath,$ORIGIN/../csrc", "-Wl,-rpath,$ORIGIN"]

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("host build failed: " + cmd[-1])

    if force or _stale(HOST_LIB, deps):
        run([cxx] + common + ["-shared", "-o", HOST_LIB + ".tmp"] + srcs + ["-L" + CSRC, "-lafec_b200", SQLITE] + rpath)
        os.replace(HOST_LIB + ".tmp", HOST_LIB)
    if force or _stale(CRAWLER, deps + [os.path.join(HOST, "crawler_main.cpp"), HOST_LIB]):
        run([cxx] + common + ["-o", CRAWLER + ".tmp", os.path.join(HOST, "crawler_main.cpp"), "-L" + HOST, "-lafec_b200_host",
                              "-L" + CSRC, "-lafec_b200", SQLITE] + rpath)
        os.replace(CRAWLER + ".tmp", CRAWLER)
    return HOST_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
