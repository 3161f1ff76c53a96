#!/usr/bin/env python
"""Build libkgrec_b200.so (the C-ABI CUDA library) for sm_100a with nvcc.

    python joint-kg-recommender_b200/csrc/build.py [--force]

Each .cu is compiled to an object in parallel, then linked into
joint-kg-recommender_b200/lib/libkgrec_b200.so (git-ignored; travels to the GPU box).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(os.path.dirname(HERE), "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB = os.path.join(LIB_DIR, "libkgrec_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-diag-suppress", "20281", "--expt-relaxed-constexpr"]


def digest(paths):
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def build(force=False, verbose=True):
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cu"))
    hdrs = sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h")))
    hdrs.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "kgrec_b200.h"))
    hdr_digest = digest(hdrs)
    jobs = []
    for src in srcs:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".sha"
        want = digest([src]) + hdr_digest
        have = open(stamp).read() if os.path.exists(stamp) and os.path.exists(obj) else ""
        if force or have != want:
            jobs.append((src, obj, stamp, want))

    def compile_one(job):
        src, obj, stamp, want = job
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        with open(stamp, "w") as f:
            f.write(want)
        return src

    if jobs:
        if verbose:
            print("[kgrec build] compiling", ", ".join(os.path.basename(j[0]) for j in jobs), flush=True)
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if jobs or not os.path.exists(LIB):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("[kgrec build] linked", LIB, flush=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
