"""In-tree build of liboisat.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m oisatgmi_b200.csrc.build [--force] [--verbose]

Objects and the shared library stay next to the sources (git-ignored, but they
travel to the GPU box with the gpurun snapshot).  `--fmad=false` is deliberate:
several kernels reproduce numpy/scipy results bit for bit and must not have
their multiplies and adds contracted; hot loops that may fuse call fma().
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["api.cu", "k0_geometry.cu", "k1_locate.cu", "k2_interp.cu", "k3_vertical.cu",
           "k4_accum.cu", "k5_oi.cu", "k7_reader.cu", "k8_output.cu", "k9_alive.cu", "k10_tables.cu", "k11_pwv.cu", "k12_flip.cu", "fused_amf.cu", "fused_split.cu", "delaunay.cpp"]
HEADERS = ["common.cuh", "vertical.cuh", "delaunay_swath.inl", "delaunay_seed.inl", "flip_rounds.h", os.path.join("..", "..", "include", "oisat.h")]
LIB = os.path.join(HERE, "liboisat.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
              "--expt-relaxed-constexpr", "--extended-lambda"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; liboisat cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(HERE, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(HERE, ".build_stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, os.path.splitext(src)[0] + ".o")
        assert obj.endswith(".o") and obj != os.path.join(HERE, src)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose or out.strip():
            sys.stderr.write("[%s]\n%s\n" % (src, out))
    if failed:
        raise RuntimeError("liboisat build failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + \
           ["-cudart", "static", "-Xcompiler", "-fPIC"]
    subprocess.check_call(link)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
