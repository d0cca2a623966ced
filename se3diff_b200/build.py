"""Builds libse3diff_b200.so (sm_100a) in-tree with nvcc.  No torch, no pybind: the library is a
plain C ABI (include/se3diff_b200.h) loaded with ctypes by se3diff_b200._lib."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libse3diff_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# translation unit -> extra flags.  The SDE-algebra kernels are compiled without FMA contraction so
# that each fp32 expression rounds exactly like the reference's torch expression.
SOURCES = {
    "api.cu": [],
    "so3_kernels.cu": ["-fmad=false"],
    "frame_kernels.cu": ["-fmad=false"],
    "r3_kernels.cu": ["-fmad=false"],
    "igso3_kernels.cu": ["-fmad=false"],
    "ipa_simt.cu": [],
    "ipa_bwd.cu": [],
    "tc_selftest.cu": [],
    "ipa_tc.cu": [],
    "ipa_tc_pp.cu": [],
    "pair_pack.cu": [],
    "pair_precompute.cu": [],
    "fused_rows.cu": [],
    "observables.cu": [],
    "backbone.cu": [],
}


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(repr(sorted(SOURCES.items())).encode())
    return h.hexdigest()


def stale() -> bool:
    """True when the library on disk was not built from the sources on disk (or there is no record of what it was built from)."""
    stamp = os.path.join(OUT_DIR, "build.sha256")
    return not (os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == _digest())


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "build.sha256")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
