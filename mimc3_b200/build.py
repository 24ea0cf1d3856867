"""Builds mimc3_b200/libmimc3cu.so in-tree with nvcc for sm_100a (no JIT cache: the .so
travels to the GPU box with the repository snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmimc3cu.so")
DROPIN = os.path.join(HERE, "libmimc3cu_dropin.a")
SOURCES = ["api.cu", "match.cu", "match2.cu", "sat.cu", "cp.cu", "conv2.cu", "post.cu", "probe.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                      # no FMA contraction: results must be bit-equal to the CPU reference
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O3,-pthread",
    "-ccbin", HOST_CXX,
]


def _stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(DROPIN):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mimc3cu.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *os.environ.get("MIMC3CU_NVCC_EXTRA", "").split(), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-ccbin", HOST_CXX, "-Xcompiler", "-pthread", "-lcudart"]
    subprocess.run(cmd, check=True)
    build_dropin()
    return LIB


def build_dropin() -> str:
    """libmimc3cu_dropin.a: the reference's MIMC_module.h entry points over the C ABI (host C++).
    Its undefined symbols (the driver's globals and GMA_*_create) resolve when the reference
    driver is linked against it -- see INTEGRATION.md."""
    obj = os.path.join(HERE, "build", "dropin.o")
    subprocess.run([HOST_CXX, "-std=c++17", "-O2", "-fPIC", "-pthread", "-c", os.path.join(CSRC, "dropin.cpp"), "-o", obj], check=True)
    if os.path.exists(DROPIN):
        os.remove(DROPIN)
    subprocess.run(["ar", "rcs", DROPIN, obj], check=True)
    return DROPIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
