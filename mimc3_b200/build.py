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
SOURCES = ["api.cu", "match.cu", "match2.cu", "sat.cu", "cp.cu", "conv2.cu", "post.cu", "comm.cu", "probe.cu"]
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


def build_variant(name: str, extra_flags) -> str:
    """Development aid: the library compiled with extra nvcc flags (e.g. -DMIMC3CU_PROFILE) into
    libmimc3cu_<name>.so next to the product library; MIMC3CU_LIB=<path> makes lib.py load it."""
    return build(force=True, out=os.path.join(HERE, f"libmimc3cu_{name}.so"), objdir=os.path.join(HERE, "build", name),
                 extra=list(extra_flags), dropin=False)


def build(force: bool = False, verbose: bool = False, out: str = LIB, objdir: str = "", extra=(), dropin: bool = True) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    objdir = objdir or os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *extra, *os.environ.get("MIMC3CU_NVCC_EXTRA", "").split(), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(log)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    # linked under a temporary name and renamed: a repository snapshot (gpurun) never sees a half-written library
    cmd = [NVCC, "-shared", "-o", out + ".tmp", *objs, "-ccbin", HOST_CXX, "-Xcompiler", "-pthread", "-lcudart", "-ldl"]
    subprocess.run(cmd, check=True)
    os.replace(out + ".tmp", out)
    if dropin:
        build_dropin()
    return out


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
    if "--variant" in sys.argv:      # python -m mimc3_b200.build --variant prof -DMIMC3CU_PROFILE
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
