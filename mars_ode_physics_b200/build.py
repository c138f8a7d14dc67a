"""Build the sm_100a shared library that implements include/mars_ode_b200.h.

The product has exactly one native artefact: ``lib/libmars_ode_b200.so`` (CUDA kernels + C ABI,
static cudart).  It is built in-tree with nvcc so that it travels to the GPU box with the
snapshot.  ``strict=True`` additionally builds ``libmars_ode_b200_strict.so`` with
``-fmad=false`` -- same code without FMA contraction, used by the parity tests to compare the
device arithmetic term by term with the float build of the oracle.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SOURCES = ["b2p_assemble.cu", "b2p_sor.cu", "b2p_sor_tmem.cu", "b2p_integrate.cu", "b2p_arena.cu", "b2p_collide.cu"]
HEADERS = ["b2p_dev.h", "b2p_math.cuh", "b2p_kern.cuh", "b2p_universal.h",
           os.path.join("..", "..", "include", "mars_ode_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return exe


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(strict=True, force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    targets = [("libmars_ode_b200.so", [])]
    if strict:
        targets.append(("libmars_ode_b200_strict.so", ["-fmad=false", "-DB2P_STRICT"]))
    built = []
    for name, extra in targets:
        out = os.path.join(LIBDIR, name)
        if force or _stale(out, deps):
            cmd = [_nvcc(), *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                   *extra, "-o", out, *srcs]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
                print(" ".join(cmd), file=sys.stderr)
            subprocess.run(cmd, check=True)
        built.append(out)
    return built


if __name__ == "__main__":
    for p in build(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(p)
