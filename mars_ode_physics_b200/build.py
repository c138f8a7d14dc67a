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
SOURCES = ["b2p_assemble.cu", "b2p_sor.cu", "b2p_sor_tmem.cu", "b2p_dense.cu", "b2p_integrate.cu", "b2p_arena.cu", "b2p_collide.cu"]
HEADERS = ["b2p_dev.h", "b2p_math.cuh", "b2p_kern.cuh", "b2p_rows.cuh", "b2p_universal.h",
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
    """Compile every .cu to an object (in parallel, only the stale ones) and link the shared libraries."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    targets = [("libmars_ode_b200.so", [], "")]
    if strict:
        targets.append(("libmars_ode_b200_strict.so", ["-fmad=false", "-DB2P_STRICT"], "_strict"))
    jobs, links = [], []
    for name, extra, tag in targets:
        objs = []
        for src in SOURCES:
            obj = os.path.join(objdir, src.replace(".cu", tag + ".o"))
            objs.append(obj)
            if force or _stale(obj, [os.path.join(CSRC, src)] + hdrs):
                cmd = [_nvcc(), *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", *extra, "-c", "-o", obj,
                       os.path.join(CSRC, src)]
                if verbose:
                    cmd[1:1] = ["-Xptxas", "-v"]
                jobs.append(cmd)
        links.append((os.path.join(LIBDIR, name), objs))

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(run, jobs))
    built = []
    for out, objs in links:
        if force or _stale(out, objs):
            run([_nvcc(), *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", out, *objs])
        built.append(out)
    return built


if __name__ == "__main__":
    for p in build(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(p)
