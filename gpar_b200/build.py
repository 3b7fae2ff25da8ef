"""Builds gpar_b200/libgpar_b200.so (sm_100a only) with nvcc.  In-tree, so the
shared object travels with the repo snapshot to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgpar_b200.so")
DEBUG_LIB = os.path.join(HERE, "libgpar_b200_debug.so")  # probes only (include/gpar_b200_debug.h), not the product
SOURCES = ["capi.cu", "gram.cu", "potrf.cu", "solve.cu"]
DEBUG_SOURCES = ["debug.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "--extended-lambda", "-Xcompiler", "-fPIC",
]


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(DEBUG_LIB):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(DEBUG_LIB))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", h) for h in ("gpar_b200.h", "gpar_b200_debug.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), suffix=""):
    """defines/suffix: kernel experiments (e.g. defines=["-DGPAR_STAGES=5"], suffix="_s5" builds
    libgpar_b200_s5.so next to the product library; select it with GPAR_B200_LIB)."""
    if suffix:
        return _build(verbose, list(defines), LIB.replace(".so", suffix + ".so"), "build" + suffix)
    if not force and not _stale():
        return LIB
    _build(verbose, [], DEBUG_LIB, "build_debug", DEBUG_SOURCES)
    return _build(verbose, [], LIB, "build")


def _build(verbose, defines, LIB, bdir, SOURCES=SOURCES):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, bdir), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, bdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
