"""In-tree build of libdsac_b200.so (nvcc, sm_100a only).

    python -m dsac_b200.build [--force] [--verbose]

The shared library is the product: the C ABI of include/dsac_b200.h plus the sm_100a
kernels.  It is built in-tree (dsac_b200/libdsac_b200.so, git-ignored) so that it
travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdsac_b200.so")
SOURCES = ["engine.cu", "host_util.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O2", "-DDSAC_BUILD=1", "-DK1_THREADS_DEF=" + os.environ.get("DSAC_K1_THREADS", "384"),
    *(["-DK1_MIN_BLOCKS=" + os.environ["DSAC_K1_MIN_BLOCKS"]] if "DSAC_K1_MIN_BLOCKS" in os.environ else []),
    *(["-DK4_THREADS_DEF=" + os.environ["DSAC_K4_THREADS"]] if "DSAC_K4_THREADS" in os.environ else []),
    *(["-DK4_MIN_BLOCKS=" + os.environ["DSAC_K4_MIN_BLOCKS"]] if "DSAC_K4_MIN_BLOCKS" in os.environ else []),
    *os.environ.get("DSAC_EXTRA_NVCC_FLAGS", "").split(),   # experiments (tools/sweep.py)
]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA engine cannot be built (there is no CPU fallback)")


SYNTH_LIB = os.path.join(HERE, "libdsac_synth.so")


def build_synth(force=False):
    """Host-only build (g++) of host_util.cpp: the synthetic-frame generator and stochasticSubSample without any CUDA
    dependency, so that CPU-only consumers (bench.py --impl reference) do not have to map the CUDA library."""
    src = os.path.join(CSRC, "host_util.cpp")
    deps = [src, os.path.join(CSRC, "pose_math.cuh"), os.path.join(HERE, "..", "include", "dsac_b200.h")]
    if not force and os.path.exists(SYNTH_LIB) and all(os.path.getmtime(SYNTH_LIB) >= os.path.getmtime(d) for d in deps):
        return SYNTH_LIB
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SYNTH_LIB, src])
    return SYNTH_LIB


def _deps():
    out = [os.path.join(HERE, "..", "include", "dsac_b200.h"), os.path.abspath(__file__)]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False, out=None, csrc=None):
    """out / csrc: build a variant of the library somewhere else (tools/sweep.py); the product is LIB from CSRC.
    DSAC_SKIP_BUILD=1: use the library as it is (a snapshot sent to the GPU box must not be rebuilt from sources that were
    being edited while the job waited in the queue)."""
    if out is None and not force and os.environ.get("DSAC_SKIP_BUILD") and os.path.exists(LIB):
        return LIB
    if out is None and not force and not needs_build():
        build_synth()
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or LIB] + \
        [os.path.join(csrc or CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libdsac_b200.so")
    if out is None:
        build_synth(force)
    return out or LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
