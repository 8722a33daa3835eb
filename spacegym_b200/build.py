"""Build the CUDA engine in-tree: spacegym_b200/libspacegym_b200.so (sm_100a only).

    python -m spacegym_b200.build [--force] [-v] [--single]

The engine is one source file (csrc/sg_engine.cu) whose kernels are templated over the env family; it is compiled
as six translation units in parallel (-DSG_PART=0..5: C ABI + Goal 2P | Goal 3P | Goal 4P | Kepler | DoNotCrash | Goal 1P)
and linked, ~1.5 min on 8 cores instead of 4.5 min for the single unit (`--single`, same code).
nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the tree.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", f) for f in ("sg_engine.cu", "sg_device.cuh", "sg_tail.cuh")] + [
    os.path.join(ROOT, "include", "spacegym_b200.h")]
OUT = os.path.join(HERE, "libspacegym_b200.so")
N_PARTS = 6

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
]  # no --split-compile: it halves the build time but the kernels come out 3-5 % slower


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in SRC)


def build(force=False, verbose=False, single=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-Xptxas", "-v"] if verbose else []
    if single:
        subprocess.check_call([nvcc] + NVCC_FLAGS + extra + ["-shared", SRC[0], "-o", OUT + ".tmp"])
        os.replace(OUT + ".tmp", OUT)
        return OUT
    tmp = tempfile.mkdtemp(prefix="sg_build_")
    try:
        objs = [os.path.join(tmp, f"part{k}.o") for k in range(N_PARTS)]
        procs = [subprocess.Popen([nvcc] + NVCC_FLAGS + extra + [f"-DSG_PART={k}", "-c", SRC[0], "-o", objs[k]])
                 for k in range(N_PARTS)]
        rcs = [p.wait() for p in procs]
        if any(rcs):
            raise subprocess.CalledProcessError(max(rcs), "nvcc (one of the SG_PART translation units)")
        subprocess.check_call([nvcc] + NVCC_FLAGS + ["-shared"] + objs + ["-o", OUT + ".tmp"])
        os.replace(OUT + ".tmp", OUT)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


CHECK_OUT = os.path.join(HERE, "libspacegym_b200_check.so")


def build_check(force=False):
    """The sanitizer-substitute build (tests/test_gpu_queue_check.py): the same engine compiled with
    -DSG_CHECK_QUEUES (device-side invariant counters on the work queues of the split step), BASELINE
    instantiation only (GoalContinuous3P-v0; one translation unit, ~25 s)."""
    if not force and os.path.exists(CHECK_OUT) and all(os.path.getmtime(s) <= os.path.getmtime(CHECK_OUT) for s in SRC):
        return CHECK_OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc] + NVCC_FLAGS + ["-DSG_CHECK_QUEUES", "-DSG_DEV_ONLY_GOAL3P", "-shared", SRC[0], "-o",
                                                 CHECK_OUT + ".tmp"])
    os.replace(CHECK_OUT + ".tmp", CHECK_OUT)
    return CHECK_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, single="--single" in sys.argv))
    print(build_check(force="--force" in sys.argv))
