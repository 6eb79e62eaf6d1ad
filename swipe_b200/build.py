"""Builds the in-tree native code: the sm_100a CUDA library behind include/swipe_b200.h and,
for tests / bench baselines only, the CPU oracle under oracle/."""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "swipe_b200", "csrc")
LIB = os.path.join(CSRC, "libswipe_b200.so")
SOURCES = ["swb_api.cu", "swb_blastdb.cu", "swb_align.cu", "swb_scoring.cu", "swb_text.cu", "swb_ubench.cu"]
DEPS = ["swb_api.cu", "sw_kernels.cuh", "swb_blastdb.cu", "swb_blastdb.h", "swb_align.cu", "swb_scoring.cu", "swb_text.cu", "swb_ubench.cu", "swb_tables.inc", os.path.join(ROOT, "include", "swipe_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    for d in deps:
        p = d if os.path.isabs(d) else os.path.join(CSRC, d)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build_lib(force=False, verbose=False):
    """Compile swipe_b200/csrc/libswipe_b200.so for sm_100a (cross-compiles without a GPU)."""
    if not force and not _stale(LIB, DEPS):
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("nvcc not found and %s is not built" % LIB)
    # several ranks of one job may get here at once: one builds (under a file lock, into a temporary
    # file that is renamed when complete), the others wait and find the library fresh
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or _stale(LIB, DEPS):
                tmp = LIB + ".tmp.%d" % os.getpid()
                cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
                subprocess.run(cmd, cwd=CSRC, check=True)
                os.replace(tmp, LIB)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


CLI = os.path.join(CSRC, "swipe-b200")


def build_cli(force=False):
    """The command-line front end (host C++ over the C ABI): swipe_b200/csrc/swipe-b200."""
    src = os.path.join(CSRC, "swipe_main.cpp")
    if not force and not _stale(CLI, ["swipe_main.cpp", LIB, os.path.join(ROOT, "include", "swipe_b200.h")]):
        return CLI
    cxx = shutil.which("g++")
    if cxx is None:
        if os.path.exists(CLI):
            return CLI
        raise RuntimeError("g++ not found and %s is not built" % CLI)
    subprocess.run([cxx, "-O2", "-std=c++17", "-Wall", "-o", CLI, src, "-L" + CSRC, "-lswipe_b200",
                    "-Wl,-rpath,$ORIGIN", "-lpthread"], check=True)
    return CLI


def build_oracle():
    """TEST INFRASTRUCTURE: the plain-C oracle and, where /root/reference exists, the unmodified
    reference kernels under oracle/_ref (see oracle/Makefile)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)


if __name__ == "__main__":
    import sys
    build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_cli(force="--force" in sys.argv)
    build_oracle()
    print(LIB)
