"""Builds cdae_b200/libcdae_b200.so (the C-ABI library of include/cdae_b200.h) in-tree with nvcc
for sm_100a.  nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU
box with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libcdae_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "api.cu")]
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [
    os.path.join(ROOT, "include", "cdae_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def up_to_date():
    if not os.path.exists(SO):
        return False
    t = os.path.getmtime(SO)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False, out=None):
    """out: build a variant library elsewhere (A/B runs: CDAE_NVCC_FLAGS='-DDECODE_UNR=2' and
    CDAE_B200_LIB=<that path> at run time)."""
    if out:
        nvcc = os.environ.get("NVCC", "nvcc")
        extra = os.environ.get("CDAE_NVCC_FLAGS", "").split()
        subprocess.check_call([nvcc] + NVCC_FLAGS + extra + SOURCES + ["-o", out, "-ldl"])
        return out
    if not force and up_to_date():
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("CDAE_NVCC_FLAGS", "").split()     # e.g. -DDECODE_MIN_BLOCKS=2 for A/B runs
    cmd = [nvcc] + NVCC_FLAGS + extra + SOURCES + ["-o", SO, "-ldl"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
