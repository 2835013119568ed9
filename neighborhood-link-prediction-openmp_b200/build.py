"""Build the CUDA shared library (sm_100a only) in-tree with nvcc."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libnlp_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "nlp_b200.cu")]
HEADERS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + [os.path.join(HERE, "..", "include", "nlp_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS if os.path.exists(f))


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libnlp_b200.so (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
