"""Builds htscodecs_b200/libhtscodecs_b200.so in-tree with nvcc for sm_100a (no torch involved)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhtscodecs_b200.so")
SOURCES = ["api.cu", "decode.cu", "encode.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall", "--fmad=false",
    "-Xptxas", "-v", "-shared", "-cudart", "shared",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for fn in os.listdir(root):
            if os.path.getmtime(os.path.join(root, fn)) > t:
                return True
    return False


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print("built", LIB)
