"""Builds the native twin of the CPU oracle (oracle/native/dcora_oracle.cpp) into
oracle/_build/liboracle_native.so with plain g++ -O3 -march=native.  Run by
__graft_entry__.build(); TEST INFRASTRUCTURE (the product never loads it)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "..", "_build")
LIB = os.path.join(OUT_DIR, "liboracle_native.so")
SRC = os.path.join(HERE, "dcora_oracle.cpp")


def build(verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        # -march=native would tie the library to this container's CPU; the GPU box may differ,
        # so the baseline is built for x86-64-v3 (AVX2 + FMA), which every candidate host has
        cmd = [os.environ.get("CXX", "g++"), "-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC",
               "-shared", SRC, "-o", LIB]
        subprocess.run(cmd, check=True)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(verbose=True)
