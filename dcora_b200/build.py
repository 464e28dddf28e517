"""In-tree build of libdcora_b200.so for sm_100a (nvcc, no torch dependency)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libdcora_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["layout.cpp", "chol.cpp", "trsv_host.cpp", "selfcheck.cpp", "kernels.cu", "sptrsv.cu", "problem.cu",
           "solver.cu", "rtr_cluster.cu", "certify.cu", "rbcd.cu", "rounding.cu"]
# (source, extra flags, object tag): the per-agent RTR cluster kernel, one translation unit per rank
VARIANTS = [("rtr_cluster_inst.cu", [f"-DDCORA_CLUSTER_R={r}"], f"r{r}") for r in range(2, 9)]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
         "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp,-O3", "--expt-relaxed-constexpr", "-diag-suppress", "128"]


def _deps_hash(src):
    h = hashlib.sha1()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cuh", ".h")) or name == src:
            with open(os.path.join(CSRC, name), "rb") as fh:
                h.update(fh.read())
    with open(os.path.join(HERE, "..", "include", "dcora_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(job):
    src, extra, tag = job if isinstance(job, tuple) else (job, [], "")
    obj = os.path.join(BUILD, src + (("." + tag) if tag else "") + ".o")
    stamp = obj + ".sha1"
    want = _deps_hash(src) + "".join(extra)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return obj, False
    cmd = [NVCC, *FLAGS, *extra, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
    subprocess.run(cmd, check=True)
    with open(stamp, "w") as fh:
        fh.write(want)
    return obj, True


def build(verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    srcs = [v for v in VARIANTS if os.path.exists(os.path.join(CSRC, v[0]))] + srcs   # slowest first
    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    objs = [o for o, _ in results]
    if any(ch for _, ch in results) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fopenmp",
               "-lgomp", "-lcudart"]
        subprocess.run(cmd, check=True)
    if verbose:
        print("built", LIB)
    return LIB


EXAMPLES = os.path.join(HERE, "..", "examples")
EXAMPLE_BIN = os.path.join(EXAMPLES, "bin")
CXX = os.environ.get("CXX", "g++")


def build_examples(verbose=False):
    """C++ host-side programs written against include/dcora_b200/DCORA.hpp
    (plain g++; they only link the C-ABI library)."""
    os.makedirs(EXAMPLE_BIN, exist_ok=True)
    out = []
    inc = os.path.join(HERE, "..", "include")
    cpp_tests = os.path.join(HERE, "..", "tests", "cpp")
    sources = [os.path.join(EXAMPLES, f) for f in sorted(os.listdir(EXAMPLES)) if f.endswith(".cpp")]
    if os.path.isdir(cpp_tests):
        sources += [os.path.join(cpp_tests, f) for f in sorted(os.listdir(cpp_tests))
                    if f.endswith(".cpp")]
    for src in sources:
        name = os.path.basename(src)
        exe = os.path.join(EXAMPLE_BIN, name[:-4])
        deps = [src, os.path.join(inc, "dcora_b200.h"), os.path.join(inc, "dcora_b200", "DCORA.hpp"),
                os.path.join(inc, "dcora_b200", "DCORA_dist.hpp")]
        if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
            cmd = [CXX, "-std=c++17", "-O2", "-Wall", "-I", inc, src, "-L", HERE, "-ldcora_b200",
                   "-Wl,-rpath,$ORIGIN/../../dcora_b200", "-Wl,-rpath,/usr/local/cuda/lib64",
                   "-o", exe]
            with open(src) as fh:
                if "DCORA_dist.hpp" in fh.read():
                    # the multi-process host links the CUDA runtime and NCCL directly
                    cmd += ["-I", "/usr/local/cuda/include", "-L", "/usr/local/cuda/lib64",
                            "-lcudart", "-lnccl"]
            subprocess.run(cmd, check=True)
        out.append(exe)
        if verbose:
            print("built", exe)
    return out


if __name__ == "__main__":
    build(verbose=True)
    build_examples(verbose=True)
