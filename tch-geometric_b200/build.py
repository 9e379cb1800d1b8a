"""Build libtchgeo_cuda.so (sm_100a) in-tree with nvcc.  No torch, no JIT cache: the .so sits next to
the Python host module so it travels with the source tree.

    python tch-geometric_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "tch_geometric", "libtchgeo_cuda.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["capi.cu", "csx_build.cu", "csx_transform.cu", "gather.cu", "negative_sampling.cu", "neighbor_sampling.cu", "partitioned.cu", "partitioned_fixed.cu", "random_walk.cu", "relabel.cu", "transport.cu", "host_unpack.cpp"]
DEPS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "graph.cuh"), os.path.join(HERE, "..", "include", "tchgeo_cuda.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-ccbin", "/usr/bin/g++",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + DEPS):
            if src.endswith(".cpp"):   # host-only code: the host compiler directly
                cmd = ["/usr/bin/g++", "-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-c", s, "-o", o]
            else:
                cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- compiler output, {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-Wno-deprecated-gpu-targets", "-o", OUT] + objs + ["-ccbin", "/usr/bin/g++", "-Xlinker", "--exclude-libs,ALL", "-lpthread"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
