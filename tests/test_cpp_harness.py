"""A C++ program (tests/cpp/abi_harness.cpp) drives the C ABI with no Python or torch in the loop: it must compile
against include/tchgeo_cuda.h, link libtchgeo_cuda.so (CPU check) and, on the GPU box, reproduce karate's CSC, a
full-neighbourhood 2-hop sample and its relabel map through the graph / plan handles."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "abi_harness.cpp")
LIBDIR = os.path.join(ROOT, "tch-geometric_b200", "tch_geometric")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build(tmp):
    exe = os.path.join(tmp, "abi_harness")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"), SRC,
           "-o", exe, "-L", LIBDIR, "-l:libtchgeo_cuda.so", "-L", os.path.join(CUDA, "lib64"), "-lcudart",
           "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + os.path.join(CUDA, "lib64")]
    subprocess.check_call(cmd)
    return exe


def test_harness_compiles_and_links(tmp_path):
    exe = build(str(tmp_path))
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 64 and "usage" in out.stderr      # ran far enough to resolve every symbol


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["karate", "fakedataset"])
def test_harness_runs_the_path_without_python(tmp_path, fixture):
    exe = build(str(tmp_path))
    d = np.load(os.path.join(ROOT, "tests", "golden", fixture + ".npz"))
    edges = os.path.join(str(tmp_path), "edges.bin")
    np.ascontiguousarray(d["edge_index"], dtype=np.int64).tofile(edges)
    out = subprocess.run([exe, edges, str(int(d["num_nodes"]))], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "abi_harness ok" in out.stdout
