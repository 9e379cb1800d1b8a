"""CPU-only: the C-ABI shared library loads, exports every symbol include/tchgeo_cuda.h declares, and
the ctypes mirror of struct tchgeo_sampling_args has the C layout.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tchgeo_cuda.h")


@pytest.fixture(scope="module")
def native():
    from tch_geometric import _native
    return _native


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"TCHGEO_API[^;(]*?\b(tchgeo_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    assert len(syms) == 21
    for s in ("tchgeo_coo_to_csx", "tchgeo_neighbor_sampling", "tchgeo_neighbor_sampling_homogenous",
              "tchgeo_random_walk", "tchgeo_unique_relabel", "tchgeo_ind2ptr", "tchgeo_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(native):
    lib = ctypes.CDLL(native.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert sorted(native.EXPORTS) == declared_symbols()
    assert lib.tchgeo_abi_version() == native.ABI_VERSION == 4


def test_library_is_sm100a_native(native):
    out = subprocess.run(["cuobjdump", "--list-elf", native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_struct_layout_matches_c(native):
    fields = [f[0] for f in native.SamplingArgs._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "tchgeo_cuda.h"\nint main(){printf("%zu", sizeof(tchgeo_sampling_args));\n'
    for f in fields:
        prog += f'printf(" %zu", offsetof(tchgeo_sampling_args, {f}));\n'
    prog += "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(c, "w").write(prog)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        vals = [int(x) for x in subprocess.check_output([exe]).split()]
    assert vals[0] == ctypes.sizeof(native.SamplingArgs)
    for f, off in zip(fields, vals[1:]):
        assert getattr(native.SamplingArgs, f).offset == off, f


def test_host_only_entry_points_work_without_a_gpu(native):
    """capacity / workspace planning and argument validation are pure host code."""
    a = native.SamplingArgs()
    rel = np.zeros(1, dtype=np.int32)
    fan = np.array([15, 10, 5], dtype=np.int64)
    seeds = np.array([1024], dtype=np.int64)
    a.num_node_types, a.num_rels, a.num_hops, a.sampler_kind = 1, 1, 3, 0
    a.rel_src = a.rel_dst = rel.ctypes.data
    a.fanouts, a.seeds_per_batch, a.num_batches = fan.ctypes.data, seeds.ctypes.data, 256
    cn, ce = np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int64)
    assert native.lib.tchgeo_neighbor_sampling_capacity(ctypes.byref(a), cn.ctypes.data, ce.ctypes.data) == 0
    assert (cn[0], ce[0]) == (937984, 936960)  # SURVEY §8(a) config-2 bounds
    assert native.lib.tchgeo_neighbor_sampling_workspace_bytes(ctypes.byref(a)) > 0
    fan[1] = 1 << 20
    assert native.lib.tchgeo_neighbor_sampling_capacity(ctypes.byref(a), cn.ctypes.data, ce.ctypes.data) == native.ERR_BAD_ARG
    assert b"fanout" in native.lib.tchgeo_last_error()
    with pytest.raises(ValueError):
        native.check(native.ERR_BAD_ARG)
    # NULL / negative arguments are rejected before any CUDA call
    assert native.lib.tchgeo_ind2ptr(None, 5, 3, None, None) == native.ERR_BAD_ARG
    assert native.lib.tchgeo_random_walk(None, 1, None, None, 1, 1, 1.0, 1.0, 0, 0, None, None, None, None) == native.ERR_BAD_ARG
    assert native.lib.tchgeo_coo_to_csx(None, None, -1, 1, 1, 1, None, None, None, None, 0, None) == native.ERR_BAD_ARG
