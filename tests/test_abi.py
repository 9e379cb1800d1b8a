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
    assert len(syms) == 57
    for s in ("tchgeo_coo_to_csx", "tchgeo_neighbor_sampling", "tchgeo_neighbor_sampling_homogenous",
              "tchgeo_random_walk", "tchgeo_unique_relabel", "tchgeo_ind2ptr", "tchgeo_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(native):
    lib = ctypes.CDLL(native.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert sorted(native.EXPORTS) == declared_symbols()
    assert lib.tchgeo_abi_version() == native.ABI_VERSION == 6


def test_library_is_sm100a_native(native):
    out = subprocess.run(["cuobjdump", "--list-elf", native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


@pytest.mark.parametrize("cname,pyname", [("tchgeo_sampling_args", "SamplingArgs"), ("tchgeo_negative_args", "NegativeArgs")])
def test_struct_layout_matches_c(native, cname, pyname):
    cls = getattr(native, pyname)
    fields = [f[0] for f in cls._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "tchgeo_cuda.h"\nint main(){printf("%%zu", sizeof(%s));\n' % cname
    for f in fields:
        prog += f'printf(" %zu", offsetof({cname}, {f}));\n'
    prog += "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(c, "w").write(prog)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        vals = [int(x) for x in subprocess.check_output([exe]).split()]
    assert vals[0] == ctypes.sizeof(cls)
    for f, off in zip(fields, vals[1:]):
        assert getattr(cls, f).offset == off, f


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
    # negative sampling: capacity planning is host code too
    n = native.NegativeArgs()
    src, dst = np.array([0, 0], dtype=np.int32), np.array([0, 1], dtype=np.int32)
    dummy = np.zeros(2, dtype=np.uint64) + 8  # non-NULL placeholders (never dereferenced by the planner)
    nrows, ncount, ninp = np.array([5, 5], dtype=np.int64), np.array([5, 9], dtype=np.int64), np.array([4, 2], dtype=np.int64)
    n.num_node_types, n.num_rels, n.num_neg, n.try_count = 2, 2, 3, 5
    n.rel_src, n.rel_dst = src.ctypes.data, dst.ctypes.data
    n.row_ptrs = n.col_indices = n.inputs = dummy.ctypes.data
    n.num_rows, n.node_count, n.num_inputs = nrows.ctypes.data, ncount.ctypes.data, ninp.ctypes.data
    cn, ce = np.zeros(2, dtype=np.int64), np.zeros(2, dtype=np.int64)
    assert native.lib.tchgeo_negative_sampling_capacity(ctypes.byref(n), cn.ctypes.data, ce.ctypes.data) == 0
    assert cn.tolist() == [4 + 12, 2 + 12] and ce.tolist() == [12, 12]
