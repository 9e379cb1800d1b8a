import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tch-geometric_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def karate():
    d = np.load(os.path.join(GOLDEN, "karate.npz"))
    return d["edge_index"], int(d["num_nodes"])


@pytest.fixture(scope="session")
def fakedataset():
    d = np.load(os.path.join(GOLDEN, "fakedataset.npz"))
    return d["edge_index"], int(d["num_nodes"])


@pytest.fixture(scope="session")
def fakehetero():
    """-> (node_counts{type}, edge_index{(src, rel, dst)}); key scheme of src/data/io.rs:28-51"""
    d = np.load(os.path.join(GOLDEN, "fakeheterodataset.npz"))
    counts, edges = {}, {}
    for k in sorted(d.files):
        if k.startswith("num_nodes_"):
            counts[k[len("num_nodes_"):]] = int(d[k])
        elif k.startswith("edge_"):
            s, r, t = k.split("_")[1].split("-")
            edges[(s, r, t)] = d[k]
    return counts, edges
