"""Regenerates the committed fixtures under tests/golden/ from the reference's own test fixtures
(/root/reference/tests/*.npz, read-only; only present in the build container).

Only the integer graph structure the sampling path consumes is kept (edge_index arrays and node
counts); node features/labels are dropped.  Key naming follows src/data/io.rs:28-51.

    python tests/golden/make_golden.py
"""
import os

import numpy as np

REF = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    d = np.load(os.path.join(REF, "karate.npz"))
    np.savez_compressed(os.path.join(HERE, "karate.npz"), edge_index=d["edge_index"].astype(np.int64),
                        num_nodes=np.int64(d["x"].shape[0]))
    d = np.load(os.path.join(REF, "fakedataset.npz"))
    np.savez_compressed(os.path.join(HERE, "fakedataset.npz"), edge_index=d["edge_index"].astype(np.int64),
                        num_nodes=np.int64(d["x"].shape[0]))
    d = np.load(os.path.join(REF, "fakeheterodataset.npz"))
    out = {}
    for k in d.files:
        if k.startswith("node_") and k.endswith("_x"):
            out["num_nodes_" + k.split("_")[1]] = np.int64(d[k].shape[0])
        elif k.startswith("edge_"):
            out[k] = d[k].astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "fakeheterodataset.npz"), **out)
    for f in ("karate", "fakedataset", "fakeheterodataset"):
        g = np.load(os.path.join(HERE, f + ".npz"))
        print(f, {k: (g[k].shape, str(g[k].dtype)) for k in g.files})


if __name__ == "__main__":
    main()
