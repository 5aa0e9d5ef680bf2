"""Golden vectors for barycentre grid subsampling, made by the reference's OWN CPU code
(oracle/_ref/libgridsub_ref.so = grid_subsampling.cpp + cloud.cpp compiled unmodified, see build_ref.build_gridsub).
The reference emits cells in unordered_map order; fixtures store them re-ordered by ascending cell index (keys recomputed
by the C oracle, whose barycentres are matched row for row against the reference's first).

    python oracle/make_golden_gridsub.py      # needs /root/reference; writes tests/golden/gridsub.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, cpu_oracle  # noqa: E402


def scan(n, seed):
    """Room-like scan: a few planes in a 4 x 3 x 2.5 m box + 5 mm noise (BASELINE configs[3] shape)."""
    g = np.random.default_rng(seed)
    pts = []
    for _ in range(6):
        a = g.integers(0, 3); v = g.uniform(0, (4, 3, 2.5)[a])
        p = g.uniform(0, 1, (n // 6, 3)) * np.array([4, 3, 2.5]); p[:, a] = v
        pts.append(p)
    return (np.concatenate(pts) + g.normal(0, 0.005, (len(pts) * (n // 6), 3))).astype(np.float32)


def canon(ref_out, oracle_out):
    """Re-order the reference's rows into ascending-cell order by matching barycentres bit for bit."""
    rp, rf, rl = ref_out
    op, of, ol, keys = oracle_out
    assert rp.shape == op.shape, (rp.shape, op.shape)
    view = lambda a: np.ascontiguousarray(a).view([("x", "<u4"), ("y", "<u4"), ("z", "<u4")]).ravel()
    ro, oo = np.argsort(view(rp), order=("x", "y", "z")), np.argsort(view(op), order=("x", "y", "z"))
    perm = np.empty(len(ro), np.int64); perm[oo] = ro          # oracle row i <-> reference row perm[i]
    return rp[perm], (None if rf is None else rf[perm]), (None if rl is None else rl[perm]), keys


def main():
    out = {}
    cases = [("room_6k_dl10", scan(6000, 1), 0.10, 4, 1), ("room_30k_dl04", scan(30000, 2), 0.04, 3, 2),
             ("cloud_1k_dl02", np.random.default_rng(3).standard_normal((1024, 3)).astype(np.float32) * 0.3, 0.02, 0, 0),
             ("dup_dl05", np.repeat(np.random.default_rng(4).uniform(-1, 1, (200, 3)).astype(np.float32), 5, 0), 0.05, 2, 1)]
    for name, pts, dl, fdim, ldim in cases:
        g = np.random.default_rng(len(name))
        feats = g.standard_normal((len(pts), fdim)).astype(np.float32) if fdim else None
        # labels without count ties inside a cell cannot be guaranteed, so ties are checked separately in the tests;
        # the fixture keeps only cells whose majority is unique
        labels = g.integers(0, 3, (len(pts), ldim)).astype(np.int32) if ldim else None
        ref = build_ref.ref_grid_subsampling(pts, feats, labels, dl)
        orc = cpu_oracle.grid_subsample(pts, feats, labels, dl)
        rp, rf, rl, keys = canon(ref, orc)
        assert np.array_equal(rp, orc[0]), name
        if fdim:
            assert np.array_equal(rf, orc[1]), name
        out[name + "_points"] = pts; out[name + "_dl"] = np.float32(dl)
        if fdim: out[name + "_features"] = feats
        if ldim: out[name + "_labels"] = labels
        out[name + "_sub_points"] = rp; out[name + "_keys"] = keys
        if fdim: out[name + "_sub_features"] = rf
        if ldim:
            out[name + "_sub_labels"] = rl
            print(name, "label rows differing from the smallest-label tie rule:", int((rl != orc[2]).any(1).sum()), "of", len(rl))
        print(name, "N", len(pts), "-> M", len(rp))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gridsub.npz"), **out)


if __name__ == "__main__":
    main()
