"""Pin the CPU oracle (oracle/ri_oracle.c) against the golden vectors produced by the reference's own CUDA
kernels on a B200 (oracle/make_golden.py -> tests/golden/*.npz).  Runs on CPU.

Bit-exact: KNN distances and indices (full order), cube voxel indices/counts, devox corner indices and weights,
cube devox outputs.  1e-5: PPF, voxel means, gradients.  Spherical cells: the oracle uses glibc acosf/atanf, the
reference CUDA libdevice; any disagreement must be a point sitting on a cell boundary (checked explicitly)."""
import numpy as np
import pytest

from _util import TOL, load_golden, rel_err, scaled_err


def test_knn(oracle, golden_dir):
    g = load_golden(golden_dir, "knn.npz")
    d1, d2, i1, i2 = oracle.knn(g["xyz1"], g["xyz2"], int(g["k"]))
    assert np.array_equal(i1, g["idx1"]) and np.array_equal(i2, g["idx2"])
    assert np.array_equal(d1, g["dist1"]) and np.array_equal(d2, g["dist2"])
    assert (g["idx1"][0, :, 100] == 0).all() and (g["dist1"][0, :, 100] == 10000.0).all()   # far query: all sentinels
    g1, g2 = oracle.knn_grad(g["xyz1"], g["xyz2"], g["graddist1"], g["graddist2"], g["idx1"], g["idx2"])
    assert scaled_err(g1, g["gradxyz1"]) <= TOL and scaled_err(g2, g["gradxyz2"]) <= TOL


def test_knn_edge_cases(oracle, golden_dir):
    e = load_golden(golden_dir, "knn_edge.npz")
    d1, d2, i1, i2 = oracle.knn(e["xq"], e["xs"], 8)
    assert np.array_equal(i1, e["idx1"]) and np.array_equal(d1, e["dist1"])
    assert np.array_equal(i2, e["idx2"]) and np.array_equal(d2, e["dist2"])
    assert (e["dist1"][:, 5:, :] == 10000.0).all() and (e["idx1"][:, 5:, :] == 0).all()     # m = 5 < k = 8
    d, i = oracle.knn_one(e["xc"], e["xc"], 20)
    assert np.array_equal(i, e["self_idx"]) and np.array_equal(d, e["self_dist"])
    d1, d2, i1, i2 = oracle.knn(e["x5"], e["y5"], 4)
    assert np.array_equal(i1, e["c5_idx1"]) and np.array_equal(d1, e["c5_dist1"])
    assert np.array_equal(i2, e["c5_idx2"]) and np.array_equal(d2, e["c5_dist2"])


def test_ppf(oracle, golden_dir):
    g = load_golden(golden_dir, "ppf.npz")
    o = oracle.ppf_backend(g["coords"], g["center"], g["normals"], g["center_normal"])
    ref = g["feat"]
    assert (ref[:, :, 8:24] == 0).all()                        # zero-length normals -> all-zero rows
    assert np.array_equal(o[:, 3], ref[:, 3])                  # ||d|| (and the 1e-20 clamp) bit-exact
    assert np.max(np.abs(o - ref)) <= 1e-6                      # angles: libm vs device f64 acos, <= 1 ulp(pi)
    assert rel_err(o[:, :3][ref[:, :3] > 1e-3], ref[:, :3][ref[:, :3] > 1e-3]) <= TOL


@pytest.mark.parametrize("r", [4, 8, 16, 32])
def test_spherical(oracle, golden_dir, r):
    g = load_golden(golden_dir, "spherical.npz")
    coords, feat = g[f"r{r}_coords"], g[f"r{r}_feat"]
    ind, cnt = oracle.sph_grid_stats(coords, r)
    ref_ind = g[f"r{r}_ind"]
    assert np.array_equal(ind == -1, ref_ind == -1)            # defined / undefined agrees everywhere
    bad = np.argwhere(ind != ref_ind)
    assert len(bad) <= 2, "oracle disagrees with the reference kernel on %d spherical cells" % len(bad)
    if len(bad):
        gc = oracle.sph_grid_cont(coords, r)
        for b, i in bad:
            frac = np.abs(gc[b, i] - np.rint(gc[b, i]))
            assert frac.min() < 1e-4, "a non-boundary point landed in a different cell"
    else:
        assert np.array_equal(cnt, g[f"r{r}_cnt"])
        assert scaled_err(oracle.scatter_mean(feat, ind, cnt, r), g[f"r{r}_out"]) <= TOL
    # the rest is checked on the REFERENCE's indices so a boundary flip cannot hide an error
    out = oracle.scatter_mean(feat, ref_ind, g[f"r{r}_cnt"], r)
    assert scaled_err(out, g[f"r{r}_out"]) <= TOL
    assert rel_err(oracle.avg_voxelize_grad(g[f"r{r}_gy"], ref_ind, g[f"r{r}_cnt"]), g[f"r{r}_gx"]) <= TOL
    o, di, dw = oracle.spherical_trilinear_devoxelize(coords, g[f"r{r}_grid"], ref_ind, r)
    assert np.array_equal(di, g[f"r{r}_dinds"])
    assert np.max(np.abs(dw - g[f"r{r}_dwgts"])) <= 1e-6        # weights carry the libm-vs-libdevice ulp
    assert scaled_err(o, g[f"r{r}_douts"]) <= TOL
    assert scaled_err(oracle.devox_grad(g[f"r{r}_dgy"], g[f"r{r}_dinds"], g[f"r{r}_dwgts"], r, True), g[f"r{r}_dgx"]) <= TOL


@pytest.mark.parametrize("r", [4, 8, 16])
def test_cube(oracle, golden_dir, r):
    g = load_golden(golden_dir, "cube.npz")
    out, ind, cnt = oracle.avg_voxelize(g[f"r{r}_feat"], g[f"r{r}_vox"], r)
    assert np.array_equal(ind, g[f"r{r}_ind"]) and np.array_equal(cnt, g[f"r{r}_cnt"])
    assert scaled_err(out, g[f"r{r}_out"]) <= TOL
    assert np.array_equal(np.rint(g[f"r{r}_norm_coords"]).astype(np.int32), g[f"r{r}_vox"])   # round-half-even
    assert rel_err(oracle.avg_voxelize_grad(g[f"r{r}_gy"], ind, cnt), g[f"r{r}_gx"]) <= TOL
    o, di, dw = oracle.trilinear_devoxelize(g[f"r{r}_norm_coords"], g[f"r{r}_grid"], r)
    assert np.array_equal(di, g[f"r{r}_dinds"]) and np.array_equal(dw, g[f"r{r}_dwgts"])
    assert np.array_equal(o, g[f"r{r}_douts"])
    assert scaled_err(oracle.devox_grad(g[f"r{r}_dgy"], di, dw, r, False), g[f"r{r}_dgx"]) <= TOL


def test_ball_query_oracle_properties(oracle):
    """ball query restatement: first-U-in-index-order, self and near-duplicates excluded, first hit fills the row."""
    g = np.random.default_rng(1)
    pts = (g.standard_normal((2, 3, 300)) * 0.3).astype(np.float32)
    idx = oracle.ball_query(pts, pts, 0.3, 16)
    x = pts.astype(np.float64)
    d2 = ((x[:, :, :, None] - x[:, :, None, :]) ** 2).sum(1)
    for b in range(2):
        for j in range(0, 300, 37):
            nb = np.nonzero((d2[b, j] < np.float32(0.3) ** 2 * (1 - 1e-6)) & (d2[b, j] > 1.01e-5))[0]
            row = idx[b, j]
            assert j not in row or len(nb) == 0
            k = min(len(nb), 16)
            assert np.array_equal(row[:k], nb[:k])
            if 0 < k < 16:
                assert (row[k:] == nb[0]).all()
            if k == 0:
                assert (row == 0).all()
    f = g.standard_normal((2, 4, 300)).astype(np.float32)
    grp = oracle.grouping(f, idx)
    assert np.array_equal(grp[1, 2], f[1, 2][idx[1]])


def test_ball_query_golden(oracle, golden_dir):
    from _util import load_golden
    g = load_golden(golden_dir, "ball_query.npz")
    for name in "abc":
        idx = oracle.ball_query(g["centers"], g["points"], float(g[name + "_radius"]), int(g[name + "_u"]))
        assert np.array_equal(idx, g[name + "_idx"])
        assert np.array_equal(oracle.grouping(g[name + "_feat"], idx), g[name + "_grp"])
        gx = oracle.grouping_grad(g[name + "_gy"], idx, g["points"].shape[2])
        assert np.abs(gx - g[name + "_gx"]).max() <= 1e-5 * np.abs(g[name + "_gx"]).max()
