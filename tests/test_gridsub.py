"""Barycentre grid subsampling: C oracle vs the golden vectors made by the reference's own CPU code
(oracle/make_golden_gridsub.py), vs the reference library live when it is present, and the CUDA kernels vs both.
Bars: the set of cells, every barycentre and every mean feature bit-exact; labels: a most-frequent label of the cell (the
reference's choice between equally frequent labels is an unordered_map's iteration order)."""
import numpy as np
import pytest

from _util import load_golden

CASES = ["room_6k_dl10", "room_30k_dl04", "cloud_1k_dl02", "dup_dl05"]


def point_keys(pts, dl, sub_keys):
    """Cell index of every input point, fp32 arithmetic of grid_subsampling.cpp:25-56."""
    pts = pts.astype(np.float32); dl = np.float32(dl)
    inv = np.float32(1) / dl
    org = np.floor(pts.min(0) * inv) * dl
    n = (np.floor((pts.max(0) - org) / dl).astype(np.uint64) + np.uint64(1))
    i = np.floor((pts - org) / dl).astype(np.uint64)
    return i[:, 0] + n[0] * i[:, 1] + n[0] * n[1] * i[:, 2]


def check_labels(pts, labels, dl, keys, got, want):
    """`got` and `want` both name a most-frequent label of every cell."""
    pk = point_keys(pts, dl, keys)
    order = np.argsort(pk, kind="stable")
    starts = np.searchsorted(pk[order], keys)
    ends = np.append(starts[1:], len(pk))
    for c in range(len(keys)):
        cell = labels[order[starts[c]:ends[c]]]
        for col in range(labels.shape[1]):
            vals, counts = np.unique(cell[:, col], return_counts=True)
            best = counts.max()
            assert counts[vals == got[c, col]][0] == best
            assert counts[vals == want[c, col]][0] == best
            assert got[c, col] == vals[counts == best].min()          # our documented tie rule


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(oracle, golden_dir, name):
    g = load_golden(golden_dir, "gridsub.npz")
    f = g[name + "_features"] if name + "_features" in g else None
    l = g[name + "_labels"] if name + "_labels" in g else None
    p, of, ol, keys = oracle.grid_subsample(g[name + "_points"], f, l, float(g[name + "_dl"]))
    assert np.array_equal(keys, g[name + "_keys"])
    assert np.array_equal(p, g[name + "_sub_points"])
    if f is not None:
        assert np.array_equal(of, g[name + "_sub_features"])
    if l is not None:
        check_labels(g[name + "_points"], l, float(g[name + "_dl"]), keys, ol, g[name + "_sub_labels"])


def test_oracle_matches_reference_live(oracle):
    """Against oracle/_ref/libgridsub_ref.so (the unmodified reference sources) on fresh inputs, incl. negative coordinates."""
    from oracle import build_ref
    rng = np.random.default_rng(7)
    pts = (rng.standard_normal((5000, 3)) * np.array([2.0, 1.0, 0.5])).astype(np.float32)
    feats = rng.standard_normal((5000, 5)).astype(np.float32)
    ref = build_ref.ref_grid_subsampling(pts, feats, None, 0.07)
    if ref is None:
        pytest.skip("reference library not built (no /root/reference and no prebuilt oracle/_ref/libgridsub_ref.so)")
    p, of, _, _ = oracle.grid_subsample(pts, feats, None, 0.07)
    rows = lambda a, b: set(map(bytes, np.concatenate([a, b], 1)))
    assert len(p) == len(ref[0])
    assert rows(p, of) == rows(ref[0], ref[1])


def test_empty_and_single(oracle):
    p, _, _, k = oracle.grid_subsample(np.zeros((1, 3), np.float32), None, None, 0.1)
    assert p.shape == (1, 3) and np.array_equal(p, np.zeros((1, 3), np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_golden_and_oracle(oracle, golden_dir, name):
    import torch
    import ri_b200
    g = load_golden(golden_dir, "gridsub.npz")
    f = g[name + "_features"] if name + "_features" in g else None
    l = g[name + "_labels"] if name + "_labels" in g else None
    dl = float(g[name + "_dl"])
    out = ri_b200.grid_sub_sampling(g[name + "_points"], f, l, grid_size=dl)
    out = list(out) if isinstance(out, tuple) else [out]
    assert isinstance(out[0], np.ndarray)
    assert np.array_equal(out[0], g[name + "_sub_points"])
    op, of, ol, keys = oracle.grid_subsample(g[name + "_points"], f, l, dl)
    assert np.array_equal(out[0], op)
    if f is not None:
        assert np.array_equal(out[1], g[name + "_sub_features"]) and np.array_equal(out[1], of)
    if l is not None:
        assert np.array_equal(out[-1], ol)                         # same tie rule as the oracle
        check_labels(g[name + "_points"], l, dl, keys, out[-1], g[name + "_sub_labels"])


@pytest.mark.gpu
def test_cuda_scan_sized_and_tensor_api(oracle):
    """A 400k-point scan (BASELINE configs[3] pre-step: subsample to ~50k), CUDA tensors in and out."""
    import torch
    import ri_b200
    from ri_b200 import synth
    pts = np.ascontiguousarray(synth.make_scan(400000, seed=5)[:3].T)          # [6,N] (xyz | normal) -> (N,3)
    t = torch.from_numpy(pts).cuda()
    sub = ri_b200.grid_sub_sampling(t, grid_size=0.05)
    assert sub.is_cuda
    op, _, _, _ = oracle.grid_subsample(pts, None, None, 0.05)
    assert np.array_equal(sub.cpu().numpy(), op)
    # idempotence-like property at full size: every barycentre lies in its own cell, so subsampling the result at the
    # same cell size cannot produce more cells
    again = ri_b200.grid_sub_sampling(sub, grid_size=0.05)
    assert again.shape[0] <= sub.shape[0]
    # empty cloud
    e = ri_b200.grid_sub_sampling(torch.empty((0, 3), device="cuda"), grid_size=0.1)
    assert e.shape == (0, 3)


@pytest.mark.gpu
def test_cuda_labels_in_large_cells(oracle):
    """A coarse grid on a scan: hundreds to thousands of points per cell — the label vote must stay linear in the cell size
    (it used to recount the whole cell for every point) and still name the most frequent label, smallest on ties.  One label
    column has more distinct values per cell than the kernel's counting table holds: that column takes the quadratic path."""
    import time
    import torch
    import ri_b200
    from ri_b200 import synth
    rng = np.random.default_rng(12)
    N = 60000
    pts = np.ascontiguousarray(synth.make_scan(N, seed=9)[:3].T)
    labels = np.stack([rng.integers(0, 13, N), rng.integers(0, 3, N), rng.integers(0, 500, N)], 1).astype(np.int32)
    torch.cuda.synchronize(); t0 = time.time()
    sub_p, sub_l = ri_b200.grid_sub_sampling(torch.from_numpy(pts).cuda(), labels=torch.from_numpy(labels).cuda(), grid_size=0.6)
    torch.cuda.synchronize(); took = time.time() - t0
    op, _, ol, _ = oracle.grid_subsample(pts, None, labels, 0.6)
    assert np.array_equal(sub_p.cpu().numpy(), op) and np.array_equal(sub_l.cpu().numpy(), ol)
    assert sub_p.shape[0] < 400 and took < 2.0


@pytest.mark.gpu
@pytest.mark.parametrize("extent,dl", [(0.1, 0.05), (0.5, 0.05), (3.0, 0.05), (40.0, 0.05), (100.0, 0.05), (300.0, 0.01),
                                       (1000.0, 0.01), (3000.0, 0.01)])
def test_cuda_sort_pass_counts(oracle, extent, dl):
    """The radix sort of the cell keys runs ceil(bits / 8) passes, decided on the device from the largest key: from one pass
    (64 cells) to eight (3 km at 1 cm: ~2^58 cells), odd and even counts (the result lands in either buffer), with
    duplicate cells (stability: points of a cell are summed in input order) — all bit-equal to the C oracle."""
    import torch
    import ri_b200
    rng = np.random.default_rng(int(extent * 7) % 1000)
    N = 30011                                                     # not a multiple of the sort's 4096-pair tile
    base = rng.uniform(-extent, extent, (N // 3, 3))
    pts = np.concatenate([base, base + rng.uniform(0, dl * 0.3, base.shape), rng.uniform(-extent, extent, (N - 2 * (N // 3), 3))])
    pts = pts[rng.permutation(N)].astype(np.float32)
    feats = rng.standard_normal((N, 2)).astype(np.float32)
    sub_p, sub_f = ri_b200.grid_sub_sampling(torch.from_numpy(pts).cuda(), features=torch.from_numpy(feats).cuda(), grid_size=dl)
    op, of, _, keys = oracle.grid_subsample(pts, feats, None, dl)
    assert np.array_equal(sub_p.cpu().numpy(), op) and np.array_equal(sub_f.cpu().numpy(), of)
    assert np.all(keys[1:] > keys[:-1])                          # ascending cell order


@pytest.mark.gpu
def test_cuda_many_scans_one_sync(oracle):
    """grid_sub_sampling_many: several scans enqueued back to back, lengths read once — same results as one call per scan."""
    import torch
    import ri_b200
    from ri_b200 import synth
    rng = np.random.default_rng(3)
    clouds = []
    for q, n in enumerate((5000, 12345, 1, 40000)):
        pts = np.ascontiguousarray(synth.make_scan(max(n, 16), seed=20 + q)[:3].T)[:n]
        f = rng.standard_normal((n, 3)).astype(np.float32)
        clouds.append((torch.from_numpy(pts).cuda(), torch.from_numpy(f).cuda() if q != 1 else None))
    many = ri_b200.grid_sub_sampling_many(clouds, grid_size=0.07)
    assert len(many) == len(clouds)
    for (p, f), (mp, mf, ml) in zip(clouds, many):
        one = ri_b200.grid_sub_sampling(p, features=f, grid_size=0.07)
        op = one[0] if isinstance(one, tuple) else one
        assert torch.equal(mp, op) and ml is None
        if f is None:
            assert mf is None
        else:
            assert torch.equal(mf, one[1])
        ref_p, _, _, _ = oracle.grid_subsample(p.cpu().numpy(), None, None, 0.07)
        assert np.array_equal(mp.cpu().numpy(), ref_p)
    assert ri_b200.grid_sub_sampling_many([]) == []
