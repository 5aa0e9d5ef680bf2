"""GPU parity of the tcgen05 mutual-NN matcher (csrc/matcher.cu, torch.ops.ri.mutual_nn) against the reference's numpy
code (datasets/deepgmr_mn40.py:232-244, restated verbatim in oracle/cpu_oracle.py) and an fp64 evaluation.

The reference arithmetic is an fp32 sgemm whose summation order is unspecified (numpy/OpenBLAS, unpinned), so parity is
stated the way north_star does: match DISTANCES within 1e-5 relative (relative to the magnitude of the terms that
cancel, |f1|^2 + |f2|^2), argmin INDICES equal wherever the minimum is separated from the runner-up by more than that
tolerance, and lowest-index tie-breaking on exact ties."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ri():
    import ri_b200
    return ri_b200


def _fp64(f1, f2):
    x, y = f1.astype(np.float64), f2.astype(np.float64)
    d = (x * x).sum(1)[:, None] + (y * y).sum(1)[None, :] - 2 * x @ y.T
    return d, (x * x).sum(1).max() + (y * y).sum(1).max()


def _run(ri, f1, f2, point_major):
    a = torch.from_numpy(f1 if point_major else np.ascontiguousarray(f1.transpose(0, 2, 1))).cuda()
    b = torch.from_numpy(f2 if point_major else np.ascontiguousarray(f2.transpose(0, 2, 1))).cuda()
    r = ri.matcher.mutual_nn(a, b, point_major=point_major)
    return {k: v.cpu().numpy() for k, v in r.items()}


def _check_pair(r, p, f1, f2, oracle):
    n1, n2 = f1.shape[0], f2.shape[0]
    d, scale = _fp64(f1, f2)
    c12, c21 = r['corr12'][p], r['corr21'][p]
    assert c12.min() >= 0 and c12.max() < n2 and c21.min() >= 0 and c21.max() < n1
    # every pick is a minimiser up to the tolerance
    assert (d[np.arange(n1), c12] - d.min(1)).max() <= TOL * scale
    assert (d[c21, np.arange(n2)] - d.min(0)).max() <= TOL * scale
    # reported distances: fp32 recomputation of diff[i, corr12[i]]
    assert np.abs(r['dist12'][p] - d[np.arange(n1), c12]).max() <= TOL * scale
    # mutual mask / compaction is exact integer logic on our own argmins
    mask = c21[c12] == np.arange(n1)
    cnt = int(r['count'][p])
    assert cnt == int(mask.sum())
    assert np.array_equal(r['idx1'][p][:cnt], np.arange(n1)[mask])
    assert np.array_equal(r['idx2'][p][:cnt], c12[mask])
    assert (r['idx1'][p][cnt:] == -1).all() and (r['idx2'][p][cnt:] == -1).all()
    # the reference's numpy code: identical index sets wherever its own fp32 matrix separates the minimum
    o1, o2, diff = oracle.find_correspondence_one_pair(f1, f2)
    srt = np.sort(diff, 1)
    clear_rows = (srt[:, 1] - srt[:, 0]) > 4 * TOL * scale if n2 > 1 else np.ones(n1, bool)
    assert np.array_equal(c12[clear_rows], diff.argmin(1)[clear_rows])
    srt = np.sort(diff, 0)
    clear_cols = (srt[1] - srt[0]) > 4 * TOL * scale if n1 > 1 else np.ones(n2, bool)
    assert np.array_equal(c21[clear_cols], diff.argmin(0)[clear_cols])
    if clear_rows.all() and clear_cols.all():
        assert np.array_equal(r['idx1'][p][:cnt], o1) and np.array_equal(r['idx2'][p][:cnt], o2)


@pytest.mark.parametrize("P,C,n1,n2,pm", [(1, 16, 128, 256, False), (2, 512, 1024, 1024, False), (2, 512, 1024, 1024, True),
                                          (3, 100, 300, 700, False), (2, 33, 1000, 130, True), (1, 7, 5, 3, False),
                                          (1, 512, 1, 1024, False), (2, 64, 2049, 257, False)])
def test_random_descriptors(ri, oracle, P, C, n1, n2, pm):
    g = np.random.default_rng(P * 1000 + C + n1 + n2)
    f1 = g.standard_normal((P, n1, C)).astype(np.float32)
    f2 = g.standard_normal((P, n2, C)).astype(np.float32)
    r = _run(ri, f1, f2, pm)
    for p in range(P):
        _check_pair(r, p, f1[p], f2[p], oracle)


def test_registration_shaped_descriptors_all_match(ri, oracle):
    """Target descriptors = permuted source descriptors + noise (what a good extractor yields on a DeepGMR pair):
    every point's true partner is its mutual nearest neighbour."""
    g = np.random.default_rng(7)
    P, n, C = 4, 1024, 512
    f1 = g.standard_normal((P, n, C)).astype(np.float32)
    perm = np.stack([g.permutation(n) for _ in range(P)])
    f2 = (np.stack([f1[p][perm[p]] for p in range(P)]) + 0.05 * g.standard_normal((P, n, C))).astype(np.float32)
    r = _run(ri, f1, f2, False)
    for p in range(P):
        _check_pair(r, p, f1[p], f2[p], oracle)
        assert int(r['count'][p]) == n
        inv = np.empty(n, np.int64); inv[perm[p]] = np.arange(n)
        assert np.array_equal(r['corr12'][p], inv)


def test_exact_ties_take_lowest_index(ri):
    """Duplicated descriptors give bit-identical distances: np.argmin returns the first index, so must we."""
    g = np.random.default_rng(3)
    n, C = 256, 64
    base = g.integers(-4, 5, size=(n // 2, C)).astype(np.float32)       # small integers: every product is exact
    f2 = np.concatenate([base, base], 0)[None]                          # row j and j + n/2 are identical
    f1 = base[None].copy()
    r = _run(ri, f1, f2, False)
    assert np.array_equal(r['corr12'][0], np.arange(n // 2))            # not j + n/2
    f1b = np.concatenate([base, base], 0)[None]
    r = _run(ri, f1b, base[None].copy(), False)
    assert np.array_equal(r['corr21'][0], np.arange(n // 2))            # lowest row index of the duplicate pair
    assert np.array_equal(r['dist12'][0], np.zeros(n, np.float32))


def test_reference_signature_wrapper(ri, oracle):
    g = np.random.default_rng(11)
    f1 = g.standard_normal((700, 128)).astype(np.float32); f2 = g.standard_normal((650, 128)).astype(np.float32)
    i1, i2 = ri.matcher.find_correspondence_one_pair(f1, f2)
    o1, o2, _ = oracle.find_correspondence_one_pair(f1, f2)
    assert i1.dtype == np.int64 and np.array_equal(i1, o1) and np.array_equal(i2, o2)


def test_full_size_properties(ri):
    """DeepGMR shape (BASELINE configs[2], one rank's share at 8 GPUs): 32 pairs x 1024 x 1024 x 512.  Size-independent
    properties: swapping the operands transposes the result; matches are one-to-one; a pair matched with itself is the
    identity with zero distance."""
    P, C, n = 32, 512, 1024
    a = torch.randn(P, C, n, device="cuda"); b = torch.randn(P, C, n, device="cuda")
    r = ri.matcher.mutual_nn(a, b); s = ri.matcher.mutual_nn(b, a)
    assert torch.equal(r['corr12'], s['corr21']) and torch.equal(r['corr21'], s['corr12'])
    assert torch.equal(r['count'], s['count'])
    for p in range(0, P, 7):
        k = int(r['count'][p]); i2 = r['idx2'][p, :k]
        assert i2.unique().numel() == k
    t = ri.matcher.mutual_nn(a, a)
    ar = torch.arange(n, device="cuda", dtype=torch.int32).expand(P, n)
    assert torch.equal(t['corr12'], ar) and torch.equal(t['idx2'], ar) and (t['count'] == n).all()
    assert t['dist12'].abs().max().item() <= 1e-5 * 2 * float((a * a).sum(1).max())


def test_matcher_errors(ri):
    a = torch.randn(2, 8, 16, device="cuda")
    with pytest.raises(RuntimeError):
        torch.ops.ri.mutual_nn(a.cpu(), a, False)
    with pytest.raises(RuntimeError):
        torch.ops.ri.mutual_nn(a, torch.randn(2, 9, 16, device="cuda"), False)
    with pytest.raises(RuntimeError):
        torch.ops.ri.mutual_nn(a, torch.randn(3, 8, 16, device="cuda"), False)


def test_tensor_core_truncates_tf32_operands():
    """The GEMM feeds the RAW fp32 descriptor as the 'hi' operand and only writes the 'lo' plane (a - trunc(a)): that is
    exact only if the tensor core ignores the low 13 mantissa bits of a tf32 operand (truncation), not if it rounds.  Pin it:
    every channel of f1 is alpha = 1 + 0.75 * 2^-10 (rounds UP to 1 + 2^-10, truncates to 1); f2[0] = 0, f2[1] = beta * ones
    with beta = 2 + 2^-9 chosen so that the true distances are |a|^2 and |a|^2 + 0.5, while a rounding tensor core would see
    the second one 2.0 smaller and pick it."""
    import ri_b200
    C, n1 = 512, 128
    alpha = np.float32(1 + 0.75 * 2.0 ** -10)
    beta = np.float32(2 + 2.0 ** -9)
    f1 = np.full((1, n1, C), alpha, np.float32)
    f2 = np.zeros((1, 2, C), np.float32); f2[0, 1] = beta
    x, y = f1[0].astype(np.float64), f2[0].astype(np.float64)
    d = (x * x).sum(1)[:, None] + (y * y).sum(1)[None] - 2 * x @ y.T
    assert 0.4 < d[0, 1] - d[0, 0] < 0.6
    r = ri_b200.matcher.mutual_nn(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda(), point_major=True)
    assert int(r["corr12"].abs().sum()) == 0, "the tensor core rounded a tf32 operand: set kWriteHi = true in csrc/matcher.cu"
    assert abs(float(r["dist12"][0, 0]) - d[0, 0]) <= 1e-5 * d[0, 0]


def test_persistent_gemm_stress_two_streams():
    """The persistent GEMM (stage ring across tiles, double-buffered TMEM accumulator, 22 warps in four roles) run 400 times,
    alternating between two streams so that calls overlap on the device (TMEM allocation contention included), for three
    shapes: every call must reproduce the first call's bits."""
    import ri_b200
    for (P, C, n1, n2) in ((32, 512, 1024, 1024), (5, 100, 300, 700), (64, 64, 128, 256)):
        g = torch.Generator(device="cuda"); g.manual_seed(P)
        d1 = torch.randn((P, C, n1), device="cuda", generator=g); d2 = torch.randn((P, C, n2), device="cuda", generator=g)
        mms = [ri_b200.matcher.MutualMatcher(P, C, n1, n2) for _ in range(2)]
        streams = [torch.cuda.Stream() for _ in range(2)]
        mms[0](d1, d2); torch.cuda.synchronize()
        want = (mms[0].corr12.clone(), mms[0].corr21.clone(), mms[0].dist12.clone(), mms[0].count.clone())
        for it in range(400):
            q = it & 1
            with torch.cuda.stream(streams[q]):
                mms[q](d1, d2)
        torch.cuda.synchronize()
        for mm in mms:
            assert torch.equal(mm.corr12, want[0]) and torch.equal(mm.corr21, want[1])
            assert torch.equal(mm.dist12, want[2]) and torch.equal(mm.count, want[3])


@pytest.mark.parametrize("P,C,n1,n2,pm", [(3, 512, 1024, 1024, False), (2, 100, 300, 700, False), (2, 33, 1000, 130, True), (4, 64, 128, 256, False)])
def test_pair_and_single_cta_forms_agree(P, C, n1, n2, pm):
    """The two persistent GEMM forms (CTA pair with cta_group::2 over 256 x 256 tiles; single CTA over 128 x 256 tiles) return the
    same argmins, matches and distances."""
    import os
    import ri_b200
    g = torch.Generator(device="cuda"); g.manual_seed(C + n1)
    shape1, shape2 = ((P, n1, C), (P, n2, C)) if pm else ((P, C, n1), (P, C, n2))
    d1 = torch.randn(shape1, device="cuda", generator=g); d2 = torch.randn(shape2, device="cuda", generator=g)
    res = {}
    L = ri_b200._lib.lib
    try:
        for mode in ("0", "1"):
            assert L.ri_debug_set_knob(b"RI_MATCH_PAIR", int(mode)) == 0
            r = ri_b200.matcher.mutual_nn(d1, d2, point_major=pm)
            torch.cuda.synchronize()
            res[mode] = {k: v.clone() for k, v in r.items()}
    finally:
        L.ri_debug_set_knob(b"RI_MATCH_PAIR", -1)
    for k in res["0"]:
        assert torch.equal(res["0"][k], res["1"][k]), k


@pytest.mark.parametrize("P,C,n1,n2,pm", [(3, 512, 1024, 1024, False), (2, 100, 300, 700, False), (2, 33, 1000, 130, True), (1, 16, 37, 53, False),
                                          (2, 33, 1000, 132, False), (3, 70, 301, 515, False), (5, 128, 1024, 512, False), (2, 256, 260, 1028, False)])
def test_indices_only_mode_equals_full_mode(P, C, n1, n2, pm):
    """dist12 = NULL (the reference method returns (idx1, idx2) only): the distance re-evaluation is skipped, every other output is
    the same as with it.  For channel-major descriptors with n1 > 128 and n % 4 == 0 the indices-only call takes the no-image path
    (operands through tensor maps as MN-major UMMA operands, norms by their own kernel): it must return the same bits as the
    pre-pass path, also where columns / channels run past the tensor (C not a multiple of 16, n not a multiple of 32 or 256)."""
    import ri_b200
    g = torch.Generator(device="cuda"); g.manual_seed(100 + C)
    shp1, shp2 = ((P, n1, C), (P, n2, C)) if pm else ((P, C, n1), (P, C, n2))
    d1 = torch.randn(shp1, device="cuda", generator=g); d2 = torch.randn(shp2, device="cuda", generator=g)
    full = ri_b200.matcher.MutualMatcher(P, C, n1, n2, point_major=pm)
    idx = ri_b200.matcher.MutualMatcher(P, C, n1, n2, point_major=pm, want_dist=False)
    full(d1, d2); idx(d1, d2); torch.cuda.synchronize()
    assert idx.dist12 is None
    for name in ("corr12", "corr21", "idx1", "idx2", "count"):
        assert torch.equal(getattr(full, name), getattr(idx, name)), name


def test_indices_only_stress_two_streams():
    """The tensor-map form of the persistent CTA-pair GEMM under the same stress as the pre-pass form: 300 calls alternating
    between two streams, every call reproduces the first call's bits (and none hangs)."""
    import ri_b200
    for (P, C, n1, n2) in ((32, 512, 1024, 1024), (5, 100, 300, 700)):
        g = torch.Generator(device="cuda"); g.manual_seed(P)
        d1 = torch.randn((P, C, n1), device="cuda", generator=g); d2 = torch.randn((P, C, n2), device="cuda", generator=g)
        mms = [ri_b200.matcher.MutualMatcher(P, C, n1, n2, want_dist=False) for _ in range(2)]
        streams = [torch.cuda.Stream() for _ in range(2)]
        mms[0](d1, d2); torch.cuda.synchronize()
        want = (mms[0].corr12.clone(), mms[0].corr21.clone(), mms[0].idx1.clone(), mms[0].count.clone())
        for it in range(300):
            q = it & 1
            with torch.cuda.stream(streams[q]):
                mms[q](d1, d2)
        torch.cuda.synchronize()
        for mm in mms:
            assert torch.equal(mm.corr12, want[0]) and torch.equal(mm.corr21, want[1])
            assert torch.equal(mm.idx1, want[2]) and torch.equal(mm.count, want[3])
