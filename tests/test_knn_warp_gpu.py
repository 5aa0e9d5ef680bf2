"""The warp-per-query k-NN (csrc/knn_warp.cu: what torch.ops.ri.knn / knn_one and the engine run for up to 1024 references)
against the one-thread-per-query kernel (csrc/knn.cu: references visited in index order, the reference kernel's own insertion
rule) and the C oracle: every index and every distance bit, including the cases where a selection by threshold + sort could
differ from a running insertion — exact ties, duplicated points, lattices, fewer references than k, non-finite coordinates,
distances beyond the 10000 cut-off.  test_parity_gpu.py pins both forms to the reference kernel (golden vectors and live)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def both(q, r, k):
    d0, i0 = torch.ops.ri.knn_brute(q, r, k)
    d1, i1 = torch.ops.ri.knn_one(q, r, k)
    torch.cuda.synchronize()
    return d0, i0, d1, i1


def assert_same(q, r, k):
    d0, i0, d1, i1 = both(q, r, k)
    assert torch.equal(i0, i1), "indices differ in %d places" % int((i0 != i1).sum())
    assert torch.equal(d0.view(torch.int32), d1.view(torch.int32)), "distance bits differ"


@pytest.mark.parametrize("B,n,m,k", [(32, 1024, 1024, 20), (3, 1000, 777, 16), (2, 33, 2048, 5), (4, 512, 512, 32),
                                     (2, 64, 64, 1), (3, 2048, 1500, 20), (2, 1, 1, 3), (2, 1025, 1025, 8), (5, 100, 31, 20)])
def test_warp_equals_brute_on_surfaces(B, n, m, k):
    from ri_b200 import synth
    a = T(synth.make_clouds(B, n, seed=n + m)[:, :3])
    b = T(synth.make_clouds(B, m, seed=n * 7 + m)[:, :3])
    assert_same(a, b, k)
    assert_same(a, a, k)                     # self query: one sorted set
    assert_same(b, a, k)


@pytest.mark.parametrize("B,n,k", [(4, 1024, 20), (2, 700, 32)])
def test_warp_equals_brute_gaussian_and_clusters(B, n, k):
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    a = torch.randn((B, 3, n), device="cuda", generator=g)
    assert_same(a, a, k)
    c = a.clone()
    c[:, :, : n // 2] = c[:, :, : n // 2] * 1e-3 + 5.0        # a tight far-away cluster: very uneven cells
    assert_same(c, c, k)
    assert_same(a, c, k)


@pytest.mark.parametrize("k", [8, 20])
def test_warp_ties_lattice_and_duplicates(k, oracle):
    # integer lattice, shuffled: every query has many equidistant neighbours, the lower ORIGINAL index has to win
    rng = np.random.default_rng(5)
    g = np.stack(np.meshgrid(np.arange(10), np.arange(10), np.arange(10), indexing="ij"), 0).reshape(3, -1).astype(np.float32)
    clouds = np.stack([g[:, rng.permutation(1000)] for _ in range(3)])
    x = T(clouds)
    assert_same(x, x, k)
    d, i = torch.ops.ri.knn_one(x, x, k)
    od, oi = oracle.knn_one(clouds, clouds, k)
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.cpu().numpy(), od)
    # duplicated points (each point four times) and an all-equal cloud
    dup = np.repeat(rng.standard_normal((2, 3, 256)).astype(np.float32), 4, axis=2)
    dup = np.stack([c[:, rng.permutation(1024)] for c in dup])
    assert_same(T(dup), T(dup), k)
    same = T(np.full((2, 3, 300), 0.25, np.float32))
    assert_same(same, same, k)
    d, i = torch.ops.ri.knn_one(same, same, k)
    assert torch.equal(i[0, :, 7].cpu(), torch.arange(k, dtype=torch.int32))          # lowest indices first


def test_warp_non_finite_and_far_points():
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    a = torch.randn((3, 3, 900), device="cuda", generator=g)
    a[0, 0, 5] = float("nan"); a[0, 2, 77] = float("inf"); a[1, 1, 100] = -float("inf")
    a[2, :, 200:260] *= 300.0                                  # beyond the 10000 cut-off: never a neighbour
    b = torch.randn((3, 3, 1100), device="cuda", generator=g)
    b[1, :, 3] = float("nan"); b[2, 0, 9] = float("inf")
    assert_same(a, a, 20)
    assert_same(a, b, 20)
    assert_same(b, a, 16)
    far = torch.randn((2, 3, 64), device="cuda", generator=g) + 500.0     # every distance > 10000: all slots stay (10000, 0)
    near = torch.randn((2, 3, 64), device="cuda", generator=g)
    d0, i0, d1, i1 = both(near, far, 4)
    assert torch.equal(i0, i1) and torch.equal(d0, d1) and float(d1.min()) == 10000.0 and int(i1.max()) == 0


def test_bilateral_op_uses_both_sorted_sets():
    from ri_b200 import synth
    a = T(synth.make_clouds(4, 1024, seed=1)[:, :3]); b = T(synth.make_clouds(4, 900, seed=2)[:, :3])
    d1, d2, i1, i2 = torch.ops.ri.knn(a, b, 20)
    e1, j1 = torch.ops.ri.knn_brute(a, b, 20); e2, j2 = torch.ops.ri.knn_brute(b, a, 20)
    assert torch.equal(i1, j1) and torch.equal(i2, j2) and torch.equal(d1, e1) and torch.equal(d2, e2)
    d1, d2, i1, i2 = torch.ops.ri.knn(a, a, 20)
    e1, j1 = torch.ops.ri.knn_brute(a, a, 20)
    assert torch.equal(i1, j1) and torch.equal(i2, j1) and torch.equal(d1, e1) and torch.equal(d2, e1)


def test_warp_is_deterministic_across_runs():
    from ri_b200 import synth
    a = T(synth.make_clouds(32, 1024, seed=11)[:, :3])
    d0, i0 = torch.ops.ri.knn_one(a, a, 20)
    for _ in range(5):
        d, i = torch.ops.ri.knn_one(a, a, 20)
        assert torch.equal(i, i0) and torch.equal(d, d0)


@pytest.mark.parametrize("n,k", [(1024, 20), (640, 32), (1000, 8)])
def test_warp_clustered_index_classes_take_the_streaming_path(n, k):
    """References whose index classes modulo 32 are tight spatial clusters: one lane of the warp-per-query kernel then owns
    all the near candidates, its threshold (k-th smallest per-lane minimum) admits hundreds of survivors, and the kernel has to
    stream them instead of compacting them into its key buffer.  Same bits as the thread-per-query kernel."""
    rng = np.random.default_rng(n + k)
    centres = rng.standard_normal((32, 3)).astype(np.float32) * 2.0
    x = np.empty((3, 3, n), np.float32)
    for b in range(3):
        cls = np.arange(n) % 32
        x[b] = (centres[cls] + rng.standard_normal((n, 3)).astype(np.float32) * (1e-3 if b < 2 else 0.0)).T   # cloud 2: exact duplicates
    assert_same(T(x), T(x), k)
    q = T(rng.standard_normal((3, 3, 333)).astype(np.float32))
    assert_same(q, T(x), k)


def test_warp_runs_cross_cloud_boundaries():
    # many small clouds: every warp's run of consecutive (cloud, query) pairs spans several clouds, odd sizes break the pairs
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    for B, n, m in ((257, 7, 50), (64, 33, 33), (100, 1, 40), (31, 95, 1000)):
        a = torch.randn((B, 3, n), device="cuda", generator=g)
        b = torch.randn((B, 3, m), device="cuda", generator=g)
        assert_same(a, b, 5)
