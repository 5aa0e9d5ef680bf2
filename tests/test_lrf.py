"""Row f4: the 'change_coords' rotation-invariant preprocessing (pvcnn_classify.py:153-184).
Golden vectors: the reference's torch operations run line for line on the CPU (oracle/make_golden_lrf.py) — a transcription,
the class itself cannot be instantiated without open3d / the CUDA backend.  Bars: the chosen frame agrees (same base points:
bases within 1e-5), new coordinates within 1e-5 of the cloud scale (fp32); and the property the op exists for: the output is
invariant to a rotation of the input cloud."""
import numpy as np
import pytest

from _util import load_golden, scaled_err, TOL


@pytest.mark.parametrize("name", ["surface", "parallel"])
def test_oracle_matches_torch_transcription(oracle, golden_dir, name):
    g = load_golden(golden_dir, "lrf.npz")
    new, ok = oracle.change_coords(g[name + "_coords"], g[name + "_mean"])
    assert ok.all()
    assert scaled_err(new, g[name + "_new"]) <= TOL


def test_oracle_flags_degenerate_clouds(oracle):
    x = np.zeros((3, 3, 64), np.float32)
    x[1, 0] = np.linspace(-1, 1, 64)                       # collinear: every direction is parallel to base_x
    x[2] = np.random.default_rng(0).standard_normal((3, 64))
    _, ok = oracle.change_coords(x)
    assert ok.tolist() == [0, 0, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["surface", "parallel"])
def test_cuda_matches_golden_and_oracle(oracle, golden_dir, name):
    import torch
    import ri_b200
    g = load_golden(golden_dir, "lrf.npz")
    x = torch.from_numpy(g[name + "_coords"]).cuda()
    new, bases = ri_b200.functional.change_coords(x, return_bases=True)
    assert np.abs(bases.cpu().numpy() - g[name + "_bases"]).max() <= 1e-5
    assert scaled_err(new.cpu().numpy(), g[name + "_new"]) <= TOL
    onew, ok = oracle.change_coords(g[name + "_coords"], x.mean(dim=2).cpu().numpy())
    assert ok.all() and scaled_err(new.cpu().numpy(), onew) <= TOL


@pytest.mark.gpu
def test_cuda_rotation_invariance_and_interleaved_input():
    """Full-size property: rotating a cloud does not change its coordinates in its own frame; [B,6,N] input accepted."""
    import torch
    import ri_b200
    from ri_b200 import synth
    pts = synth.make_clouds(32, 1024, seed=5)                              # [B,6,N]
    rng = np.random.default_rng(1)
    Q, _ = np.linalg.qr(rng.standard_normal((32, 3, 3)))
    Q *= np.sign(np.linalg.det(Q))[:, None, None]
    rot = np.einsum("bij,bjn->bin", Q, pts[:, :3]).astype(np.float32)
    a = ri_b200.functional.change_coords(torch.from_numpy(pts).cuda())
    b = ri_b200.functional.change_coords(torch.from_numpy(rot).cuda())
    assert a.shape == (32, 3, 1024)
    assert float((a - b).abs().max()) < 2e-5
    # the frame is orthonormal: distances to the centroid are preserved
    c = torch.from_numpy(pts[:, :3]).cuda()
    c = c - c.mean(2, keepdim=True)
    assert float((a.norm(dim=1) - c.norm(dim=1)).abs().max()) < 1e-5


@pytest.mark.gpu
def test_cuda_degenerate_clouds_assert_like_the_reference():
    import torch
    import ri_b200
    x = torch.zeros(2, 3, 64, device="cuda")
    x[1, 0] = torch.linspace(-1, 1, 64)
    with pytest.raises(AssertionError):
        ri_b200.functional.change_coords(x)
    out = ri_b200.functional.change_coords(x, check=False)
    assert float(out.abs().max()) == 0.0
