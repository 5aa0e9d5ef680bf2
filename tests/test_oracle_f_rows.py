"""CPU checks of the oracle restatements behind rows f1 (local PPF) and f3 (pose, metrics): the numpy code against the reference's
own torch operations run on the CPU (local PPF, pvcnn_classify.py:252-270) and against closed-form cases (Kabsch, RE/TE/RMSE)."""
import numpy as np
import torch


def test_local_ppf_oracle_equals_reference_torch_ops_on_cpu(oracle):
    rng = np.random.default_rng(3)
    B, N, U = 2, 300, 16
    xyz = rng.uniform(-1, 1, (B, 3, N)).astype(np.float32)
    nrm = rng.standard_normal((B, 3, N)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    idx = oracle.ball_query(xyz, xyz, 0.4, U)                                   # [B,M,U]
    got = oracle.local_ppf(xyz, nrm, idx)
    # the reference's ops: BallQuery.forward (ball_query.py:16-35) then pvcnn_classify.py:258-269, torch on the CPU
    coords, normals = torch.from_numpy(xyz), torch.from_numpy(nrm)
    gi = torch.from_numpy(idx.astype(np.int64)).reshape(B, 1, N * U).expand(-1, 3, -1)
    nb_c = torch.gather(coords, 2, gi).reshape(B, 3, N, U) - coords.unsqueeze(-1)        # grouping - centres
    nb_n = torch.gather(normals, 2, gi).reshape(B, 3, N, U)
    g = torch.cat([nb_c, nb_n], 1).permute(0, 1, 3, 2)                                    # [b, 6, u, m]
    neighbor_coords, neighbor_normals = g[:, :3], g[:, 3:]
    ck = coords.unsqueeze(2).expand(-1, -1, U, -1); nk = normals.unsqueeze(2).expand(-1, -1, U, -1)
    d = ck - neighbor_coords
    d_norm = torch.norm(d, dim=1, p=2, keepdim=True)
    d_unit = d / d_norm
    want = torch.cat((torch.acos(neighbor_normals.mul(d_unit).sum(dim=1, keepdim=True).clamp(-1, 1)),
                      torch.acos(nk.mul(d_unit).sum(dim=1, keepdim=True).clamp(-1, 1)),
                      torch.acos(neighbor_normals.mul(nk).sum(dim=1, keepdim=True).clamp(-1, 1)), d_norm), dim=1).numpy()
    assert got.shape == want.shape == (B, 4, U, N)
    assert np.abs(got[:, 3] - want[:, 3]).max() <= 1e-6
    edge = np.abs(np.cos(want[:, :3])) > 1 - 1e-4
    assert np.abs(got[:, :3] - want[:, :3])[~edge].max() <= 2e-6


def test_kabsch_and_metrics_closed_form(oracle):
    rng = np.random.default_rng(1)
    a = rng.uniform(-1, 1, (200, 3))
    ang = 0.7
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]])
    t = np.array([0.3, -0.2, 0.5])
    T = oracle.kabsch(a, a @ R.T + t)
    assert np.allclose(T[:3, :3], R, atol=1e-12) and np.allclose(T[:3, 3], t, atol=1e-12)
    gt = np.eye(4); gt[:3, :3] = R; gt[:3, 3] = t
    est = np.eye(4); est[:3, 3] = t + np.array([0.0, 0.0, 0.1])                 # identity rotation, 0.1 off in z
    m = oracle.registration_metrics(gt[None], est[None], a[None])
    assert abs(m[0, 0] - np.degrees(ang)) < 1e-9 and abs(m[0, 1] - 0.1) < 1e-12
    want_rmse = np.mean(np.linalg.norm((a @ np.eye(3) + est[:3, 3]) - (a @ R.T + t), axis=1))
    assert abs(m[0, 2] - want_rmse) < 1e-12
    assert oracle.count_inliers(gt, a, a @ R.T + t, 1e-9) == 200
