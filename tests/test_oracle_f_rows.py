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
    # the reference's torch calls in their order (BallQuery.forward, ball_query.py:16-35, then pvcnn_classify.py:258-269), CPU
    xyz_t, nrm_t = torch.from_numpy(xyz), torch.from_numpy(nrm)
    gi = torch.from_numpy(idx.astype(np.int64)).reshape(B, 1, N * U).expand(-1, 3, -1)
    g_xyz = (torch.gather(xyz_t, 2, gi).reshape(B, 3, N, U) - xyz_t.unsqueeze(-1)).permute(0, 1, 3, 2)     # relative, [b,3,u,m]
    g_nrm = torch.gather(nrm_t, 2, gi).reshape(B, 3, N, U).permute(0, 1, 3, 2)
    c_xyz = xyz_t.unsqueeze(2).expand(-1, -1, U, -1); c_nrm = nrm_t.unsqueeze(2).expand(-1, -1, U, -1)
    d = c_xyz - g_xyz
    length = torch.norm(d, dim=1, p=2, keepdim=True)
    unit = d / length
    ang = lambda p, q: torch.acos(p.mul(q).sum(dim=1, keepdim=True).clamp(-1, 1))
    want = torch.cat((ang(g_nrm, unit), ang(c_nrm, unit), ang(g_nrm, c_nrm), length), dim=1).numpy()
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
