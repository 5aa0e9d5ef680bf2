"""Synthetic, seeded, ModelNet40-shaped inputs (there are no datasets in this environment).

Shapes follow what the reference's loaders hand to the model: xyz centred, unit-ish scale, unit normals, fp32,
layout [B, 6, N] = (xyz | normal) (/root/reference/datasets/modelnet40.py:43-63), random SO(3) applied
(configs/.../SO3_SO3/__init__.py:18), N(0, 0.01^2) jitter clipped to +-0.05 (datasets/deepgmr_partial.py:100-106).
Surfaces: unit sphere, axis-aligned box, two-plane "L", anisotropic ellipsoid — analytic normals.
Everything is numpy on the host; callers move it to the GPU (or pinned memory) themselves.
"""
import numpy as np

__all__ = ['make_clouds', 'make_features', 'make_pairs', 'random_rotations', 'make_scan']


def random_rotations(rng, n):
    """n uniformly random rotation matrices [n,3,3] (QR of a Gaussian, sign-fixed)."""
    a = rng.standard_normal((n, 3, 3))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diagonal(r, axis1=1, axis2=2))[:, None, :]
    det = np.linalg.det(q)
    q[:, :, 0] *= det[:, None]
    return q


def _sphere(rng, n):
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v, v.copy()


def _ellipsoid(rng, n):
    s = rng.uniform(0.3, 1.0, 3)
    v, _ = _sphere(rng, n)
    p = v * s
    nrm = v / s                                  # gradient of the implicit surface
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return p, nrm


def _box(rng, n):
    h = rng.uniform(0.3, 1.0, 3)                 # half extents
    area = np.array([h[1] * h[2], h[0] * h[2], h[0] * h[1]])
    axis = rng.choice(3, size=n, p=area / area.sum())
    sign = rng.choice([-1.0, 1.0], size=n)
    p = rng.uniform(-1.0, 1.0, (n, 3)) * h
    nrm = np.zeros((n, 3))
    p[np.arange(n), axis] = sign * h[axis]
    nrm[np.arange(n), axis] = sign
    return p, nrm


def _lshape(rng, n):
    half = n // 2
    p = np.zeros((n, 3)); nrm = np.zeros((n, 3))
    p[:half, :2] = rng.uniform(-1.0, 1.0, (half, 2)); nrm[:half, 2] = 1.0           # floor z = 0
    p[half:, 0] = rng.uniform(-1.0, 1.0, n - half)
    p[half:, 2] = rng.uniform(0.0, 1.5, n - half)
    p[half:, 1] = -1.0; nrm[half:, 1] = 1.0                                         # wall y = -1
    perm = rng.permutation(n)
    return p[perm], nrm[perm]


_SURFACES = (_sphere, _box, _lshape, _ellipsoid)


def make_clouds(B, N, seed=0, rotate=True, jitter=0.01):
    """-> float32 [B, 6, N]: centred xyz (largest radius scaled to <= 1) and unit normals."""
    rng = np.random.default_rng(seed)
    out = np.empty((B, 6, N), np.float32)
    rots = random_rotations(rng, B) if rotate else np.broadcast_to(np.eye(3), (B, 3, 3))
    for b in range(B):
        p, nrm = _SURFACES[int(rng.integers(len(_SURFACES)))](rng, N)
        p = p @ rots[b].T
        nrm = nrm @ rots[b].T
        if jitter > 0:
            p = p + np.clip(rng.standard_normal((N, 3)) * jitter, -0.05, 0.05)
        p = p - p.mean(0, keepdims=True)
        p = p / max(np.linalg.norm(p, axis=1).max(), 1e-12)
        out[b, :3] = p.T
        out[b, 3:] = nrm.T
    return out


def make_features(B, C, N, seed=0):
    """Unit-variance per-point features [B, C, N] (stand-in for the LRF coords + local-PPF MLP channels)."""
    rng = np.random.default_rng(seed + 7919)
    return rng.standard_normal((B, C, N), dtype=np.float32)


def make_pairs(P, N, seed=0, max_angle_deg=360.0, max_trans=0.8, jitter=0.01):
    """Registration pairs: target = random SE(3) of the source + independent jitter
    (/root/reference/utils/open3d_func.py:85-102).  -> (src [P,6,N], tgt [P,6,N], R [P,3,3], t [P,3])."""
    rng = np.random.default_rng(seed + 104729)
    src = make_clouds(P, N, seed=seed, jitter=jitter)
    axis = rng.random((P, 3)); axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    ang = np.deg2rad(rng.uniform(0.0, max_angle_deg, P))
    K = np.zeros((P, 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -axis[:, 2], axis[:, 1], axis[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -axis[:, 0], -axis[:, 1], axis[:, 0]
    R = np.eye(3)[None] + np.sin(ang)[:, None, None] * K + (1 - np.cos(ang))[:, None, None] * (K @ K)
    t = rng.uniform(-max_trans, max_trans, (P, 3))
    tgt = np.empty_like(src)
    tgt[:, :3] = np.einsum('pij,pjn->pin', R, src[:, :3]) + t[:, :, None]
    tgt[:, :3] += np.clip(rng.standard_normal(tgt[:, :3].shape) * jitter, -0.05, 0.05)
    tgt[:, 3:] = np.einsum('pij,pjn->pin', R, src[:, 3:])
    return src, tgt.astype(np.float32), R.astype(np.float32), t.astype(np.float32)


def make_scan(N, seed=0, noise=0.005):
    """ICL-NUIM-shaped indoor scan: points on 6-10 random planes inside a 4 x 3 x 2.5 m room, sigma = 5 mm.
    -> float32 [6, N] (xyz | normal), not centred."""
    rng = np.random.default_rng(seed + 15485863)
    room = np.array([4.0, 3.0, 2.5])
    nplanes = int(rng.integers(6, 11))
    per = np.full(nplanes, N // nplanes); per[: N - per.sum()] += 1
    pts, nrms = [], []
    for q in range(nplanes):
        ax = q % 3 if q < 6 else int(rng.integers(3))
        off = (0.0 if (q // 3) % 2 == 0 else room[ax]) if q < 6 else rng.uniform(0.2, 0.8) * room[ax]
        p = rng.uniform(0, 1, (per[q], 3)) * room
        if q >= 6:                                 # furniture-sized patch
            lo = rng.uniform(0, 0.6, 3) * room
            p = lo + rng.uniform(0, 1, (per[q], 3)) * 0.4 * room
        p[:, ax] = off
        n = np.zeros((per[q], 3)); n[:, ax] = 1.0 if off < room[ax] / 2 else -1.0
        pts.append(p); nrms.append(n)
    p = np.concatenate(pts) + rng.standard_normal((N, 3)) * noise
    n = np.concatenate(nrms)
    perm = rng.permutation(N)
    return np.concatenate([p[perm].T, n[perm].T], 0).astype(np.float32)
