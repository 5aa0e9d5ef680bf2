"""Generate golden input/output vectors from the UNMODIFIED reference CUDA backend (oracle/_ref).

The reference has no CPU path (every op CHECK_CUDAs, src/utils.hpp:15) and this build container has no GPU, so
the fixtures are produced on the GPU box:

    gpurun -- python oracle/make_golden.py            # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/          # here, then commit

Inputs are small and seeded; adversarial cases follow SURVEY.md §8c (duplicate points / distance ties, query ==
reference, m < k, a point at the centroid, points on x = 0 and x = y = 0, the farthest point (gamma >= 1),
zero-length normals, parallel normals, coordinates on cell boundaries).  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def cloud(rng, B, N, kind="mixed"):
    x = rng.standard_normal((B, 3, N)).astype(np.float32)
    if kind == "surface":
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        x *= rng.uniform(0.4, 1.0, (B, 3, 1)).astype(np.float32)
        x += (rng.standard_normal((B, 3, N)) * 0.01).astype(np.float32)
    return x.astype(np.float32)


def norm_coords_torch(torch, coords):
    """Spherical_Voxelization.forward prologue (PVCNN/modules/spherical_vox.py:16-19), torch on the GPU."""
    nc = coords - coords.mean(2, keepdim=True)
    return nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values + 1e-20)


def main(out_dir):
    import torch
    from oracle.build_ref import load_ref
    ref = load_ref()
    if ref is None or not torch.cuda.is_available():
        raise SystemExit("needs oracle/_ref/_multi_shape_pvcnn_backend.so and a GPU")
    os.makedirs(out_dir, exist_ok=True)
    dev = "cuda"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    A = lambda t: t.detach().cpu().numpy()
    rng = np.random.default_rng(20261018)

    # ---------------------------------------------------------------- KNN (+ backward)
    x1 = cloud(rng, 3, 256); x2 = cloud(rng, 3, 200)
    x2[:, :, 50:60] = x2[:, :, 40:50]            # duplicated references -> exact distance ties
    x2[:, :, :16] = x1[:, :, :16]                # query == reference (d = 0)
    x1[0, :, 100] = 200.0                        # a query farther than sqrt(10000) from everything
    k = 8
    d1, d2, i1, i2 = ref.knn_forward_cuda(T(x1), T(x2), k)
    gd1 = rng.standard_normal(A(d1).shape).astype(np.float32)
    gd2 = rng.standard_normal(A(d2).shape).astype(np.float32)
    gd1[0, 0, :5] = 20000.0                      # hits the `2g >= 20000` skip (knn.cu:68)
    g1, g2 = ref.knn_backward_cuda(T(x1), T(x2), T(gd1), T(gd2), i1, i2)
    np.savez_compressed(os.path.join(out_dir, "knn.npz"), xyz1=x1, xyz2=x2, k=k, dist1=A(d1), dist2=A(d2),
                        idx1=A(i1), idx2=A(i2), graddist1=gd1, graddist2=gd2, gradxyz1=A(g1), gradxyz2=A(g2))
    # m < k and a self query with k = 20 on a surface-like cloud
    xs = cloud(rng, 2, 5)
    xq = cloud(rng, 2, 40)
    e1, e2, j1, j2 = ref.knn_forward_cuda(T(xq), T(xs), 8)
    xc = cloud(rng, 2, 512, "surface")
    s1, _, sj1, _ = ref.knn_forward_cuda(T(xc), T(xc), 20)
    x5 = rng.standard_normal((2, 5, 96)).astype(np.float32)     # generic channel count c = 5
    y5 = rng.standard_normal((2, 5, 64)).astype(np.float32)
    f1, f2, fj1, fj2 = ref.knn_forward_cuda(T(x5), T(y5), 4)
    np.savez_compressed(os.path.join(out_dir, "knn_edge.npz"), xq=xq, xs=xs, dist1=A(e1), dist2=A(e2), idx1=A(j1),
                        idx2=A(j2), xc=xc, self_dist=A(s1), self_idx=A(sj1),
                        x5=x5, y5=y5, c5_dist1=A(f1), c5_dist2=A(f2), c5_idx1=A(fj1), c5_idx2=A(fj2))

    # ---------------------------------------------------------------- PPF
    L = 384
    pc = cloud(rng, 2, L); cc = cloud(rng, 2, L)
    pn = cloud(rng, 2, L); cn = cloud(rng, 2, L)
    cc[:, :, :8] = pc[:, :, :8]                  # d = 0 columns
    pn[:, :, 8:16] = 0.0                         # zero-length point normals
    cn[:, :, 16:24] = 0.0                        # zero-length centre normals
    cn[:, :, 24:40] = pn[:, :, 24:40] * 3.0      # parallel normals (acos near 0)
    cn[:, :, 40:56] = -pn[:, :, 40:56]           # anti-parallel normals (acos near pi)
    pn[:, :, 56:64] = (cc - pc)[:, :, 56:64]     # normal parallel to the offset
    feat = ref.spherical_ppf_forward(T(pc), T(cc), T(pn), T(cn))
    np.savez_compressed(os.path.join(out_dir, "ppf.npz"), coords=pc, center=cc, normals=pn, center_normal=cn, feat=A(feat))

    # ---------------------------------------------------------------- spherical voxelize / devox (+ backward)
    sph = {}
    for r in (4, 8, 16, 32):
        B, N, C = 2, 320, 5
        raw = cloud(rng, B, N, "surface")
        raw[:, 0, 10:20] = 0.0                    # x == 0, y != 0
        raw[:, 0, 20:24] = 0.0; raw[:, 1, 20:24] = 0.0      # x == y == 0 (on the pole axis)
        nc = A(norm_coords_torch(torch, T(raw)))
        nc[:, :, 30] = 0.0                        # gamma == 0
        nc[0, :, 31] = np.array([0.0, 0.0, 0.5], np.float32)     # exactly on the +z axis
        nc[0, :, 32] = np.array([0.0, 0.0, -0.5], np.float32)    # exactly on the -z axis (beta = pi -> undefined)
        nc[1, :, 33] = np.array([0.5, 0.0, 0.0], np.float32)     # cell-boundary-ish values
        nc[1, :, 34] = np.array([-0.25, 0.0, 0.0], np.float32)
        nc[1, :, 35] = np.array([0.0, -0.75, 0.0], np.float32)
        feats = rng.standard_normal((B, C, N)).astype(np.float32)
        out, ind, cnt = ref.spherical_avg_voxelize_forward(T(feats), T(nc), r)
        gy = rng.standard_normal(A(out).shape).astype(np.float32)
        gx = ref.spherical_avg_voxelize_backward(T(gy), ind, cnt)
        grid = rng.standard_normal((B, 7, r ** 3)).astype(np.float32)
        douts, dinds, dwgts = ref.spherical_trilinear_devoxelize_forward(r, True, T(nc), T(grid), ind)
        dgy = rng.standard_normal(A(douts).shape).astype(np.float32)
        dgx = ref.spherical_trilinear_devoxelize_backward(T(dgy), dinds, dwgts, r)
        sph.update({f"r{r}_coords": nc, f"r{r}_feat": feats, f"r{r}_out": A(out), f"r{r}_ind": A(ind), f"r{r}_cnt": A(cnt),
                    f"r{r}_gy": gy, f"r{r}_gx": A(gx), f"r{r}_grid": grid, f"r{r}_douts": A(douts),
                    f"r{r}_dinds": A(dinds), f"r{r}_dwgts": A(dwgts), f"r{r}_dgy": dgy, f"r{r}_dgx": A(dgx)})
    np.savez_compressed(os.path.join(out_dir, "spherical.npz"), **sph)

    # ---------------------------------------------------------------- cube voxelize / devox (+ backward)
    cube = {}
    for r in (4, 8, 16):
        B, N, C = 2, 320, 5
        raw = cloud(rng, B, N, "surface")
        t = T(raw)
        nc = t - t.mean(2, keepdim=True)
        nc = (nc + 1) / 2.0                        # Voxelization(normalize=False) (voxelization.py:24)
        nc = torch.clamp(nc * r, 0, r - 1)
        nc[:, :, :6] = torch.tensor([0.0, 0.5, 1.5, 2.5, r - 1.0, r - 1.5], device=dev)   # ties for round-half-even
        vox = torch.round(nc).to(torch.int32)
        feats = rng.standard_normal((B, C, N)).astype(np.float32)
        out, ind, cnt = ref.avg_voxelize_forward(T(feats), vox.contiguous(), r)
        gy = rng.standard_normal(A(out).shape).astype(np.float32)
        gx = ref.avg_voxelize_backward(T(gy), ind, cnt)
        grid = rng.standard_normal((B, 7, r ** 3)).astype(np.float32)
        douts, dinds, dwgts = ref.trilinear_devoxelize_forward(r, True, nc.contiguous(), T(grid))
        dgy = rng.standard_normal(A(douts).shape).astype(np.float32)
        dgx = ref.trilinear_devoxelize_backward(T(dgy), dinds, dwgts, r)
        cube.update({f"r{r}_norm_coords": A(nc), f"r{r}_vox": A(vox), f"r{r}_feat": feats, f"r{r}_out": A(out),
                     f"r{r}_ind": A(ind), f"r{r}_cnt": A(cnt), f"r{r}_gy": gy, f"r{r}_gx": A(gx), f"r{r}_grid": grid,
                     f"r{r}_douts": A(douts), f"r{r}_dinds": A(dinds), f"r{r}_dwgts": A(dwgts), f"r{r}_dgy": dgy,
                     f"r{r}_dgx": A(dgx)})
    np.savez_compressed(os.path.join(out_dir, "cube.npz"), **cube)
    # ---------------------------------------------------------------- ball query + grouping (+ backward)
    bq = {}
    pts = cloud(rng, 3, 400, "surface"); ctr = pts[:, :, :150].copy()
    ctr[:, :, 140:] += 5.0                        # centres with no neighbour at all -> zero rows
    pts[:, :, 300:310] = pts[:, :, 0:10]          # duplicates of centres 0..9: d2 = 0 excluded like the centre itself
    pts[:, :, 310] = pts[:, :, 11] + np.float32(0.002)      # d2 ~ 1.2e-5: just above the 1e-5 lower bound
    pts[:, :, 311] = pts[:, :, 12] + np.float32(0.0015)     # d2 ~ 6.75e-6: just below it
    for name, (radius, u) in {"a": (0.3, 16), "b": (0.15, 64), "c": (2.0, 8)}.items():
        idx = ref.ball_query(T(ctr), T(pts), radius, u)
        feats = rng.standard_normal((3, 6, 400)).astype(np.float32)
        grp = ref.grouping_forward(T(feats), idx)
        gy = rng.standard_normal(A(grp).shape).astype(np.float32)
        gx = ref.grouping_backward(T(gy), idx, 400)
        bq.update({f"{name}_radius": radius, f"{name}_u": u, f"{name}_idx": A(idx), f"{name}_feat": feats,
                   f"{name}_grp": A(grp), f"{name}_gy": gy, f"{name}_gx": A(gx)})
    np.savez_compressed(os.path.join(out_dir, "ball_query.npz"), centers=ctr, points=pts, **bq)
    print("golden vectors written to", out_dir, sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
