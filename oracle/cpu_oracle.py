"""ctypes front end of oracle/ri_oracle.c — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this module.  The product package (ri_b200) never does; it fails loudly when its CUDA library is missing.

Each wrapper takes/returns C-contiguous numpy arrays with the reference's layouts ([B,C,N], points
innermost) and cites the reference code the C function restates (paths relative to /root/reference).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ri_oracle.c")
LIB = os.path.join(HERE, "libri_oracle.so")

_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off: the C file spells every fma itself, the compiler must add none."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off",
                               "-fno-fast-math", "-fopenmp", SRC, "-o", LIB, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def knn(xyz1, xyz2, k):
    """Bilateral kNN as knn_forward_cuda does (knn/knn.cpp:6-25, knn/knn.cu:81-87): returns
    dist1[B,k,n], dist2[B,k,m], idx1, idx2."""
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    B, c, n = xyz1.shape
    m = xyz2.shape[2]
    d1 = np.empty((B, k, n), np.float32); i1 = np.empty((B, k, n), np.int32)
    d2 = np.empty((B, k, m), np.float32); i2 = np.empty((B, k, m), np.int32)
    lib().ri_oracle_knn(_p(xyz1), _p(xyz2), B, c, n, m, k, _p(d1), _p(i1))
    lib().ri_oracle_knn(_p(xyz2), _p(xyz1), B, c, m, n, k, _p(d2), _p(i2))
    return d1, d2, i1, i2


def knn_one(xyz1, xyz2, k):
    """One direction only (queries xyz1, references xyz2)."""
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    B, c, n = xyz1.shape
    m = xyz2.shape[2]
    d1 = np.empty((B, k, n), np.float32); i1 = np.empty((B, k, n), np.int32)
    lib().ri_oracle_knn(_p(xyz1), _p(xyz2), B, c, n, m, k, _p(d1), _p(i1))
    return d1, i1


def knn_grad(xyz1, xyz2, gd1, gd2, idx1, idx2):
    """knn_backward_cuda (knn/knn.cpp:27-52, knn/knn.cu:89-97)."""
    xyz1, xyz2, gd1, gd2, idx1, idx2 = _f(xyz1), _f(xyz2), _f(gd1), _f(gd2), _i(idx1), _i(idx2)
    B, c, n = xyz1.shape
    m = xyz2.shape[2]
    k = idx1.shape[1]
    g1 = np.zeros((B, c, n), np.float32); g2 = np.zeros((B, c, m), np.float32)
    lib().ri_oracle_knn_grad(_p(xyz1), _p(xyz2), _p(gd1), _p(idx1), B, c, n, m, k, _p(g1), _p(g2))
    lib().ri_oracle_knn_grad(_p(xyz2), _p(xyz1), _p(gd2), _p(idx2), B, c, m, n, k, _p(g2), _p(g1))
    return g1, g2


def ppf_backend(coords, center, normals, center_normal):
    """_backend.spherical_ppf_forward argument order (spherical_ppf/ppf.cpp:17-36)."""
    coords, center, normals, center_normal = _f(coords), _f(center), _f(normals), _f(center_normal)
    B, _, L = coords.shape
    feat = np.zeros((B, 4, L), np.float32)
    lib().ri_oracle_ppf(_p(coords), _p(center), _p(normals), _p(center_normal), B, L, _p(feat))
    return feat


def ppf(centers_coords, points_coords, centers_normals, points_normals):
    """functional/ppf.py:8-22 — note the argument swap on the way to the backend (ppf.py:22)."""
    return ppf_backend(points_coords, centers_coords, points_normals, centers_normals)


def sph_grid_stats(coords, r):
    coords = _f(coords)
    B, _, N = coords.shape
    ind = np.empty((B, N), np.int32); cnt = np.empty((B, r ** 3), np.int32)
    lib().ri_oracle_sph_grid_stats(_p(coords), B, N, r, _p(ind), _p(cnt))
    return ind, cnt


def sph_grid_cont(coords, r):
    coords = _f(coords)
    B, _, N = coords.shape
    gc = np.empty((B, N, 3), np.float64)
    lib().ri_oracle_sph_grid_cont(_p(coords), B, N, r, _p(gc))
    return gc


def cube_grid_stats(vox_coords, r):
    vox_coords = _i(vox_coords)
    B, _, N = vox_coords.shape
    ind = np.empty((B, N), np.int32); cnt = np.empty((B, r ** 3), np.int32)
    lib().ri_oracle_cube_grid_stats(_p(vox_coords), B, N, r, _p(ind), _p(cnt))
    return ind, cnt


def scatter_mean(features, ind, cnt, r):
    features, ind, cnt = _f(features), _i(ind), _i(cnt)
    B, C, N = features.shape
    s = r ** 3
    out = np.empty((B, C, s), np.float32)
    lib().ri_oracle_avg_voxelize(_p(features), _p(ind), _p(cnt), B, C, N, s, _p(out))
    return out


def spherical_avg_voxelize(features, coords, r):
    """_backend.spherical_avg_voxelize_forward (spherical_voxelization/spherical_vox.cpp:17-46): (out, ind, cnt)."""
    ind, cnt = sph_grid_stats(coords, r)
    return scatter_mean(features, ind, cnt, r), ind, cnt


def avg_voxelize(features, vox_coords, r):
    """_backend.avg_voxelize_forward (voxelization/vox.cpp:17-43): (out, ind, cnt)."""
    ind, cnt = cube_grid_stats(vox_coords, r)
    return scatter_mean(features, ind, cnt, r), ind, cnt


def avg_voxelize_grad(grad_y, ind, cnt):
    """avg_voxelize_backward == spherical_avg_voxelize_backward (vox.cpp:54-78)."""
    grad_y, ind, cnt = _f(grad_y), _i(ind), _i(cnt)
    B, C, s = grad_y.shape
    N = ind.shape[1]
    gx = np.empty((B, C, N), np.float32)
    lib().ri_oracle_avg_voxelize_grad(_p(grad_y), _p(ind), _p(cnt), B, C, N, s, _p(gx))
    return gx


def trilinear_devoxelize(coords, features, r):
    """_backend.trilinear_devoxelize_forward (interpolate/trilinear_devox.cpp:18-55): (outs, inds, wgts)."""
    coords, features = _f(coords), _f(features)
    B, C = features.shape[:2]
    features = features.reshape(B, C, -1)
    N = coords.shape[2]
    outs = np.empty((B, C, N), np.float32); inds = np.empty((B, 8, N), np.int32); wgts = np.empty((B, 8, N), np.float32)
    lib().ri_oracle_trilinear_devox(_p(coords), _p(features), B, C, N, r, _p(outs), _p(inds), _p(wgts))
    return outs, inds, wgts


def spherical_trilinear_devoxelize(coords, features, g_inds, r):
    """_backend.spherical_trilinear_devoxelize_forward (interpolate/spherical_trilinear_devox.cpp:19-56)."""
    coords, features, g_inds = _f(coords), _f(features), _i(g_inds)
    B, C = features.shape[:2]
    features = features.reshape(B, C, -1)
    N = coords.shape[2]
    outs = np.empty((B, C, N), np.float32); inds = np.empty((B, 8, N), np.int32); wgts = np.empty((B, 8, N), np.float32)
    lib().ri_oracle_sph_trilinear_devox(_p(coords), _p(features), _p(g_inds), B, C, N, r, _p(outs), _p(inds), _p(wgts))
    return outs, inds, wgts


def devox_grad(grad_y, inds, wgts, r, spherical):
    grad_y, inds, wgts = _f(grad_y), _i(inds), _f(wgts)
    B, C, N = grad_y.shape
    s = r ** 3
    gx = np.empty((B, C, s), np.float32)
    lib().ri_oracle_devox_grad(_p(grad_y), _p(inds), _p(wgts), B, C, N, s, int(bool(spherical)), _p(gx))
    return gx


def voxel_edge_gather(avg, features, inds):
    """PVCNN/modules/pvconv.py:68-90 -> [B,2C,N]."""
    features, inds = _f(features), _i(inds)
    B, C, N = features.shape
    avg = _f(avg).reshape(B, C, -1)
    s = avg.shape[2]
    out = np.empty((B, 2 * C, N), np.float32)
    lib().ri_oracle_voxel_edge_gather(_p(avg), _p(features), _p(inds), B, C, N, s, _p(out))
    return out


def ball_query(centers, points, radius, num_neighbors):
    """ball_query/ball_query.cu:19-50: centers [B,3,M], points [B,3,N] -> int32 [B,M,U]."""
    centers, points = _f(centers), _f(points)
    B, _, M = centers.shape
    N = points.shape[2]
    out = np.empty((B, M, num_neighbors), np.int32)
    lib().ri_oracle_ball_query.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    lib().ri_oracle_ball_query(_p(centers), _p(points), B, N, M, float(radius), int(num_neighbors), _p(out))
    return out


def grouping(features, indices):
    """grouping/grouping.cu:18-44: features [B,C,N], indices [B,M,U] -> [B,C,M,U]."""
    features, indices = _f(features), _i(indices)
    B, C, N = features.shape
    _, M, U = indices.shape
    out = np.empty((B, C, M, U), np.float32)
    lib().ri_oracle_grouping(_p(features), _p(indices), B, C, N, M, U, _p(out))
    return out


def grouping_grad(grad_y, indices, N):
    """grouping/grouping.cu:58-84: grad_y [B,C,M,U], indices [B,M,U] -> grad_x [B,C,N]."""
    grad_y, indices = _f(grad_y), _i(indices)
    B, C, M, U = grad_y.shape
    out = np.empty((B, C, N), np.float32)
    lib().ri_oracle_grouping_grad(_p(grad_y), _p(indices), B, C, N, M, U, _p(out))
    return out


def find_correspondence_one_pair(feat1, feat2):
    """datasets/deepgmr_mn40.py:232-244, restated with the same numpy calls (fp32 in, fp32 sgemm)."""
    diff = (np.power(np.linalg.norm(feat1, axis=1, keepdims=True), 2)
            + np.power(np.linalg.norm(feat2, axis=1, keepdims=True).T, 2)
            - 2 * np.dot(feat1, feat2.T))
    c1 = np.argmin(diff, axis=1)
    c2 = np.argmin(diff, axis=0)
    mask = (c2[c1] == np.arange(c1.shape[0]))
    return np.arange(c1.shape[0])[mask], c1[mask], diff


# ---- module-level prologues (torch-free numpy restatements; fp32 throughout) -------------------------

def spherical_voxelization_module(features, coords, r):
    """Spherical_Voxelization.forward (PVCNN/modules/spherical_vox.py:14-23).  NOTE: numpy's mean/norm
    reductions are not bit-identical to torch's; parity tests feed the SAME norm_coords to both sides."""
    coords = _f(coords)
    nc = coords - coords.mean(2, keepdims=True, dtype=np.float32)
    nrm = np.sqrt((nc * nc).sum(1, keepdims=True, dtype=np.float32)).max(2, keepdims=True)
    nc = (nc / (nrm + np.float32(1e-20))).astype(np.float32)
    out, ind, _ = spherical_avg_voxelize(features, nc, r)
    return out, ind, nc


def voxelization_module(features, coords, r, normalize=True, eps=0.0):
    """Voxelization.forward (PVCNN/modules/voxelization.py:16-35)."""
    coords = _f(coords)
    nc = coords - coords.mean(2, keepdims=True, dtype=np.float32)
    if normalize:
        nrm = np.sqrt((nc * nc).sum(1, keepdims=True, dtype=np.float32)).max(2, keepdims=True)
        nc = nc / (nrm * np.float32(2.0) + np.float32(eps)) + np.float32(0.5)
    else:
        nc = (nc + np.float32(1)) / np.float32(2.0)
    nc = np.clip(nc * np.float32(r), 0, r - 1).astype(np.float32)
    vox = np.rint(nc).astype(np.int32)                     # torch.round == half-to-even == rint
    out, ind, _ = avg_voxelize(features, vox, r)
    return out, ind, nc


def grid_subsample(points, features=None, labels=None, grid_size=0.1):
    """utils/grid_subsampleing.py:3-21 -> cpp_subsampling grid_subsampling.cpp:4-106.  points (N,3), features (N,d),
    labels (N,) or (N,l) -> (sub_points, sub_features, sub_labels, cell_keys), cells in ascending cell index."""
    pts = _f(points)
    N = pts.shape[0]
    f = None if features is None else _f(features).reshape(N, -1)
    l = None if labels is None else _i(labels).reshape(N, -1)
    fdim = 0 if f is None else f.shape[1]
    ldim = 0 if l is None else l.shape[1]
    op = np.empty((max(N, 1), 3), np.float32); of = np.empty((max(N, 1), max(fdim, 1)), np.float32)
    ol = np.empty((max(N, 1), max(ldim, 1)), np.int32); keys = np.empty(max(N, 1), np.uint64)
    fn = lib().ri_oracle_grid_subsample
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_float] + [ctypes.c_void_p] * 4
    M = fn(_p(pts), None if f is None else _p(f), None if l is None else _p(l), N, fdim, ldim, float(grid_size),
           _p(op), _p(of), _p(ol), _p(keys))
    return op[:M].copy(), (of[:M, :fdim].copy() if fdim else None), (ol[:M, :ldim].copy() if ldim else None), keys[:M].copy()


# ---------------------------------------------------------------------------------------------- registration (row f3)
def re_te_one_pair(gt, est):
    """RE_TE_one_pair, datasets/deepgmr_mn40.py:152-164 (numpy, as the reference): rotation error in degrees, |dt|."""
    import math
    gt_R = gt[:3, :3]
    est_R = est[:3, :3]
    A = (np.trace(np.dot(gt_R.T, est_R)) - 1) / 2
    if A > 1:
        A = 1
    elif A < -1:
        A = -1
    return math.degrees(math.fabs(math.acos(A))), np.linalg.norm(gt[:3, 3] - est[:3, 3])


def apply_transform_2dim_numpy(pts, trans):
    """utils/open3d_func.py:104-110."""
    return pts.dot(trans[:3, :3].T) + trans[:3, 3][np.newaxis, :]


def registration_metrics(gt, est, pts):
    """Per pair (rre, rte, rmse) as MeterModelNet40_registration.update computes them (deepgmr_mn40.py:119-126).
    gt, est [P,4,4], pts [P,n,3] -> [P,3] float64."""
    out = np.zeros((len(gt), 3))
    for i in range(len(gt)):
        rre, rte = re_te_one_pair(gt[i], est[i])
        d = apply_transform_2dim_numpy(pts[i], est[i]) - apply_transform_2dim_numpy(pts[i], gt[i])
        out[i] = (rre, rte, np.mean(np.linalg.norm(d, axis=1)))
    return out


def kabsch(a, b):
    """Least-squares rigid transform mapping a [m,3] onto b [m,3] (rotation + translation, no scale) — the estimator the
    reference selects in Open3D (TransformationEstimationPointToPoint(False), utils/open3d_func.py:46), by SVD in fp64."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    ca, cb = a.mean(0), b.mean(0)
    H = (a - ca).T @ (b - cb)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ D @ U.T
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = cb - R @ ca
    return T


def count_inliers(T, a, b, dist):
    d = apply_transform_2dim_numpy(np.asarray(a, np.float64), np.asarray(T, np.float64)) - b
    return int((np.linalg.norm(d, axis=1) < dist).sum())


# ---------------------------------------------------------------------------------------------- LRF change_coords (row f4)
def change_coords(coords, mean=None):
    """PVCNN/models/pvcnn_classify.py:153-184, float32 numpy, loop for loop.  coords [B,3,N] -> (new [B,3,N], ok [B]).
    `mean` [B,3]: the per-cloud mean to use (pass torch's to see the same bits); default numpy's."""
    f32 = np.float32
    coords = _f(coords)
    b, _, n = coords.shape
    m = coords.mean(axis=2, dtype=f32) if mean is None else _f(mean)
    norm_coords = coords - m[:, :, None]                                            # :154
    nrm = lambda v: f32(np.sqrt(f32(f32(f32(v[0] * v[0]) + f32(v[1] * v[1])) + f32(v[2] * v[2]))))
    radius = np.sqrt((norm_coords[:, 0] ** 2 + norm_coords[:, 1] ** 2) + norm_coords[:, 2] ** 2).astype(f32)
    rank = np.argsort(-radius, axis=1, kind="stable")                               # :155 (ties: lowest index first)
    out = np.zeros((b, 3, n), f32); ok = np.zeros(b, np.int32)
    for i in range(b):
        base_x = norm_coords[i, :, rank[i, 0]]                                      # :159
        if not nrm(base_x) > 1e-5:                                                  # :160 assert
            continue
        base_x = base_x / nrm(base_x)
        found = False
        for j in range(1, n):                                                       # :162-169
            base_y = norm_coords[i, :, rank[i, j]]
            if nrm(base_y) < 1e-5:
                continue
            base_y = base_y / nrm(base_y)
            lamda = f32(f32(f32(base_x[0] * base_y[0]) + f32(base_x[1] * base_y[1])) + f32(base_x[2] * base_y[2]))
            if lamda < 0.9 and lamda > -0.9:
                found = True
                break
        if not found:                                                               # :170 assert
            continue
        d = f32(f32(f32(base_x[0] * base_y[0]) + f32(base_x[1] * base_y[1])) + f32(base_x[2] * base_y[2]))
        base_x = (base_x - base_y * d).astype(f32)                                  # :175
        if nrm(base_x) < 1e-5:                                                      # :176 assert
            continue
        base_x = base_x / nrm(base_x)                                               # :177
        base_z = np.cross(base_x, base_y).astype(f32)                               # :179
        base_z = base_z / nrm(base_z)                                               # :180
        out[i, 0] = base_x @ norm_coords[i]; out[i, 1] = base_y @ norm_coords[i]; out[i, 2] = base_z @ norm_coords[i]   # :181-184
        ok[i] = 1
    return out, ok


# ---------------------------------------------------------------------------------------------- local PPF (row f1, fused)
def local_ppf(points_coords, points_normals, neighbors, centers_coords=None, centers_normals=None):
    """PVCNN/models/pvcnn_classify.py:252-270 on top of PVCNN/modules/ball_query.py:16-35, float32 numpy, operation for
    operation: rel = p_nbr - c; d = c - rel; dn = sqrt((d0^2 + d1^2) + d2^2); du = d / dn; three clamped acos of dot products
    summed as (p0 + p1) + p2.  points_* [B,3,N], neighbors [B,M,U] -> [B,4,U,M]."""
    f32 = np.float32
    pc, pn = _f(points_coords), _f(points_normals)
    cc = pc if centers_coords is None else _f(centers_coords)
    cn = pn if centers_normals is None else _f(centers_normals)
    nb = _i(neighbors)
    B, M, U = nb.shape
    gi = np.broadcast_to(nb.reshape(B, 1, M * U), (B, 3, M * U))
    q = np.take_along_axis(pc, gi, 2).reshape(B, 3, M, U)            # F.grouping(points_coords, idx)
    r = np.take_along_axis(pn, gi, 2).reshape(B, 3, M, U)
    c = cc[:, :, :, None]; n = cn[:, :, :, None]
    rel = (q - c).astype(f32)                                        # ball_query.py:24
    d = (c - rel).astype(f32)                                        # pvcnn_classify.py:262
    sq = (d * d).astype(f32)
    dn = np.sqrt(((sq[:, 0] + sq[:, 1]).astype(f32) + sq[:, 2]).astype(f32)).astype(f32)      # :263
    with np.errstate(divide="ignore", invalid="ignore"):
        du = (d / dn[:, None]).astype(f32)                           # :264
        dot = lambda a, b: (((a[:, 0] * b[:, 0]).astype(f32) + (a[:, 1] * b[:, 1]).astype(f32)).astype(f32)
                            + (a[:, 2] * b[:, 2]).astype(f32)).astype(f32)
        nb_ = np.broadcast_to(n, r.shape)
        ac = lambda x: np.arccos(np.clip(x, f32(-1), f32(1))).astype(f32)
        out = np.stack([ac(dot(r, du)), ac(dot(nb_, du)), ac(dot(r, nb_)), dn], 1)           # [B,4,M,U]  (:265-268)
    return np.ascontiguousarray(out.transpose(0, 1, 3, 2))           # (b, 4, k, m)
