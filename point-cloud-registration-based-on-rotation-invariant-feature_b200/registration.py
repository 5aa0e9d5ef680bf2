"""Registration back end on the GPU: descriptors -> mutual matches -> rigid pose -> RRE / RTE / RMSE, whole batches of
pairs per call (csrc/matcher.cu + csrc/pose.cu).

Mirrors the reference's meter `MeterModelNet40_registration` (/root/reference/datasets/deepgmr_mn40.py:98-244): same
constructor argument, same `update((feat1, feat2), (pt1, pt2, gt_trans))` / `compute()` contract, same thresholds
(:107-109), same result keys (:141).  The reference copies the descriptors to the host and calls Open3D / TEASER++ once
per pair (:121, :165-231); here `func='ransac'` runs the RANSAC configuration the reference hands to Open3D
(utils/open3d_func.py:43-49: voxel_size 0.08, edge-length similarity 0.9, max_iter 1000) on the GPU for all pairs at once,
and `func='kabsch'` the plain least squares over the mutual matches.  FGR, ICP and TEASER++ are third-party solvers the
reference only calls into; they are not re-implemented (func='fgr' / 'icp' / 'teaserpp' raise).
"""
import time

import torch

from . import matcher

__all__ = ['estimate_poses', 'registration_metrics', 'register_pairs', 'MeterModelNet40_registration']


def estimate_poses(src, tgt, idx1, idx2, count, func='ransac', voxel_size=0.08, max_iter=1000, edge_similarity=0.9,
                   refine_iters=3, seed=0):
    """src [P,n1,3], tgt [P,n2,3] CUDA fp32; idx1/idx2/count: the mutual matches (matcher.mutual_nn).
    -> (T [P,4,4] fp32 mapping src onto tgt, inliers [P] int32)."""
    if func not in ('ransac', 'kabsch'):
        raise ValueError("func must be 'ransac' or 'kabsch' (Open3D FGR / ICP and TEASER++ are third-party solvers)")
    hyps = int(max_iter) if func == 'ransac' else 0
    return torch.ops.ri.pose_from_matches(src.float().contiguous(), tgt.float().contiguous(), idx1.contiguous(),
                                          idx2.contiguous(), count.contiguous(), hyps, float(voxel_size),
                                          float(edge_similarity), int(refine_iters), int(seed))


def registration_metrics(gt_trans, est_trans, pts):
    """[P,4,4], [P,4,4], [P,n,3] -> [P,3] float64 (RRE in degrees, RTE, RMSE) — RE_TE_one_pair + the RMSE of update()."""
    return torch.ops.ri.registration_metrics(gt_trans.float().contiguous(), est_trans.float().contiguous(),
                                             pts.float().contiguous())


def register_pairs(feat1, feat2, pt1, pt2, func='ransac', **kw):
    """feat1/feat2 [P,C,n] (the extractor's layout), pt1/pt2 [P,n,3], all CUDA -> (T [P,4,4], inliers [P], matches dict)."""
    m = matcher.mutual_nn(feat1, feat2, point_major=False, want_dist=False)      # the meter uses the matches only
    T, inl = estimate_poses(pt1, pt2, m['idx1'], m['idx2'], m['count'], func=func, **kw)
    return T, inl, m


class MeterModelNet40_registration:
    """Same state, thresholds and result keys as the reference meter (deepgmr_mn40.py:98-141)."""

    def __init__(self, func='ransac', device='cuda'):
        self.rre = 0
        self.rte = 0
        self.num = 0
        self.succ = 0
        self.rmse = 0
        self.rmse_succ = 0
        self.reg_time = 0
        self.rot_thresh = 1e-05
        self.rmse_thresh = 0.2
        self.translate_thresh = 0.005
        self.func = func
        self.device = torch.device(device)

    def update(self, output, target):
        """output = (feat1, feat2) [b,C,n]; target = (pt1 [b,n,3], pt2 [b,n,3], gt_trans [b,4,4]) — numpy or torch."""
        with torch.no_grad():
            dev = self.device
            feat1, feat2 = (torch.as_tensor(f, dtype=torch.float32).to(dev) for f in output)
            pt1, pt2, gt = (torch.as_tensor(t, dtype=torch.float32).to(dev) for t in target)
            torch.cuda.synchronize(dev)
            t0 = time.time()
            T, _, _ = register_pairs(feat1, feat2, pt1, pt2, func=self.func)
            torch.cuda.synchronize(dev)
            reg_time = time.time() - t0
            m = registration_metrics(gt, T, pt1).cpu()
            rre, rte, rmse = m[:, 0], m[:, 1], m[:, 2]
            b = pt1.shape[0]
            self.succ += int(((rre < self.rot_thresh) & (rte < self.translate_thresh)).sum())
            self.rmse_succ += int((rmse < self.rmse_thresh).sum())
            self.rre += float(rre.sum())
            self.rte += float(rte.sum())
            self.rmse += float(rmse.sum())
            self.reg_time += reg_time
            self.num += b

    def compute(self):
        return {'succ': self.succ / self.num, 'rre': self.rre / self.num, 'rte': self.rte / self.num,
                'rmse': self.rmse / self.num, 'reg_time': self.reg_time / self.num, 'rmse_succ': self.rmse_succ / self.num}
