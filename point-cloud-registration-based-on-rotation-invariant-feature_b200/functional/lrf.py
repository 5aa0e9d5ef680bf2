"""change_coords — the rotation-invariant preprocessing of the shipped models (`rot_invariant_preprocess='change_coords'`,
/root/reference/PVCNN/models/pvcnn_classify.py:153-184) as one kernel instead of a Python loop over clouds and points
with a host synchronisation per step."""
import torch

__all__ = ['change_coords', 'global_ppf']


def change_coords(coords, return_bases=False, check=True):
    """coords FloatTensor[B,3,N] (or the [B,6,N] xyz|normal batch) -> new coordinates FloatTensor[B,3,N] in each cloud's own
    frame (x: farthest point from the centroid; y: the farthest point with |cos| < 0.9 to it; z = x × y after Gram-Schmidt).
    `check=True` keeps the reference's asserts (:160, :170, :176) as one AssertionError after the kernel (one host sync);
    `check=False` leaves failing clouds as zeros and never synchronises."""
    coords = coords.float().contiguous()
    mean = coords[:, :3, :].mean(dim=2)                      # torch's own reduction: defines the bits the ranking sees
    out, bases, ok = torch.ops.ri.lrf_change_coords(coords, mean.contiguous())
    if check:
        assert bool((ok == 1).all()), "change_coords: no admissible base_x / base_y for some cloud (pvcnn_classify.py:160,170,176)"
    return (out, bases) if return_bases else out


def global_ppf(coords, normals):
    """The 'extra_feature_channels == 4' features of pvcnn_classify.py:200-204 (and the 'ppf' branch :114-116): PPF of every
    point against the cloud's centroid and mean normal.  coords, normals [B,3,N] -> [B,4,N]."""
    from .ppf import ppf
    n = coords.shape[2]
    centers_coords = coords.mean(dim=2, keepdim=True).expand(-1, -1, n)
    centers_normals = normals.mean(dim=2, keepdim=True).expand(-1, -1, n)
    return ppf(centers_coords, coords, centers_normals, normals)
