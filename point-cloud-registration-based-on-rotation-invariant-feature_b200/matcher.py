"""Mutual-nearest-neighbour descriptor matching — the registration step that follows the feature extractor.

Mirrors `MeterModelNet40_registration.find_correspondence_one_pair`
(/root/reference/datasets/deepgmr_mn40.py:232-244; same code in deepgmr_partial.py:335-347, mn40_hdf.py:466-477):
    diff = |f1|^2 + |f2|^2^T - 2 f1 f2^T ; c1 = argmin(diff, 1) ; c2 = argmin(diff, 0) ; mask = c2[c1] == arange(n1)
but runs on the GPU for a whole batch of pairs in one call (tcgen05 split-precision contraction with the argmins fused
into the epilogue, csrc/matcher.cu), so the `.cpu().numpy()` hop of deepgmr_mn40.py:88 disappears.
"""
import numpy as np
import torch

__all__ = ['mutual_nn', 'find_correspondence_one_pair', 'MutualMatcher']


def mutual_nn(desc1, desc2, point_major=False, want_dist=True):
    """desc1 [P,C,n1], desc2 [P,C,n2] CUDA fp32 (or [P,n,C] with point_major=True).
    Returns a dict: corr12 [P,n1], corr21 [P,n2] (int32 argmins), dist12 [P,n1] (fp32 distance of each row's nearest; absent with
    want_dist=False — the reference method returns the matches only, and without the distances the library skips a pass over the
    descriptors and, for channel-major input, the re-tiled image), idx1/idx2 [P,n1] (mutual matches in ascending idx1, -1 padded)
    and count [P]."""
    d1, d2 = desc1.float().contiguous(), desc2.float().contiguous()
    if not want_dist:
        c12, c21, i1, i2, cnt = torch.ops.ri.mutual_nn_indices(d1, d2, bool(point_major))
        return {'corr12': c12, 'corr21': c21, 'idx1': i1, 'idx2': i2, 'count': cnt}
    c12, c21, d12, i1, i2, cnt = torch.ops.ri.mutual_nn(d1, d2, bool(point_major))
    return {'corr12': c12, 'corr21': c21, 'dist12': d12, 'idx1': i1, 'idx2': i2, 'count': cnt}


def find_correspondence_one_pair(feat1, feat2, device='cuda'):
    """Drop-in for the reference method: feat1 [n1,C], feat2 [n2,C] (numpy or torch, any device) ->
    (idx1, idx2) int64 numpy arrays of the mutual matches, exactly the reference's return value."""
    f1 = torch.as_tensor(feat1, dtype=torch.float32).to(device)[None].contiguous()
    f2 = torch.as_tensor(feat2, dtype=torch.float32).to(device)[None].contiguous()
    r = mutual_nn(f1, f2, point_major=True, want_dist=False)
    n = int(r['count'][0].item())
    return (r['idx1'][0, :n].cpu().numpy().astype(np.int64), r['idx2'][0, :n].cpu().numpy().astype(np.int64))


class MutualMatcher:
    """Pre-allocated, stream-ordered matcher for a fixed problem shape (what bench tools time): no per-call allocation,
    kernels enqueued straight through the C ABI."""

    def __init__(self, P, C, n1, n2, point_major=False, device='cuda', want_dist=True):
        from . import _lib
        self._L, self._check = _lib.lib, _lib.check
        self.P, self.C, self.n1, self.n2, self.point_major = int(P), int(C), int(n1), int(n2), bool(point_major)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('MutualMatcher runs on a CUDA device only (no CPU path exists)')
        i32, dev = torch.int32, self.device
        self.corr12 = torch.empty((P, n1), dtype=i32, device=dev)
        self.corr21 = torch.empty((P, n2), dtype=i32, device=dev)
        # want_dist=False: indices only, like the reference's find_correspondence_one_pair — skips the fp32 re-evaluation of
        # the matched distances (the only pass of the finish that touches the descriptors)
        self.dist12 = torch.empty((P, n1), dtype=torch.float32, device=dev) if want_dist else None
        self.idx1 = torch.empty((P, n1), dtype=i32, device=dev)
        self.idx2 = torch.empty((P, n1), dtype=i32, device=dev)
        self.count = torch.empty((P,), dtype=i32, device=dev)
        self._nws = self._L.ri_mutual_nn_workspace_bytes(P, C, n1, n2)
        self._ws = torch.empty(self._nws, dtype=torch.uint8, device=dev)

    def __call__(self, desc1, desc2):
        st = torch.cuda.current_stream().cuda_stream
        self._check(self._L.ri_mutual_nn_tf32x3(desc1.data_ptr(), desc2.data_ptr(), self.P, self.C, self.n1, self.n2,
                                                1 if self.point_major else 0, self.corr12.data_ptr(),
                                                self.corr21.data_ptr(),
                                                self.dist12.data_ptr() if self.dist12 is not None else None, self.idx1.data_ptr(),
                                                self.idx2.data_ptr(), self.count.data_ptr(), self._ws.data_ptr(),
                                                self._nws, st), 'ri_mutual_nn')
        return self

    @property
    def flops(self):
        """Useful flops of one call (the single fp32 contraction the reference's sgemm does)."""
        return 2.0 * self.P * self.n1 * self.n2 * self.C
