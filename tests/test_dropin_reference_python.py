"""The drop-in claim, demonstrated: the reference's OWN, unmodified Python — PVCNN/modules/functional/*.py, PVCNN/modules/*.py
(PVConv, Voxelization, Spherical_Voxelization, BallQuery, SharedMLP, SE3d ...) and PVCNN/models/pvcnn_classify.py — runs a
forward pass of PVCNN_classifier twice on the same seeded weights and inputs:

  (A) with `PVCNN.modules.functional.backend._backend` = the reference's CUDA backend (oracle/_ref, unmodified sources
      recompiled for sm_100a), i.e. the reference as it is;
  (B) with `_backend` = ri_b200.backend._backend, i.e. libri_b200.so behind the same names.

Asserted: the voxel indices every PVConv block computed are identical, and the model outputs agree within 1e-5 relative (the
reference's own run-to-run noise — its voxelizer sums with float atomics — is measured and reported beside it).  Configurations:
exp13 = sph_dg (/root/reference/configs/modelnet40/pvcnn/experiments/SO3_SO3/exp13.py:10-20 on top of
configs/modelnet40/pvcnn/__init__.py:5-12) and deepgmr_mn40_cu_dg = cu_dg (.../deepgmr_mn40_cu_dg/__init__.py:14-22).

The reference package is imported from oracle/_ref/pyref (staged by oracle/build_ref.py::stage_python: python files only,
git-ignored), `open3d` — imported at the top of pvcnn_classify.py, used only by the 'fpfh' branch — is stubbed, and the module
that would JIT-build the CUDA sources (functional/backend.py:14-39) is replaced by a two-line stand-in whose `_backend` forwards
to whichever library is under test."""
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BLOCKS = ((64, 1, 32), (128, 1, 32), (256, 1, None), (512, 1, None))           # configs/modelnet40/pvcnn/__init__.py:7
CONFIGS = {
    "sph_dg": dict(voxel_shape="spherical", extra_feature_channels=0, is_classify=True),      # exp13.py
    "cu_dg": dict(voxel_shape="cube", extra_feature_channels=4, is_classify=False),           # deepgmr_mn40_cu_dg/__init__.py
}


class _Switch:
    """Stands where functional/backend.py's `_backend` stands; forwards every attribute to the library under test."""
    target = None

    def __getattr__(self, name):
        return getattr(_Switch.target, name)


@pytest.fixture(scope="module")
def reference_package(ref_backend):
    from oracle.build_ref import stage_python
    root = stage_python()
    if root is None:
        pytest.fail("oracle/_ref/pyref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    saved = {k: v for k, v in sys.modules.items() if k == "open3d" or k.startswith("PVCNN")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, root)
    sys.modules["open3d"] = types.ModuleType("open3d")
    stand_in = types.ModuleType("PVCNN.modules.functional.backend")
    stand_in._backend = _Switch()
    stand_in.__all__ = ["_backend"]
    sys.modules["PVCNN.modules.functional.backend"] = stand_in
    import PVCNN.models.pvcnn_classify as mod
    import PVCNN.modules as modules
    yield mod, modules
    sys.path.remove(root)
    for k in [k for k in sys.modules if k == "open3d" or k.startswith("PVCNN")]:
        del sys.modules[k]
    sys.modules.update(saved)


def _forward(model, modules, backend, x):
    _Switch.target = backend
    inds = []
    hooks = [m.register_forward_hook(lambda _m, _i, out: inds.append(out[1].clone()))
             for m in model.modules() if isinstance(m, (modules.Voxelization, modules.Spherical_Voxelization))]
    with torch.no_grad():
        y = model(x)
    for h in hooks:
        h.remove()
    torch.cuda.synchronize()
    return y, inds


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_reference_model_runs_unchanged_on_libri_b200(reference_package, ref_backend, name, capsys):
    import ri_b200
    from ri_b200 import synth
    mod, modules = reference_package
    cfg = CONFIGS[name]
    torch.manual_seed(1234)
    model = mod.PVCNN_classifier(blocks=BLOCKS, dim_k=512, point_kernel_formal="dgcnn_kernel", num_classes=40,
                                 with_coeff=True, with_se=True, rot_invariant_preprocess="change_coords",
                                 with_local_feat="ppf", with_transform_fine_tune=False, use_new_coords_for_voxel=False,
                                 width_multiplier=1, voxel_resolution_multiplier=1, **cfg).cuda().eval()
    with torch.no_grad():                                    # fresh BatchNorm statistics are (0, 1): give them some spread
        for m in model.modules():
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
                m.running_mean.normal_(0.0, 0.1); m.running_var.uniform_(0.5, 1.5)
    B, N = 8, 1024
    x = torch.from_numpy(synth.make_clouds(B, N, seed=99)).cuda()
    y_ref, inds_ref = _forward(model, modules, ref_backend, x)
    y_ref2, _ = _forward(model, modules, ref_backend, x)                     # the reference against itself: float atomics
    y_ours, inds_ours = _forward(model, modules, ri_b200.backend._backend, x)
    assert len(inds_ref) == len(inds_ours) == 2                              # two PVConv blocks voxelize
    for a, b in zip(inds_ref, inds_ours):
        assert a.dtype == b.dtype and torch.equal(a, b), "voxel indices differ in %d places" % int((a != b).sum())
    scale = float(y_ref.abs().max())
    noise = float((y_ref - y_ref2).abs().max()) / scale
    err = float((y_ref - y_ours).abs().max()) / scale
    with capsys.disabled():
        print("\n[%s] output %s, max |ours - reference| / max |reference| = %.2e (reference vs itself: %.2e)"
              % (name, tuple(y_ref.shape), err, noise))
    assert torch.isfinite(y_ours).all()
    assert err <= max(1e-5, 4 * noise)
