"""Golden vectors for the 'change_coords' preprocessing.

The reference code for this path is pure torch inside PVCNN_classifier.forward (PVCNN/models/pvcnn_classify.py:153-184);
the class cannot be instantiated here (its other branches need open3d and the CUDA backend), so the fixture is produced by
running THOSE LINES' torch operations, in their order, on the CPU: same calls (`mean`, `norm(dim=1)`, `argsort(descending)`,
`.norm()`, `(a*b).sum()`, `bmm`, `cross`), same loops, same thresholds.  Pin strength: a transcription, not the imported
class — stated as such in DESIGN.md.

    python oracle/make_golden_lrf.py        # writes tests/golden/lrf.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def change_coords_torch(coords):
    """The torch calls of pvcnn_classify.py:153-184 in their order (mean, norm(dim=1), argsort(descending), per-vector norm(),
    (a*b).sum(), bmm, cross, norm(dim=1)) — written out here, not imported: see the module docstring."""
    nb, _, npts = coords.shape
    centred = coords - coords.mean(dim=2, keepdim=True)                                   # :154
    by_radius = torch.argsort(centred.norm(dim=1), dim=1, descending=True)                # :155
    ex = torch.zeros(nb, 3, 1).to(centred)
    ey = torch.zeros(nb, 3, 1).to(centred)
    for c in range(nb):
        first = centred[c, :, by_radius[c, 0]]                                            # :159
        assert first.norm() > 1e-5                                                        # :160
        first = first / first.norm()
        second, cosang = None, None
        for q in range(1, npts):                                                          # :162-169
            cand = centred[c, :, by_radius[c, q]]
            if cand.norm() < 1e-5:
                continue
            cand = cand / cand.norm()
            cosang = (first * cand).sum()
            if cosang < 0.9 and cosang > -0.9:
                second = cand
                break
        assert second is not None                                                         # :170
        ex[c, :, :] = first.unsqueeze(1)
        ey[c, :, :] = second.unsqueeze(1)
    ex -= ey * (ex.permute(0, 2, 1).bmm(ey))                                              # :175 Gram-Schmidt, y kept
    assert (ex.norm(dim=1, keepdim=True) < 1e-5).sum() < 1                                # :176
    ex /= ex.norm(dim=1, keepdim=True)                                                    # :177
    ez = ex.cross(ey, dim=1)                                                              # :179
    ez = ez / ez.norm(dim=1, keepdim=True)                                                # :180
    rows = [axis.permute(0, 2, 1).bmm(centred) for axis in (ex, ey, ez)]                  # :181-183
    return torch.cat(rows, dim=1), torch.cat((ex, ey, ez), dim=2).permute(0, 2, 1)


def main():
    spec = __import__("importlib.util").util.spec_from_file_location("ri_synth", os.path.join(
        ROOT, "point-cloud-registration-based-on-rotation-invariant-feature_b200", "synth.py"))
    synth = __import__("importlib.util").util.module_from_spec(spec); spec.loader.exec_module(synth)
    out = {}
    clouds = synth.make_clouds(16, 1024, seed=77)[:, :3]
    # adversarial rows: the 2nd..5th farthest points (anti)parallel to the farthest one, so base_y must be searched for
    adv = synth.make_clouds(4, 500, seed=78)[:, :3].copy()
    for q in range(4):
        c = adv[q] - adv[q].mean(1, keepdims=True)
        far = np.argmax(np.linalg.norm(c, axis=0))
        for s, sc in enumerate((0.99, -0.98, 0.97, -0.96)):
            adv[q][:, (far + 1 + s) % 500] = adv[q].mean(1) + c[:, far] * sc
    for name, x in (("surface", clouds), ("parallel", adv)):
        t = torch.from_numpy(np.ascontiguousarray(x))
        new, bases = change_coords_torch(t)
        out[name + "_coords"] = x
        out[name + "_mean"] = t.mean(dim=2).numpy()
        out[name + "_new"] = new.numpy()
        out[name + "_bases"] = bases.contiguous().numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lrf.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
