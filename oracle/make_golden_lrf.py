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
    b, _, n = coords.shape
    norm_coords = coords - coords.mean(dim=2, keepdim=True)
    rank = torch.argsort(norm_coords.norm(dim=1), dim=1, descending=True)
    batch_base_x = torch.zeros(b, 3, 1).to(norm_coords)
    batch_base_y = torch.zeros(b, 3, 1).to(norm_coords)
    for i in range(b):
        base_x = norm_coords[i, :, rank[i, 0]]
        assert (base_x.norm() > 1e-5)
        base_x = base_x / base_x.norm()
        for j in range(1, n):
            base_y = norm_coords[i, :, rank[i, j]]
            if base_y.norm() < 1e-5:
                continue
            base_y = base_y / base_y.norm()
            lamda = (base_x * base_y).sum()
            if (lamda < 0.9 and lamda > -0.9):
                break
        assert (lamda < 0.9 and lamda > -0.9)
        batch_base_x[i, :, :] = base_x.unsqueeze(1)
        batch_base_y[i, :, :] = base_y.unsqueeze(1)
    batch_base_x -= batch_base_y * (batch_base_x.permute(0, 2, 1).bmm(batch_base_y))
    assert (batch_base_x.norm(dim=1, keepdim=True) < 1e-5).sum() < 1
    batch_base_x /= batch_base_x.norm(dim=1, keepdim=True)
    batch_base_z = batch_base_x.cross(batch_base_y, dim=1)
    batch_base_z = batch_base_z / batch_base_z.norm(dim=1, keepdim=True)
    new_x = batch_base_x.permute(0, 2, 1).bmm(norm_coords)
    new_y = batch_base_y.permute(0, 2, 1).bmm(norm_coords)
    new_z = batch_base_z.permute(0, 2, 1).bmm(norm_coords)
    return torch.cat((new_x, new_y, new_z), dim=1), torch.cat((batch_base_x, batch_base_y, batch_base_z), dim=2).permute(0, 2, 1)


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
