"""Golden vectors for the 'change_coords' preprocessing.

The reference code for this path is pure torch inside PVCNN_classifier.forward (PVCNN/models/pvcnn_classify.py:153-184);
the class cannot be instantiated here (its other branches need open3d and the CUDA backend), so the fixture is produced by
EXECUTING those statements: oracle/ref_extract.py cuts the `change_coords` branch out of the reference file by its syntax
tree, wraps it in a function of (coords, b, n) and runs it on the CPU.  Pin strength: reference-executed.

    python oracle/make_golden_lrf.py        # writes tests/golden/lrf.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def change_coords_torch(coords):
    """The reference branch itself: its statements are cut out of pvcnn_classify.py by their syntax tree
    (oracle/ref_extract.py) and executed on `coords` — nothing of it is retyped here."""
    from oracle import ref_extract
    f, lines = ref_extract.change_coords_function()
    features, ex, ey, ez = f(coords)
    change_coords_torch.lines = lines
    return features, torch.cat((ex, ey, ez), dim=2).permute(0, 2, 1)


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
    out["reference_lines"] = np.array(change_coords_torch.lines)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lrf.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
