"""Golden vectors for the mutual-NN matcher made by EXECUTING the reference's own method.

`find_correspondence_one_pair` (/root/reference/datasets/deepgmr_mn40.py:232-244) is cut out of the reference file by its
syntax tree (oracle/ref_extract.py) and run with numpy on seeded descriptors; the cases are regenerated from their seeds by
the tests (tests/test_matcher_golden.py::make_case — keep the two in step), so the fixture stores only the reference's
outputs: idx1, idx2 and, per row, the value and position of the reference's fp32 `diff` minimum and the runner-up gap.

    python oracle/make_golden_matcher.py        # writes tests/golden/matcher.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_extract  # noqa: E402

CASES = {
    # name: (kind, n1, n2, C, seed)
    "random_512": ("random", 1024, 1024, 512, 11),
    "random_128": ("random", 700, 650, 128, 12),
    "random_small": ("random", 37, 53, 16, 13),
    "registration_512": ("registration", 1024, 1024, 512, 14),
    "registration_64": ("registration", 512, 512, 64, 15),
    "ties_duplicates": ("ties", 256, 256, 64, 16),
    "one_row": ("random", 1, 300, 32, 17),
}


def make_case(kind, n1, n2, C, seed):
    """Seeded descriptors [n1,C], [n2,C] fp32 (numpy Generator streams are stable across versions)."""
    g = np.random.default_rng(seed)
    if kind == "random":
        return g.standard_normal((n1, C)).astype(np.float32), g.standard_normal((n2, C)).astype(np.float32)
    if kind == "registration":                     # target = permuted source + noise, a third of the points replaced
        f1 = g.standard_normal((n1, C)).astype(np.float32)
        perm = g.permutation(n1)[:n2]
        f2 = (f1[perm] + 0.1 * g.standard_normal((n2, C))).astype(np.float32)
        out = g.random(n2) < 0.33
        f2[out] = g.standard_normal((int(out.sum()), C)).astype(np.float32)
        return f1, f2
    if kind == "ties":                             # small integers: every product and sum is exact, duplicates tie exactly
        base = g.integers(-4, 5, size=(n1 // 2, C)).astype(np.float32)
        return np.concatenate([base, base], 0)[g.permutation(n1)], np.concatenate([base, base], 0)[:n2]
    raise ValueError(kind)


def main():
    f, lines = ref_extract.matcher_function()
    out = {"reference_lines": np.array(lines)}
    for name, spec in CASES.items():
        f1, f2 = make_case(*spec)
        idx1, idx2 = f(f1, f2)
        out[name + "_idx1"] = idx1.astype(np.int64)
        out[name + "_idx2"] = idx2.astype(np.int64)
        out[name + "_spec"] = np.array([spec[1], spec[2], spec[3], spec[4]])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "matcher.npz"), **out)
    print("reference lines", lines, {k: v.shape for k, v in out.items() if k.endswith("idx1")})


if __name__ == "__main__":
    main()
