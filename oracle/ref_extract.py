"""Run pieces of the reference's own Python source (test infrastructure only; needs /root/reference, i.e. this container).

The reference's registration meter and classifier live in modules whose imports (h5py, open3d, the JIT-built CUDA backend)
are absent here, but the two pieces of this path are plain numpy / torch: they are cut out of the reference FILES by their
syntax tree — never retyped, never copied into this repository — compiled, and executed:

  matcher_function()       MeterModelNet40_registration.find_correspondence_one_pair   (datasets/deepgmr_mn40.py:232-244)
  change_coords_function() the `elif self.rot_invariant_preprocess=='change_coords':` branch of PVCNN_classifier.forward up
                           to the assignment of `features`                              (PVCNN/models/pvcnn_classify.py:153-184)
"""
import ast
import os

REF = "/root/reference"


def available():
    return os.path.isdir(REF)


def _parse(rel):
    path = os.path.join(REF, rel)
    src = open(path).read()
    return path, src, ast.parse(src, filename=path)


def matcher_function():
    """-> f(feat1 [n1,c], feat2 [n2,c]) -> (idx1, idx2): the reference method itself, bound to a dummy self."""
    import numpy as np
    path, src, tree = _parse("datasets/deepgmr_mn40.py")
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "find_correspondence_one_pair")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np}
    exec(compile(mod, path, "exec"), ns)
    f = ns["find_correspondence_one_pair"]
    return (lambda feat1, feat2: f(None, feat1, feat2)), (fn.lineno, fn.end_lineno)


def change_coords_function():
    """-> f(coords [b,3,n] torch) -> features [b,3,n] (and the three base vectors): the reference branch body itself."""
    import torch
    path, src, tree = _parse("PVCNN/models/pvcnn_classify.py")
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == "PVCNN_classifier")
    fwd = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward")
    branch = None
    for node in ast.walk(fwd):
        if isinstance(node, ast.If) and isinstance(node.test, ast.Compare) and node.test.comparators \
                and isinstance(node.test.comparators[0], ast.Constant) and node.test.comparators[0].value == "change_coords":
            branch = node
            break
    assert branch is not None, "change_coords branch not found"
    body = []
    for st in branch.body:
        body.append(st)
        if isinstance(st, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "features" for t in st.targets):
            break
    wrapper = ast.parse("def change_coords(coords, b, n):\n    pass\n    return features, batch_base_x, batch_base_y, batch_base_z")
    fdef = wrapper.body[0]
    fdef.body = body + [fdef.body[-1]]
    ast.fix_missing_locations(wrapper)
    ns = {"torch": torch}
    exec(compile(wrapper, path, "exec"), ns)
    f = ns["change_coords"]
    return (lambda coords: f(coords, coords.shape[0], coords.shape[2])), (body[0].lineno, body[-1].end_lineno)
