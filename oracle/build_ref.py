"""Build the UNMODIFIED reference CUDA backend into oracle/_ref/ (test infrastructure only).

The reference (`/root/reference/PVCNN/modules/functional/backend.py:14-39`) JIT-compiles 21
sources into one pybind11 module `_multi_shape_pvcnn_backend`.  This recipe compiles the very
same files *where they lie* under /root/reference for sm_100a and drops the resulting .so into
`oracle/_ref/` (git-ignored, but shipped to the GPU box by gpurun).  No reference source is
copied into this repository.

The .so is the executable specification used by `tests/` (GPU parity), by
`oracle/make_golden.py` (golden fixture generation) and by `bench.py --impl reference-cuda`.
It is never imported by the product package.

Usage:  python oracle/build_ref.py            (no-op if /root/reference is absent or .so is fresh)
"""
import os
import sys
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/PVCNN/modules/functional/src"
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "_multi_shape_pvcnn_backend"

# the exact list of backend.py:17-37
SOURCES = [
    'ball_query/ball_query.cpp', 'ball_query/ball_query.cu',
    'grouping/grouping.cpp', 'grouping/grouping.cu',
    'interpolate/neighbor_interpolate.cpp', 'interpolate/neighbor_interpolate.cu',
    'interpolate/trilinear_devox.cpp', 'interpolate/trilinear_devox.cu',
    'sampling/sampling.cpp', 'sampling/sampling.cu',
    'voxelization/vox.cpp', 'voxelization/vox.cu',
    'interpolate/spherical_trilinear_devox.cpp', 'interpolate/spherical_trilinear_devox.cu',
    'spherical_voxelization/spherical_vox.cpp', 'spherical_voxelization/spherical_vox.cu',
    'spherical_ppf/ppf.cpp', 'spherical_ppf/ppf.cu',
    'knn/knn.cpp', 'knn/knn.cu',
    'bindings.cpp',
]


def so_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def build(force=False, verbose=False):
    if not os.path.isdir(REF_SRC):
        return None  # GPU box: only the prebuilt file is used
    if os.path.exists(so_path()) and not force:
        return so_path()
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 8))
    from torch.utils.cpp_extension import load
    build_dir = os.path.join(OUT_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    load(name=NAME,
         extra_cflags=['-O3', '-std=c++17'],          # same as backend.py:15
         sources=[os.path.join(REF_SRC, f) for f in SOURCES],
         build_directory=build_dir, verbose=verbose, is_python_module=False)
    shutil.copy2(os.path.join(build_dir, NAME + ".so"), so_path())
    shutil.rmtree(build_dir, ignore_errors=True)
    return so_path()


def load_ref():
    """Import the prebuilt reference backend (needs `import torch` first). Returns module or None."""
    p = so_path()
    if not os.path.exists(p):
        return None
    import torch  # noqa: F401  (libtorch symbols)
    import importlib.util
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("reference backend:", p)
