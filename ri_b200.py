"""Import alias: `import ri_b200` loads the package that lives in
`point-cloud-registration-based-on-rotation-invariant-feature_b200/` (a directory name Python cannot import
directly because of the hyphens) and registers it in sys.modules under this name."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "point-cloud-registration-based-on-rotation-invariant-feature_b200")
_spec = importlib.util.spec_from_file_location("ri_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ri_b200"] = _mod
_spec.loader.exec_module(_mod)
