"""Import shim: the package directory name required by the build contract contains '-', so `import sr3_b200`
resolves to it through this loader."""
import importlib.util
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "image-super-resolution-via-iterative-refinement_b200")
_spec = importlib.util.spec_from_file_location("sr3_b200", os.path.join(_PKG, "__init__.py"), submodule_search_locations=[_PKG])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sr3_b200"] = _mod
_spec.loader.exec_module(_mod)
