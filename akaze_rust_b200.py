"""Import shim: the package directory is named `akaze-rust_b200` (not a valid Python identifier), so this
module loads it under the importable name `akaze_rust_b200`. `import akaze_rust_b200` from the repo root."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "akaze-rust_b200")
_spec = importlib.util.spec_from_file_location(
    "akaze_rust_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["akaze_rust_b200"] = _mod
_spec.loader.exec_module(_mod)
