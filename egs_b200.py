"""Import shim: the package directory is named `elastic-gpu-scheduler_b200` (not a Python
identifier), so load it under the module name `egs_b200`."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "elastic-gpu-scheduler_b200")
_spec = importlib.util.spec_from_file_location(
    "egs_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["egs_b200"] = _mod
_spec.loader.exec_module(_mod)
