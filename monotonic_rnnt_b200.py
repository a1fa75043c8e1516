"""Import shim: the package directory is named ``monotonic-rnnt_b200`` (not a valid Python identifier),
so ``import monotonic_rnnt_b200`` lands here and this module replaces itself with the real package."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "monotonic-rnnt_b200")
_spec = importlib.util.spec_from_file_location(
    "monotonic_rnnt_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["monotonic_rnnt_b200"] = _mod
_spec.loader.exec_module(_mod)
