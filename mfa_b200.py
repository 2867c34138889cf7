"""Import shim: the package directory is named `metal-flash-attention_b200/` (not a valid Python
identifier), so `import mfa_b200` loads it from there and registers it under this name."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metal-flash-attention_b200")
_spec = importlib.util.spec_from_file_location(
    "mfa_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_module = importlib.util.module_from_spec(_spec)
sys.modules["mfa_b200"] = _module
_spec.loader.exec_module(_module)
