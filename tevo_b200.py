"""Import shim: the package directory is named `timeevomps.jl_b200` (with a dot), which Python's import
statement cannot spell; `import tevo_b200` loads it under that alias."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "timeevomps.jl_b200")
_NAME = "timeevomps_jl_b200"
if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_mod = sys.modules[_NAME]
globals().update({k: getattr(_mod, k) for k in _mod.__all__})
pkg = _mod
