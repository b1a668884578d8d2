"""Import shim: `import cgvc` loads the package in `voice-converter-cyclegan_b200/` (whose name is not an identifier)."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "voice-converter-cyclegan_b200")
_spec = _u.spec_from_file_location("cgvc", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["cgvc"] = _mod
_spec.loader.exec_module(_mod)
