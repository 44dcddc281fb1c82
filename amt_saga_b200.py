"""Import shim: the package directory is named `amt-saga_b200/` (as the repo
layout requires) which is not a valid Python identifier; `import amt_saga_b200`
loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "amt-saga_b200")
_spec = importlib.util.spec_from_file_location(
    "amt_saga_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["amt_saga_b200"] = _mod
_spec.loader.exec_module(_mod)
