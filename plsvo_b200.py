"""Import shim: makes the hyphenated package directory `pl-svo_b200/` importable as `plsvo_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pl-svo_b200")
_spec = importlib.util.spec_from_file_location(
    "plsvo_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["plsvo_b200"] = _mod
_spec.loader.exec_module(_mod)
