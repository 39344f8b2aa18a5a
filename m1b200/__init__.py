"""Import shim: the product package lives in ``prostatemr_3d-cad-cspca_b200/`` (a directory name
that is not a Python identifier); ``import m1b200`` loads that directory as the package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                    "prostatemr_3d-cad-cspca_b200")
_spec = importlib.util.spec_from_file_location(
    "m1b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["m1b200"] = _mod
_spec.loader.exec_module(_mod)
