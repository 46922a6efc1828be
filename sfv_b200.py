"""Import alias for the package directory ``symbols-from-video_b200/`` (a hyphen
is not a legal Python identifier).  ``import sfv_b200`` == that package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "symbols-from-video_b200")
_spec = importlib.util.spec_from_file_location("sfv_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sfv_b200"] = _mod
_spec.loader.exec_module(_mod)
