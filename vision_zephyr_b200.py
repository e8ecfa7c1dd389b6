"""Import shim: the package directory is named `vision-zephyr_b200/` (not a Python identifier);
importing this module registers that directory as the package `vision_zephyr_b200`."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "vision-zephyr_b200")
_spec = _ilu.spec_from_file_location("vision_zephyr_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["vision_zephyr_b200"] = _mod
_spec.loader.exec_module(_mod)
