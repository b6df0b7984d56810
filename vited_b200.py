"""Import alias for the package directory ``vit-ed_b200/``.

Python identifiers cannot contain '-', so ``import vited_b200`` loads ``vit-ed_b200/__init__.py`` under this name
(sub-modules resolve through the package's search path, e.g. ``vited_b200.grid``).
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vit-ed_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
