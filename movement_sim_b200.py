"""Import shim: the product package lives in ``movement-sim_b200/`` (the reference's name plus
``_b200``), which is not a Python identifier.  ``import movement_sim_b200`` lands here and loads that
directory as the package of the same name."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "movement-sim_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
