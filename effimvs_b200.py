"""Import alias for the package directory ``effi-mvs-plus_b200/`` (a hyphen is not a
valid identifier, so the directory is registered under the module name ``effimvs_b200``).

    import effimvs_b200
    from effimvs_b200 import dropin, net, fusion
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "effi-mvs-plus_b200")
_spec = _ilu.spec_from_file_location("effimvs_b200", _os.path.join(_PKG_DIR, "__init__.py"),
                                     submodule_search_locations=[_PKG_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["effimvs_b200"] = _mod
_spec.loader.exec_module(_mod)
