"""Import alias: the product package lives in the directory ``lp-gnn_b200/`` (the name the
project layout prescribes), which is not a valid Python identifier.  Importing ``lpgnn_b200``
loads that directory as the package ``lpgnn_b200``; nothing else lives here."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "lp-gnn_b200")
_spec = _ilu.spec_from_file_location("lpgnn_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["lpgnn_b200"] = _mod
_spec.loader.exec_module(_mod)
