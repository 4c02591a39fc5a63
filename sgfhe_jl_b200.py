"""Import shim: the package directory is named `sgfhe.jl_b200/` (a dot is not importable as-is).

    import sgfhe_jl_b200 as sg
    params = sg.Params(64); ...
"""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "sgfhe.jl_b200")
_spec = _u.spec_from_file_location("sgfhe_jl_b200", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["sgfhe_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
