"""Import shim: `import sfb_b200` loads the package in `sphericalfourierbesseldecompositions.jl_b200/`
(the directory is named after the reference and is not a valid Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sphericalfourierbesseldecompositions.jl_b200")
_spec = importlib.util.spec_from_file_location("sfb_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sfb_b200"] = _mod
_spec.loader.exec_module(_mod)
