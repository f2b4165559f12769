"""Import alias: `import hpv_b200` loads the package in ./hp-vpinns_b200/ (whose name is not an identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hp-vpinns_b200")
_spec = importlib.util.spec_from_file_location("hpv_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["hpv_b200"] = _mod
_spec.loader.exec_module(_mod)
