"""Import shim: the package directory is `hmc.jl_b200/` (a dot is not importable), so this module loads it
from its path and installs it as `hmc_jl_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hmc.jl_b200")
_spec = importlib.util.spec_from_file_location("hmc_jl_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["hmc_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
