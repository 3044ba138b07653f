"""Import shim: the package directory is named `reachabilityanalysis.jl_b200` (it contains a
dot, so a plain `import` cannot name it).  `import reach_b200` loads that directory as the
module `reach_b200` -- tests, bench.py and __graft_entry__.py all go through this."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "reachabilityanalysis.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "reach_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["reach_b200"] = _mod
_spec.loader.exec_module(_mod)
