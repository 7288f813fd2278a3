"""Import alias: the package directory name contains a hyphen
(macro-particle_simulation_for_magnetic_reconnection_b200), so it is loaded by
name here; `import mrg_b200 as mrg` gives the package."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("macro-particle_simulation_for_magnetic_reconnection_b200")
sys.modules[__name__] = _pkg
