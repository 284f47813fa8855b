"""Import shim: the package directory is `needleman-wunsch_b200/` (a hyphen is
not a valid identifier), so `import nw_b200` gives the same module."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_m = importlib.import_module("needleman-wunsch_b200")
sys.modules[__name__] = _m
