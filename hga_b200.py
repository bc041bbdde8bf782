"""Importable alias of the package directory `hybrid-genome-assembler_b200/` (a hyphen is not a valid Python
identifier, so `import hga_b200` forwards to it)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("hybrid-genome-assembler_b200")
sys.modules[__name__] = _pkg
