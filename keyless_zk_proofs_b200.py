"""Importable alias of the package directory ``keyless-zk-proofs_b200/`` (a hyphen is not a valid identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("keyless-zk-proofs_b200")
