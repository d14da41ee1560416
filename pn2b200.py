"""Importable alias of the product package (its directory name contains a hyphen)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("khairil_tum-facade_semantic_segmentation_b200")
