"""Importable alias of the ``audio-deepfake-detection-fmsl_b200`` package (whose directory name,
mirroring the reference repository, is not a Python identifier):  ``import b200_frontend as fe``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("audio-deepfake-detection-fmsl_b200")
sys.modules[__name__] = _pkg
