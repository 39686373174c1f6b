"""Imports the product package (its directory name is not a Python identifier)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "wavelet-noise-in-ray-tracing_b200"


def load():
    return importlib.import_module(PKG)


def load_sub(name):
    return importlib.import_module(f"{PKG}.{name}")
