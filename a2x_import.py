"""Repo-root helper: ``pkg()`` returns the ``airv2x-perception_b200`` package (its name has a hyphen)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_NAME = "airv2x-perception_b200"


def pkg(sub=None):
    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)
    return importlib.import_module(PKG_NAME + ("." + sub if sub else ""))
