"""B200-native collaborative-perception hot path (hand-written sm_100a CUDA behind a C ABI).

Import with ``importlib.import_module("airv2x-perception_b200")`` (the directory name is not a Python identifier)
or through the repo-root helper ``a2x_import.pkg()``.
"""
__version__ = "0.1.0"
