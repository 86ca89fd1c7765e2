"""TEST INFRASTRUCTURE ONLY (container-side): import the real reference from /root/reference on CPU.

Used by scripts/make_golden.py to validate the oracle restatement and to generate tests/golden/*.npz.
/root/reference does not exist on the GPU box, so nothing under tests/ (-m gpu), bench.py or smoke() imports this.
Recipe: SURVEY.md Appendix A/E — absent cosmetic packages (matplotlib, shapely, open3d, ...) are replaced by
MagicMock modules through a meta-path finder appended LAST, so real packages always win.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

REF_ROOT = "/root/reference"
COSMETIC = ("matplotlib", "efficientnet_pytorch", "shapely", "icecream", "open3d", "pyquaternion", "tkinter",
            "tensorboardX", "pypcd", "skimage", "timm", "more_itertools", "mpl_toolkits", "easydict", "cumm",
            "spconv", "pyparsing")


class _Loader(importlib.abc.Loader):
    def create_module(self, spec):
        m = MagicMock()
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        return m

    def exec_module(self, module):
        pass


class _Finder(importlib.abc.MetaPathFinder):
    def __init__(self, names):
        self.names = names

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.names:
            return importlib.machinery.ModuleSpec(name, _Loader(), is_package=True)
        return None


_installed = False


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "opencood"))


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    missing = []
    for n in COSMETIC:
        try:
            importlib.import_module(n)
        except Exception:
            missing.append(n)
    sys.meta_path.append(_Finder(tuple(missing)))
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def load_hypes(rel_yaml):
    install()
    from opencood.hypes_yaml import yaml_utils
    return yaml_utils.load_yaml(os.path.join(REF_ROOT, "opencood", "hypes_yaml", rel_yaml))


def create_model(hypes):
    install()
    from opencood.tools import train_utils
    return train_utils.create_model(hypes)
