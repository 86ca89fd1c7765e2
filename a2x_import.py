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


MODEL_NAMES = ("airv2x_where2com", "airv2x_cobevt", "airv2x_v2xvit", "point_pillar_where2comm", "point_pillar_cobevt",
               "point_pillar_v2xvit")
LOSS_NAMES = ("point_pillar_loss_multiclass", "point_pillar_loss")


def install(names=MODEL_NAMES, losses=LOSS_NAMES):
    """Register the B200 drop-in modules under the reference's registry paths, so that the UNMODIFIED
    `opencood.tools.train_utils.create_model(hypes)` (train_utils.py:288-325: `importlib.import_module("opencood.models." +
    core_method)` + class-name match) returns them. Also publishes the package under the importable alias
    `airv2x_perception_b200` (the directory name has a hyphen). `losses`: likewise `opencood.loss.<name>` for
    `train_utils.create_loss(hypes)` (train_utils.py:328-368) -> the criterion classes on the fused loss kernel.
    Returns the previous sys.modules entries for uninstall()."""
    prev = {}
    for n in names:
        key = "opencood.models." + n
        prev[key] = sys.modules.get(key)
        sys.modules[key] = pkg("opencood.models." + n)
    for n in losses or ():
        key = "opencood.loss." + n
        prev[key] = sys.modules.get(key)
        sys.modules[key] = pkg("opencood.loss." + n)
    sys.modules.setdefault("airv2x_perception_b200", pkg())
    return prev


def uninstall(prev):
    for key, mod in prev.items():
        if mod is None:
            sys.modules.pop(key, None)
        else:
            sys.modules[key] = mod


def install_dataset(name="IntermediateFusionDatasetAirv2x"):
    """Make the UNMODIFIED `opencood.data_utils.datasets.build_dataset(hypes, visualize, train)` (datasets/__init__.py:94-106:
    a lookup of `fusion.core_method` in the module's `__all__` dict) return this repo's dataset class. Needs the reference
    package importable; returns the previous entry (pass it to `uninstall_dataset`)."""
    import opencood.data_utils.datasets as D
    prev = D.__all__.get(name)
    D.__all__[name] = getattr(pkg("intermediate_fusion_dataset"), name)
    return prev


def uninstall_dataset(prev, name="IntermediateFusionDatasetAirv2x"):
    import opencood.data_utils.datasets as D
    if prev is None:
        D.__all__.pop(name, None)
    else:
        D.__all__[name] = prev
