"""CPU (-m "not gpu"), container side: the REAL `opencood.tools.train_utils.create_model(hypes)` of the reference
dispatches to the B200 drop-in modules once `a2x_import.install()` has registered them, for all six registry names, with
the shipped yamls unmodified, and a state_dict produced by the REAL reference model loads with strict=True (identical key
names and shapes). Skipped where /root/reference does not exist (the GPU box)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")

CASES = [
    ("airv2x_where2com", "airv2x/lidar/det/airv2x_intermediate_where2com.yaml", "Airv2xWhere2com"),
    ("airv2x_cobevt", "airv2x/lidar/det/airv2x_intermediate_cobevt.yaml", "Airv2xCoBEVT"),
    ("airv2x_v2xvit", "airv2x/lidar/det/airv2x_intermediate_v2xvit.yaml", "Airv2xV2XVit"),
    ("point_pillar_where2comm", "V2X-R/LiDAR/V2XR_where2comm.yaml", "PointPillarWhere2comm"),
    ("point_pillar_cobevt", "V2X-R/LiDAR/V2XR_cobevt.yaml", "PointPillarCoBEVT"),
    ("point_pillar_v2xvit", "V2X-R/LiDAR/V2XR_v2xvit.yaml", "PointPillarV2XVit"),
]


def _load(yaml_rel, tmp_path):
    """the shipped yaml; V2XR_cobevt / V2XR_v2xvit ship a voxel height that gives nz = 2, on which the reference's own
    PointPillarScatter asserts (point_pillar_scatter.py:13), so — like scripts/make_golden_legacy_fusion.py — their
    voxel_size is set to [0.4, 0.4, 4] through the yaml's own key"""
    import re

    src = open(os.path.join(ref_import.REF_ROOT, "opencood", "hypes_yaml", yaml_rel)).read()
    if yaml_rel.endswith(("V2XR_cobevt.yaml", "V2XR_v2xvit.yaml")):
        src = re.sub(r"(voxel_size: &voxel_size )\[[^\]]+\]", r"\1[0.4, 0.4, 4]", src)
    p = os.path.join(str(tmp_path), "hypes.yaml")
    open(p, "w").write(src)
    ref_import.install()
    from opencood.hypes_yaml import yaml_utils
    return yaml_utils.load_yaml(p)


@pytest.mark.parametrize("name,yaml_rel,cls", CASES)
def test_real_create_model_dispatches_to_the_b200_module(name, yaml_rel, cls, tmp_path, monkeypatch):
    import a2x_import

    monkeypatch.chdir(tmp_path)
    os.makedirs("debug", exist_ok=True)                       # the reference writes debug images relative to the cwd
    hypes = _load(yaml_rel, tmp_path)
    assert hypes["model"]["core_method"] == name
    from opencood.tools import train_utils

    sys.modules.pop("opencood.models." + name, None)
    real = train_utils.create_model(hypes)                    # the reference's own torch module
    assert type(real).__module__ == "opencood.models." + name and "b200" not in (type(real).__module__)
    real_sd = {k: v.clone() for k, v in real.state_dict().items()}
    prev = a2x_import.install()
    try:
        ours = train_utils.create_model(hypes)                # same call, unmodified reference code
        assert type(ours).__name__ == cls
        assert type(ours).__module__.startswith("airv2x-perception_b200.")
        missing, unexpected = ours.load_state_dict(real_sd, strict=True)
        assert not missing and not unexpected
        assert sum(p.numel() for p in ours.parameters()) == sum(p.numel() for p in real.parameters())
        for k, v in ours.state_dict().items():
            assert torch.equal(v, real_sd[k]), k
        with pytest.raises(RuntimeError):                     # no CPU path: the module refuses to run off a CUDA device
            ours.eval()
            ours({})
        import importlib
        assert importlib.import_module("airv2x_perception_b200") is a2x_import.pkg()
    finally:
        a2x_import.uninstall(prev)
        sys.modules.pop("opencood.models." + name, None)


def test_bevencode_state_dict_matches_the_reference_module():
    """`lss.BevEncode(inC, outC)` (the camera branch's BEV encoder on the tap-GEMM kernels) has the key names, order and
    shapes of the REAL `opencood.models.sub_modules.lss_submodule.BevEncode`: a state_dict of the reference module loads
    with strict=True, and the seeded initialisation of the golden fixture (order dependent) is reproduced."""
    import a2x_import

    ref_import.install()
    from opencood.models.sub_modules.lss_submodule import BevEncode as RefBevEncode

    ref = RefBevEncode(64, 64)
    mine = a2x_import.pkg("lss").BevEncode(64, 64)
    rsd, msd = ref.state_dict(), mine.state_dict()
    assert list(rsd.keys()) == list(msd.keys())
    assert all(tuple(rsd[k].shape) == tuple(msd[k].shape) for k in rsd)
    mine.load_state_dict(rsd, strict=True)
    assert sum(p.numel() for p in ref.parameters()) == sum(p.numel() for p in mine.parameters())
    with pytest.raises(RuntimeError, match="CUDA"):
        mine.eval()(torch.zeros(1, 64, 16, 16))


@pytest.mark.parametrize("yaml_rel,cls", [("airv2x/lidar/det/airv2x_intermediate_where2com.yaml", "PointPillarLossMultiClass"),
                                          ("V2X-R/LiDAR/V2XR_where2comm.yaml", "PointPillarLoss")])
def test_real_create_loss_dispatches_to_the_b200_criterion(yaml_rel, cls, tmp_path):
    """the UNMODIFIED `train_utils.create_loss(hypes)` (train_utils.py:328-368) returns the criterion on the fused loss
    kernel once `a2x_import.install()` has registered `opencood.loss.<core_method>`; constructor keys as shipped"""
    import a2x_import

    hypes = _load(yaml_rel, tmp_path)
    from opencood.tools import train_utils

    name = (hypes["loss"][hypes["task"]] if hypes.get("task") else hypes["loss"])["core_method"]
    sys.modules.pop("opencood.loss." + name, None)
    real = train_utils.create_loss(hypes)
    assert type(real).__name__ == cls and type(real).__module__ == "opencood.loss." + name
    prev = a2x_import.install()
    try:
        ours = train_utils.create_loss(hypes)
        assert type(ours).__name__ == cls and type(ours).__module__.startswith("airv2x-perception_b200.")
        assert ours.cls_weight == real.cls_weight and ours.reg_coe == real.reg_coe
        assert hasattr(ours, "logging") and ours.loss_dict == {}
        if cls == "PointPillarLossMultiClass":
            assert ours.cls_num == real.cls_num
        z = torch.zeros(1, 2, 4, 4)
        with pytest.raises(RuntimeError, match="CUDA"):      # no CPU path
            ours({"psm": z, "rm": torch.zeros(1, 14, 4, 4), "obj": z},
                 {"targets": torch.zeros(1, 4, 4, 14), "pos_equal_one": torch.zeros(1, 4, 4, 2),
                  "class_ids": torch.zeros(1, 4, 4, 2, dtype=torch.int64)})
    finally:
        a2x_import.uninstall(prev)
        sys.modules.pop("opencood.loss." + name, None)
