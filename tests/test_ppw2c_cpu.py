"""CPU (-m "not gpu"): BASELINE config 1 (`point_pillar_where2comm`, 2 agents, 8k points, 128 x 128 BEV): the oracle
against the golden vectors recorded from the REAL reference, and the drop-in module's registry surface."""
import json
import os

import numpy as np
import torch

from oracle import w2c_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load():
    cfg = json.load(open(os.path.join(GOLD, "ppw2c_small_config.json")))
    gold = np.load(os.path.join(GOLD, "ppw2c_small.npz"))
    return cfg, gold


def golden_state_dict(model, gold):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd.update(O.det_init_state_dict(shapes, seed=int(gold["param_seed"])))
    sd["cls_head.bias"] = sd["cls_head.bias"] + float(gold["cls_bias_shift"])
    return sd


def golden_scene(cfg, gold):
    return O.make_scene_legacy(cfg["preprocess"], int(gold["n_agents"]), int(gold["n_points"]), int(gold["scene_seed"]),
                               cfg["preprocess"]["args"]["max_voxel_test"])


def test_oracle_and_registry_surface():
    import a2x_import

    M = a2x_import.pkg("opencood.models.point_pillar_where2comm")
    cls = [v for k, v in vars(M).items() if k.lower() == "point_pillar_where2comm".replace("_", "")]
    assert len(cls) == 1 and cls[0] is M.PointPillarWhere2comm
    cfg, gold = load()
    model = M.PointPillarWhere2comm(cfg["model_args"])
    assert sum(p.numel() for p in model.parameters()) == 8057386
    sd = golden_state_dict(model, gold)
    assert sd["shrink_conv.layers.0.double_conv.0.weight"].shape == (256, 384, 3, 3)
    assert sd["pillar_vfe.pfn_layers.0.linear.weight"].shape == (64, 10)
    torch.set_num_threads(8)
    with torch.no_grad():
        out, _ = O.pp_where2comm_forward(sd, cfg["model_args"], golden_scene(cfg, gold), training=False)
    for k in ("psm", "rm"):
        assert np.abs(out[k].numpy() - gold["eval_" + k]).max() < 1e-5, k
    assert abs(float(out["com"]) - float(gold["eval_com"])) < 1e-7 and out["comm_rate"] == int(gold["eval_comm_rate"])
    assert 0.0 < float(gold["eval_com"]) < 1.0          # the fixture exercises a non-trivial, resized mask
    try:
        model(dict())
    except Exception as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("forward on CPU parameters must raise")


def test_legacy_loss_oracle_matches_reference_golden():
    """O.point_pillar_loss (the 1-class loss of the legacy models) == the recorded value / gradients of the real
    reference PointPillarLoss (scripts/make_golden_legacy_loss.py) — groundwork for the legacy training step"""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import make_golden_legacy_loss as G

    gold = np.load(os.path.join(GOLD, "pploss.npz"))
    psm, rm, lab = G.inputs()
    a, b = psm.clone().requires_grad_(True), rm.clone().requires_grad_(True)
    tot, reg, conf = O.point_pillar_loss({"psm": a, "rm": b}, lab, 1.0, 2.0)
    tot.backward()
    assert abs(float(tot) - float(gold["total"])) < 1e-9 * float(gold["total"])
    assert abs(float(reg) - float(gold["reg"])) < 1e-6 and abs(float(conf) - float(gold["conf"])) < 1e-4
    assert np.abs(a.grad.numpy() - gold["dpsm"]).max() < 1e-7
    assert np.abs(b.grad[:, :, ::4, ::4].numpy() - gold["drm_sample"]).max() < 1e-7
