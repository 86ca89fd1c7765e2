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


def test_legacy_train_mode_oracle_matches_reference_golden():
    """train-mode forward (batch-statistic BN, top-K mask with the reference's `random` draws) + PointPillarLoss + autograd
    of the oracle == the recorded run of the real reference `point_pillar_where2comm` (scripts/make_golden_legacy_train.py):
    groundwork for the legacy training step on the kernels"""
    import random
    import sys

    import a2x_import

    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import make_golden_legacy_train as G

    M = a2x_import.pkg("opencood.models.point_pillar_where2comm")
    cfg, gold = load()
    tg = np.load(os.path.join(GOLD, "ppw2c_train_small.npz"))
    model = M.PointPillarWhere2comm(cfg["model_args"])
    sd = golden_state_dict(model, gold)
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "gaussian" not in k else v.clone())
         for k, v in sd.items()}
    torch.set_num_threads(8)
    random.seed(int(tg["k_seed"]))
    out, _ = O.pp_where2comm_forward(p, cfg["model_args"], golden_scene(cfg, gold), training=True)
    lab = G.labels(out["psm"].shape[2], out["psm"].shape[3], cfg["model_args"]["anchor_number"])
    loss = O.point_pillar_loss(out, lab, 1.0, 2.0)[0]
    loss.backward()
    for k in ("psm", "rm"):
        assert np.abs(out[k].detach().numpy() - tg["train_" + k]).max() < 1e-5, k
    assert abs(float(loss.detach()) - float(tg["loss"])) < 1e-6 * float(tg["loss"])
    assert abs(float(out["com"]) - float(tg["train_com"])) < 1e-7
    n = 0
    for name in [k[5:] for k in tg.files if k.startswith("grad_")]:
        g = p[name].grad
        got = g.flatten()[:: max(1, g.numel() // 256)][:256].numpy()
        ref = tg["grad_" + name]
        assert np.abs(got - ref).max() <= 1e-4 * (np.abs(ref).max() + 1e-30) + 1e-9, name
        n += 1
    assert n == 77
