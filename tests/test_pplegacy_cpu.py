"""CPU (-m "not gpu"): the legacy transformer-fusion models `point_pillar_cobevt` / `point_pillar_v2xvit` (3 agents, 8k
points, 128 x 128 pillars): the oracles against the golden vectors recorded from the REAL reference
(scripts/make_golden_legacy_fusion.py), and the drop-in modules' registry surface."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import cobevt_oracle as CO, v2xvit_oracle as VO, w2c_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {"ppcobevt": ("point_pillar_cobevt", "PointPillarCoBEVT", 10513344, CO.pp_cobevt_forward),
         "ppv2xvit": ("point_pillar_v2xvit", "PointPillarV2XVit", 13543561, VO.pp_v2xvit_forward)}


def load(name):
    cfg = json.load(open(os.path.join(GOLD, name + "_small_config.json")))
    return cfg, np.load(os.path.join(GOLD, name + "_small.npz"))


def build(name):
    import a2x_import

    mod, cls, _, _ = CASES[name]
    M = a2x_import.pkg("opencood.models." + mod)
    found = [v for k, v in vars(M).items() if k.lower() == mod.replace("_", "")]
    assert len(found) == 1 and found[0] is getattr(M, cls)           # train_utils.create_model's lookup rule
    cfg, gold = load(name)
    return getattr(M, cls)(cfg["model_args"]), cfg, gold


def golden_state_dict(model, gold):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if "relative_position_index" not in k}
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd.update({k: v for k, v in O.det_init_state_dict(shapes, seed=int(gold["param_seed"])).items()
               if not k.endswith("rte.emb.emb.weight")})
    return sd


def pairwise(L):
    """the poses of scripts/make_golden_legacy_fusion.py"""
    t = torch.eye(4).view(1, 1, 1, 4, 4).repeat(1, L, L, 1, 1)
    a = 0.15
    t[0, 0, 1, :2, :2] = torch.tensor([[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]])
    t[0, 0, 1, 0, 3], t[0, 0, 1, 1, 3] = 4.8, -1.6
    t[0, 0, 2, 0, 3], t[0, 0, 2, 1, 3] = -3.2, 2.4
    return t


def train_labels(H, W, A, seed=99):
    """the planted labels of scripts/make_golden_legacy_fusion.py (train-mode pin)"""
    g = torch.Generator().manual_seed(seed)
    pos = torch.zeros(1, H, W, A, dtype=torch.float64)
    pos.view(-1)[torch.randperm(H * W * A, generator=g)[:20]] = 1.0
    tg = 0.3 * torch.randn(1, H, W, 7 * A, generator=g, dtype=torch.float64) * pos.repeat_interleave(7, -1)
    return {"pos_equal_one": pos, "targets": tg}


def oracle_train(name, sd, cfg, gold):
    """train-mode forward + PointPillarLoss + autograd through the oracle: (loss, logits, {name: grad}, running stats)"""
    args = cfg["model_args"]
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    torch.set_num_threads(8)
    out, bufs = CASES[name][3](p, args, golden_scene(cfg, gold), training=True)
    lab = train_labels(out["psm"].shape[2], out["psm"].shape[3], args["anchor_number"])
    loss = O.point_pillar_loss(out, lab, 1.0, 2.0)[0]
    loss.backward()
    return loss.detach(), out, {k: v.grad for k, v in p.items() if torch.is_tensor(v) and v.grad is not None}, bufs, lab


def golden_scene(cfg, gold):
    dd = O.make_scene_legacy(cfg["preprocess"], int(gold["n_agents"]), int(gold["n_points"]), int(gold["scene_seed"]),
                             cfg["preprocess"]["args"]["max_voxel_test"])
    dd["pairwise_t_matrix"] = pairwise(cfg["model_args"]["max_cav"])
    return dd


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_and_registry_surface(name):
    model, cfg, gold = build(name)
    assert sum(p.numel() for p in model.parameters()) == CASES[name][2]
    sd = golden_state_dict(model, gold)
    assert sd["shrink_conv.layers.0.double_conv.0.weight"].shape == (256, 384, 3, 3)
    assert sd["pillar_vfe.pfn_layers.0.linear.weight"].shape == (64, 10) and sd["cls_head.weight"].shape == (2, 256, 1, 1)
    torch.set_num_threads(8)
    with torch.no_grad():
        out, _ = CASES[name][3](sd, cfg["model_args"], golden_scene(cfg, gold), training=False)
    for k in ("psm", "rm"):
        assert np.abs(out[k].numpy() - gold["eval_" + k]).max() < 2e-5, k
    assert out["comm_rate"] == int(gold["eval_comm_rate"])
    with pytest.raises(RuntimeError, match="CUDA"):
        model(golden_scene(cfg, gold))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_train_mode_matches_reference_golden(name):
    """train mode (batch-statistic BatchNorm, dropout p = 0): the oracle's loss, logits and every parameter-gradient norm
    against the values recorded from the REAL reference model + the reference's PointPillarLoss"""
    model, cfg, gold = build(name)
    loss, out, grads, _, _ = oracle_train(name, golden_state_dict(model, gold), cfg, gold)
    assert abs(float(loss) - float(gold["train_loss"])) < 1e-5 * abs(float(gold["train_loss"]))
    assert np.abs(out["psm"].detach().numpy() - gold["train_psm"]).max() < 5e-5
    for n, ref in zip(gold["train_grad_names"], gold["train_grad_norms"]):
        got = float(grads[str(n)].norm())
        assert abs(got - float(ref)) < 2e-3 * float(ref) + 1e-6, (n, got, float(ref))
