"""CPU (-m "not gpu"): the V2X-ViT oracle against the golden vectors recorded from the REAL reference, the exactness of
skipping padded agents, the host-side warp geometry, and the drop-in module's registry surface."""
import json
import os

import numpy as np
import torch

import v2xvit_common as VC
from oracle import v2xvit_oracle as VO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pkg(name):
    import a2x_import

    return a2x_import.pkg(name)


def test_oracle_matches_reference_golden_and_padding_is_skippable():
    cfg, gold = VC.load_small()
    M = _pkg("opencood.models.airv2x_v2xvit")
    model = M.Airv2xV2XVit(cfg["model_args"])
    sd = VC.golden_state_dict(model, gold)
    assert np.abs(sd["fusion_net.encoder.rte.emb.emb.weight"].numpy() - gold["rte_table"]).max() < 1e-7
    dd = VC.golden_scene(cfg, gold)
    torch.set_num_threads(8)
    with torch.no_grad():
        out, _ = VO.v2xvit_forward(sd, cfg["model_args"], dd, training=False)
    for k in ("psm", "rm", "obj"):
        assert np.abs(out[k].numpy() - gold["eval_" + k]).max() < 2e-5, k
    assert out["comm_rate"] == int(gold["eval_comm_rate"])
    # the CUDA path runs on the valid agents only: same weights with max_cav = #agents
    args = json.loads(json.dumps(cfg["model_args"]))
    args["max_cav"] = {"vehicle": 2, "rsu": 1, "drone": 1}
    dd2 = dict(dd)
    dd2["prior_encoding"] = dd["prior_encoding"][:, :4]
    dd2["spatial_correction_matrix"] = dd["spatial_correction_matrix"][:, :4]
    with torch.no_grad():
        out2, _ = VO.v2xvit_forward(sd, args, dd2, training=False)
    for k in ("psm", "rm", "obj"):
        assert float((out[k] - out2[k]).abs().max()) < 1e-5, k


def test_host_warp_geometry_matches_oracle():
    W = _pkg("warp")
    _, scm = VC.scene_extras(["vehicle", "vehicle", "rsu", "drone"], 6)
    H, Wd = 32, 64
    got = W.sttf_theta(scm, 0.4, 4, H, Wd)
    T = VO.transformation_matrix(VO.discretized_matrix(scm, 0.4, 4).reshape(-1, 2, 3), (H, Wd))
    want = VO.warp_theta(T, (H, Wd)).reshape(1, 6, 2, 3)
    assert float((got - want).abs().max()) < 1e-6
    assert float((got[0, 0] - torch.tensor([[1.0, 0, 0], [0, 1.0, 0]])).abs().max()) < 1e-6   # ego: identity
    pw = torch.eye(4).repeat(1, 2, 2, 1, 1)
    pw[0, 0, 1, 0, 3], pw[0, 0, 1, 0, 1] = 8.0, 0.5
    n = W.normalize_pairwise(pw, 50, 176, 2, 0.4)
    assert abs(float(n[0, 0, 1, 0, 2]) - 8.0 / (2 * 0.4 * 176) * 2) < 1e-6 and abs(float(n[0, 0, 1, 0, 1]) - 0.5 * 50 / 176) < 1e-6


def test_registry_surface_full_config():
    cfg = json.load(open(os.path.join(ROOT, "configs", "airv2x_intermediate_v2xvit.json")))
    M = _pkg("opencood.models.airv2x_v2xvit")
    target = "airv2x_v2xvit".replace("_", "")
    cls = [v for k, v in vars(M).items() if k.lower() == target]
    assert len(cls) == 1 and cls[0] is M.Airv2xV2XVit
    model = M.Airv2xV2XVit(cfg["model_args"])
    assert sum(p.numel() for p in model.parameters()) == 12758879
    sd = model.state_dict()
    assert sd["fusion_net.encoder.layers.2.0.layers.0.0.fn.relation_att"].shape == (4, 8, 32, 32)
    assert sd["fusion_net.encoder.layers.1.0.layers.0.1.fn.pwmsa.0.pos_embedding"].shape == (3, 3)
    assert sd["fusion_net.encoder.layers.0.0.layers.0.1.fn.split_attn.fc2.weight"].shape == (768, 256)
    assert sd["fusion_net.encoder.prior_feed.weight"].shape == (256, 259)
    assert sd["fusion_net.encoder.rte.emb.emb.weight"].shape == (100, 256)
    try:
        model(dict())
    except Exception as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("forward on CPU parameters must raise")
