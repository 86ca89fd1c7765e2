"""CPU (-m "not gpu"): the CoBEVT oracle against the golden vectors recorded from the REAL reference, and the drop-in
module's registry surface (class name, state_dict keys / shapes / parameter count)."""
import json
import os

import numpy as np
import torch

import cobevt_common as CC
from oracle import cobevt_oracle as CO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(args):
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_cobevt")
    return M.Airv2xCoBEVT(args)


def test_oracle_matches_reference_golden():
    cfg, gold = CC.load_small()
    model = _model(cfg["model_args"])
    sd = CC.golden_state_dict(model, gold)
    dd = CC.golden_scene(cfg, gold)
    torch.set_num_threads(8)
    with torch.no_grad():
        out, _ = CO.cobevt_forward(sd, cfg["model_args"], dd, training=False)
    for k in ("psm", "rm", "obj"):
        assert np.abs(out[k].numpy() - gold["eval_" + k]).max() < 1e-5, k
    # fusion network alone, ragged agent mask (3 and 5 real agents of L = 7)
    g = torch.Generator().manual_seed(int(gold["fusion_seed"]))
    x = torch.randn(2, 7, 256, 8, 16, generator=g)
    mask = torch.tensor([[1, 1, 1, 0, 0, 0, 0], [1, 1, 1, 1, 1, 0, 0]])
    x = x * mask[:, :, None, None, None]
    com = mask[:, None, None, None, :].expand(-1, 8, 16, 1, -1)
    with torch.no_grad():
        o = CO.swap_fusion_encoder(sd, cfg["model_args"]["fax_fusion"], x, com)
    assert np.abs(o.numpy() - gold["fusion_out"]).max() < 1e-5


def test_oracle_naive_compressor_matches_reference_golden():
    """a10: `compression: 2` inserts NaiveCompressor after the shrink header (airv2x_cobevt.py:121-123)"""
    cfg, gold = CC.load_small()
    args = json.loads(json.dumps(cfg["model_args"]))
    args["compression"] = 2
    model = _model(args)
    assert model.state_dict()["naive_compressor.encoder.0.weight"].shape == (128, 256, 3, 3)
    assert model.state_dict()["naive_compressor.decoder.3.weight"].shape == (256, 256, 3, 3)
    sd = CC.golden_state_dict_compressed(model, gold)
    torch.set_num_threads(8)
    with torch.no_grad():
        out, _ = CO.cobevt_forward(sd, args, CC.golden_scene(cfg, gold), training=False)
    for k in ("psm", "rm", "obj"):
        assert np.abs(out[k].numpy() - gold["cmp2_eval_" + k]).max() < 1e-5, k


def test_registry_surface_full_config():
    """create_model's lookup rule (tools/train_utils.py:302-325) + the reference's parameter inventory"""
    cfg = json.load(open(os.path.join(ROOT, "configs", "airv2x_intermediate_cobevt.json")))
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_cobevt")
    target = "airv2x_cobevt".replace("_", "")
    cls = [v for k, v in vars(M).items() if k.lower() == target]
    assert len(cls) == 1 and cls[0] is M.Airv2xCoBEVT
    model = M.Airv2xCoBEVT(cfg["model_args"])
    assert sum(p.numel() for p in model.parameters()) == 9741454
    sd = model.state_dict()
    assert sd["fusion_net.layers.0.window_attention.fn.to_qkv.weight"].shape == (768, 256)
    assert sd["fusion_net.layers.2.grid_attention.fn.relative_position_bias_table.weight"].shape == (13 * 49, 8)
    assert sd["fusion_net.layers.1.grid_ffd.fn.net.3.bias"].shape == (256,)
    assert sd["fusion_net.mlp_head.3.weight"].shape == (256, 256)
    assert torch.equal(sd["fusion_net.layers.0.window_attention.fn.relative_position_index"],
                       CO.relative_position_index(7, 4))
    try:
        model(dict())
    except Exception as e:  # no CPU path: must fail loudly, never fall back
        assert "CUDA" in str(e) or "cuda" in str(e)
    else:
        raise AssertionError("forward on CPU parameters must raise")


def test_compressor_train_mode_oracle_matches_reference_golden():
    """NaiveCompressor in train mode (batch-statistic BN, backward): oracle == the recorded run of the real reference
    module (scripts/make_golden_compressor_train.py) — groundwork for training with compression > 0"""
    import os
    import sys

    import numpy as np
    import torch

    from oracle import cobevt_oracle as CO, w2c_oracle as O

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "scripts"))
    import make_golden_compressor_train as G

    g = np.load(os.path.join(root, "tests", "golden", "compressor_train.npz"))
    c, r = int(g["c"]), int(g["r"])
    shapes = {}
    for conv, bn, ci, co in (("encoder.0", "encoder.1", c, c // r), ("decoder.0", "decoder.1", c // r, c), ("decoder.3", "decoder.4", c, c)):
        shapes[conv + ".weight"], shapes[conv + ".bias"] = (co, ci, 3, 3), (co,)
        for k, shp in (("weight", (co,)), ("bias", (co,)), ("running_mean", (co,)), ("running_var", (co,)), ("num_batches_tracked", ())):
            shapes["%s.%s" % (bn, k)] = shp
    sd = {"naive_compressor." + k: v for k, v in O.det_init_state_dict(shapes, seed=int(g["seed"])).items()}
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    x, w = G.inputs()
    x = x.requires_grad_(True)
    y = CO.naive_compressor(p, x, True, {})
    (y * w).sum().backward()
    assert np.abs(y.detach()[:, ::16, ::2, ::2].numpy() - g["y"]).max() < 1e-5
    assert np.abs(x.grad[:, ::16, ::2, ::2].numpy() - g["dx"]).max() < 1e-5
    ref = g["dw_enc"]
    assert np.abs(p["naive_compressor.encoder.0.weight"].grad[::8, ::16].numpy() - ref).max() < 1e-4 * np.abs(ref).max()
