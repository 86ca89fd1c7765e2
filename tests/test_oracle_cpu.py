"""CPU: the oracle restatement against the golden vectors produced by the REAL reference (scripts/make_golden.py),
and the C voxeliser restatement against its pure-Python twin + edge cases."""
import random

import numpy as np
import torch

import w2c_common as C
from oracle import voxelize as V, w2c_oracle as O


def test_oracle_eval_matches_reference_golden(pkg):
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    cfg, gold = C.load_small()
    args = cfg["model_args"]
    model = M.Airv2xWhere2com(args)  # parameter container only (CPU); supplies the reference state_dict key set
    sd = C.golden_state_dict(model, gold)
    dd = C.golden_scene(cfg, gold)
    keep = {}
    with torch.no_grad():
        out, _ = O.where2com_forward(sd, args, dd, training=False, keep=keep)
    for k in ("psm", "rm", "obj"):
        assert np.abs(out[k].numpy() - gold["eval_" + k]).max() < 1e-5, k
    assert abs(float(out["com"]) - float(gold["eval_com"])) < 1e-7
    assert out["comm_rate"] == int(gold["eval_comm_rate"])
    for k in ("spatial_features", "spatial_features_2d", "psm_single", "mask", "fused_l0", "fused_l2", "fused_feature"):
        assert np.abs(C.sample(keep[k]) - gold["eval_keep_" + k]).max() < 1e-5, k


def test_state_dict_keys_match_reference(pkg):
    """every parameter the reference's gradients were recorded for exists with the same name in the drop-in module"""
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    cfg, gold = C.load_small()
    model = M.Airv2xWhere2com(cfg["model_args"])
    names = {n for n, _ in model.named_parameters()}
    ref = {k[len("grad_"):] for k in gold.files if k.startswith("grad_")}
    assert ref <= names, sorted(ref - names)
    assert sum(p.numel() for p in model.parameters()) == 7276088  # Airv2xWhere2com parameter count
    bufs = {n for n, _ in model.named_buffers()}
    assert {k[len("buf_"):] for k in gold.files if k.startswith("buf_")} <= bufs


def test_voxelize_c_matches_python_twin():
    rng = np.random.default_rng(0)
    r = [-25.6, -12.8, -3, 25.6, 12.8, 1]
    vs = [0.4, 0.4, 4]
    pts = np.concatenate([rng.uniform(-30, 30, (2500, 2)), rng.uniform(-4, 2, (2500, 1)), rng.uniform(0, 1, (2500, 1))], 1)
    pts = pts.astype(np.float32)
    a = V.voxelize(pts, r, vs, 32, 300)
    b = V.voxelize_py(pts, r, vs, 32, 300)
    assert a["voxel_features"].shape[0] == 300  # cap hit
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_voxelize_edge_cases():
    r = [-25.6, -12.8, -3, 25.6, 12.8, 1]
    vs = [0.4, 0.4, 4]
    # empty cloud
    e = V.voxelize(np.zeros((0, 4), np.float32), r, vs)
    assert e["voxel_features"].shape == (0, 32, 4)
    # 100 points in one pillar: first 32 kept in input order
    p = np.tile(np.array([[0.1, 0.1, 0.0, 0.0]], np.float32), (100, 1))
    p[:, 3] = np.arange(100)
    o = V.voxelize(p, r, vs)
    assert o["voxel_num_points"].tolist() == [32]
    assert o["voxel_features"][0, :, 3].tolist() == list(range(32))
    # boundaries: lower edge is cell 0, upper edge is outside, NaN dropped
    p = np.array([[-25.6, -12.8, -3, 1], [25.6, 0, 0, 2], [0, 12.8, 0, 3], [np.nan, 0, 0, 4], [25.59, 12.79, 0.99, 5]], np.float32)
    o = V.voxelize(p, r, vs)
    assert o["voxel_coords"].tolist() == [[0, 0, 0], [0, 63, 127]]
    # point filters (utils/pcd_utils.py)
    q = np.array([[0, 0, 0, 0], [2.95, 1.1, 0, 0], [2.96, 0, 0, 0], [-25.6, 0, 0, 0]], np.float32)
    m = V.mask_points(q, r, ego_box=True)
    assert m[:, 0].tolist() == [np.float32(2.96)]
