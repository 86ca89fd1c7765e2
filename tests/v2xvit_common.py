"""Shared helpers for the V2X-ViT parity tests (oracle = checker; product path = CUDA)."""
import json
import math
import os

import numpy as np
import torch

from oracle import w2c_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_small():
    cfg = json.load(open(os.path.join(GOLDEN_DIR, "v2xvit_small_config.json")))
    gold = np.load(os.path.join(GOLDEN_DIR, "v2xvit_small.npz"), allow_pickle=False)
    return cfg, gold


def golden_state_dict(model, gold):
    """Same deterministic parameters scripts/make_golden_v2xvit.py loaded into the reference model (the RTE sinusoid
    table keeps its analytic values)."""
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = O.det_init_state_dict(shapes, seed=int(gold["param_seed"]))
    sd = {k: v for k, v in sd.items() if not k.endswith("rte.emb.emb.weight")}
    full = {k: v.clone() for k, v in model.state_dict().items()}
    full.update(sd)
    return full


def scene_extras(agents, L):
    """the prior_encoding / spatial_correction_matrix of scripts/make_golden_v2xvit.py"""
    prior = torch.zeros(1, L, 3)
    scm = torch.eye(4, dtype=torch.float64).repeat(1, L, 1, 1)
    for i, t in enumerate(agents):
        prior[0, i] = torch.tensor([0.1 * i, float(i % 3), 1.0 if t == "rsu" else 0.0])
    a = 0.2
    scm[0, 1, :2, :2] = torch.tensor([[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]], dtype=torch.float64)
    scm[0, 1, 0, 3], scm[0, 1, 1, 3] = 6.0, -3.0
    scm[0, 2, 0, 3], scm[0, 2, 1, 3] = -4.8, 1.6
    return prior, scm


def golden_scene(cfg, gold):
    agents = [str(a) for a in gold["agents"]]
    dd = O.make_scene(cfg["preprocess"], agents, int(gold["n_points"]), int(gold["scene_seed"]),
                      cfg["preprocess"]["args"]["max_voxel_train"])
    dd["prior_encoding"], dd["spatial_correction_matrix"] = scene_extras(agents, int(gold["max_cav_num"]))
    return dd
