"""Shared helpers for the CoBEVT parity tests (oracle = checker; product path = CUDA)."""
import json
import os

import numpy as np

from oracle import w2c_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_small():
    cfg = json.load(open(os.path.join(GOLDEN_DIR, "cobevt_small_config.json")))
    gold = np.load(os.path.join(GOLDEN_DIR, "cobevt_small.npz"), allow_pickle=False)
    return cfg, gold


def golden_state_dict(model, gold):
    """Same deterministic parameters scripts/make_golden_cobevt.py loaded into the reference model."""
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if "relative_position_index" not in k}
    sd = O.det_init_state_dict(shapes, seed=int(gold["param_seed"]))
    full = {k: v.clone() for k, v in model.state_dict().items()}
    full.update(sd)
    return full


def golden_scene(cfg, gold):
    agents = [str(a) for a in gold["agents"]]
    return O.make_scene(cfg["preprocess"], agents, int(gold["n_points"]), int(gold["scene_seed"]),
                        cfg["preprocess"]["args"]["max_voxel_train"])


def golden_state_dict_compressed(model, gold):
    """parameters of the `compression: 2` golden run: the base model's seeded parameters + separately seeded compressor"""
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()
              if "relative_position_index" not in k and not k.startswith("naive_compressor")}
    sd = O.det_init_state_dict(shapes, seed=int(gold["param_seed"]))
    sh2 = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("naive_compressor")}
    sd.update(O.det_init_state_dict(sh2, seed=int(gold["cmp_param_seed"])))
    full = {k: v.clone() for k, v in model.state_dict().items()}
    full.update(sd)
    return full
