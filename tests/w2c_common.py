"""Shared helpers for the Where2comm parity tests (oracle = checker; product path = CUDA)."""
import json
import os
import random

import numpy as np
import torch

from oracle import w2c_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_small():
    cfg = json.load(open(os.path.join(GOLDEN_DIR, "w2c_small_config.json")))
    gold = np.load(os.path.join(GOLDEN_DIR, "w2c_small.npz"), allow_pickle=False)
    return cfg, gold


def state_dict_shapes(model):
    return {k: tuple(v.shape) for k, v in model.state_dict().items()}


def golden_state_dict(model, gold):
    """Same deterministic parameters scripts/make_golden.py loaded into the reference model."""
    sd = O.det_init_state_dict(state_dict_shapes(model), seed=int(gold["param_seed"]))
    full = {k: v.clone() for k, v in model.state_dict().items()}
    full.update(sd)
    full["cls_head.bias"] = full["cls_head.bias"] + float(gold["cls_bias_shift"])
    return full


def golden_scene(cfg, gold):
    agents = [str(a) for a in gold["agents"]]
    return O.make_scene(cfg["preprocess"], agents, int(gold["n_points"]), int(gold["scene_seed"]),
                        cfg["preprocess"]["args"]["max_voxel_train"])


def sample(t, n=4096):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].double().cpu().numpy()


def to_device(dd, dev):
    out = {}
    for k, v in dd.items():
        if isinstance(v, dict):
            out[k] = to_device(v, dev)
        elif torch.is_tensor(v):
            out[k] = v.to(dev)
        else:
            out[k] = v
    return out
