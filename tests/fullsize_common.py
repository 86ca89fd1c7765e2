"""Recipes shared by scripts/make_golden_full.py (real reference, container side) and tests/test_gpu_fullsize.py /
tests/test_fullsize_cpu.py: the full-size synthetic scenes (BASELINE.json configs 2-5: 60 000 points per agent, the
bench's x~N(0,35) / y~N(0,15) clouds), the seeded parameters, and what of an output map a fixture keeps."""
import math

import numpy as np
import torch

from oracle import w2c_oracle as O

AGENTS = ["vehicle", "vehicle", "rsu", "rsu", "drone"]     # configs 2, 3: 5 agents
N_POINTS = 60000
SIGMA_XY = (35.0, 15.0)
SCENE_SEED = 21
LABEL_SEED = 17
K_SEED = 9
W2C_PARAM_SEED, W2C_CLS_SHIFT = 1234, -5.0
COBEVT_PARAM_SEED = 4321
V2XVIT_PARAM_SEED = 2468
RANGE_504 = (-100.8, -100.8, 100.8, 100.8)                 # config 5: 504 x 504 pillars at 0.4 m
# config 4: (fixture prefix, agents, max_cav override). The shipped yaml caps L at 3 + 2 + 2 = 7
# (airv2x_intermediate_cobevt.yaml:22-25); BASELINE's "8 agents" needs max_cav 3/3/2.
COBEVT_CASES = (
    ("a5", AGENTS, None),
    ("a7", ["vehicle"] * 3 + ["rsu"] * 2 + ["drone"] * 2, None),
    ("a8", ["vehicle"] * 3 + ["rsu"] * 3 + ["drone"] * 2, {"vehicle": 3, "rsu": 3, "drone": 2}),
)


def scene(preprocess, training, agents=None, seed=SCENE_SEED, n_points=N_POINTS):
    """pre-voxelised data_dict of one scene (the reference's collate layout); eval uses max_voxel_test (70 000, no cap
    hit), training max_voxel_train (32 000: the cap drops pillars)"""
    mv = preprocess["args"]["max_voxel_train" if training else "max_voxel_test"]
    return O.make_scene(preprocess, list(agents or AGENTS), n_points, seed, mv, sigma_xy=SIGMA_XY)


def sample(t, n=4096):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].double().cpu().numpy()


def edge(t):
    """the last two rows and last two columns of an NCHW map: the cells partial GEMM tiles write"""
    t = t.detach()
    return torch.cat([t[..., -2:, :].reshape(-1), t[..., :, -2:].reshape(-1)]).double().cpu().numpy()


def v2xvit_extras(agents, L):
    """prior_encoding [v/30, time_delay, infra] and spatial_correction_matrix (one neighbour mis-aligned by 0.2 rad /
    (6, -3) m, another by a pure translation), as in scripts/make_golden_v2xvit.py"""
    prior = torch.zeros(1, L, 3)
    scm = torch.eye(4, dtype=torch.float64).repeat(1, L, 1, 1)
    for i, t in enumerate(agents):
        prior[0, i] = torch.tensor([0.1 * i, float(i % 3), 1.0 if t == "rsu" else 0.0])
    a = 0.2
    scm[0, 1, :2, :2] = torch.tensor([[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]], dtype=torch.float64)
    scm[0, 1, 0, 3], scm[0, 1, 1, 3] = 6.0, -3.0
    scm[0, 2, 0, 3], scm[0, 2, 1, 3] = -4.8, 1.6
    return prior, scm


def seeded_state_dict(model, seed, skip=(), cls_shift=0.0):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if not any(s in k for s in skip)}
    sd = O.det_init_state_dict(shapes, seed=seed)
    full = {k: v.clone() for k, v in model.state_dict().items()}
    full.update(sd)
    if cls_shift:
        full["cls_head.bias"] = full["cls_head.bias"] + cls_shift
    return full


def compare_with_golden(out, gold, prefix, tol):
    """max abs difference between an output dict and a fixture's strided sample + edge cells"""
    worst = 0.0
    for k in ("psm", "rm", "obj"):
        t = out[k].detach().float().cpu()
        assert tuple(t.shape) == tuple(int(v) for v in gold[prefix + k + "_shape"]), (k, t.shape)
        e1 = float(np.abs(sample(t) - gold[prefix + k + "_sample"]).max())
        e2 = float(np.abs(edge(t) - gold[prefix + k + "_edge"]).max())
        assert e1 < tol and e2 < tol, (prefix + k, e1, e2)
        worst = max(worst, e1, e2)
    return worst


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    return o
