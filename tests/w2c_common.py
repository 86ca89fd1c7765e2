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


def engine_buf(model, name):
    for k, t in model.engine.bufs.items():
        if len(k) == 3 and k[0] == name:
            return t
    raise KeyError(name)


def oracle_train_step(sd, cfg, dd, labels, k_seed, mask_override=None):
    """The oracle's train-mode forward + PointPillarLossMultiClass + autograd backward on the CPU. With
    `mask_override` (the [N,h,w] communication mask the CUDA path selected) the top-K tie-breaks are teacher-forced.
    Returns (outputs, total loss, {param: grad}, keep)."""
    args = cfg["model_args"]
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "gaussian" not in k
             and "num_batches" not in k else v.clone()) for k, v in sd.items()}
    keep = {"mask_override": mask_override}
    random.seed(k_seed)
    out, bufs = O.where2com_forward(p, args, dd, training=True, keep=keep)
    loss = O.point_pillar_loss_multiclass(out, labels, args["num_class"], cfg["loss_args"]["cls_weight"],
                                          cfg["loss_args"]["reg"])[0]
    loss.backward()
    grads = {k: v.grad for k, v in p.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
    return out, loss.detach(), grads, keep


def check_mask_ties(mask_gpu, keep, tol=1e-4, max_frac=5e-3):
    """The CUDA mask may differ from the oracle's own only where the smoothed confidence ties (within `tol`) with the
    per-agent cut (the K-th largest value): returns the number of differing pixels."""
    own = keep.get("mask_own", keep["mask"]).reshape(mask_gpu.shape)
    diff = own != mask_gpu
    n = int(diff.sum())
    if n == 0:
        return 0
    assert n <= max_frac * mask_gpu.numel(), "mask differs at %d pixels" % n
    smooth = keep["smooth"].reshape(mask_gpu.shape[0], -1)
    d = diff.reshape(mask_gpu.shape[0], -1)
    m = own.reshape(mask_gpu.shape[0], -1)
    for a in range(mask_gpu.shape[0]):
        if not d[a].any():
            continue
        sel = smooth[a][m[a] > 0]
        if sel.numel() == 0 or a == 0 and bool((m[a] > 0).all()):
            continue  # ego row is forced to one
        cut = float(sel.min())
        assert float((smooth[a][d[a]] - cut).abs().max()) < tol, "agent %d: mask differs away from the top-K cut" % a
    return n


def make_batch(preprocess, scenes, n_points, seed, max_voxels, sigma_xy=(10.0, 5.0)):
    """B ragged scenes in the reference's collate layout (intermediate_fusion_dataset.py:833-868): per agent type one
    merged voxel batch over all scenes (agent index = position within the type, scene-major), `record_len` per scene,
    `batch_idxs` = the scenes that contain the type. scenes: list of agent-type lists (vehicles, RSUs, drones order).
    Also returns the raw clouds (scene-major agent order) for the raw-point boundary."""
    from oracle import voxelize as V

    rng = preprocess["cav_lidar_range"]
    per_type = {t: [] for t in O.AGENT_TYPES}
    rl = {t: [] for t in O.AGENT_TYPES}
    clouds, k = [], 0
    for b, agents in enumerate(scenes):
        first = True
        for t in O.AGENT_TYPES:
            n_t = sum(1 for a in agents if a == t)
            rl[t].append(n_t)
            for _ in range(n_t):
                pts = O.synth_points(seed * 100 + k, n_points, rng, sigma_xy=sigma_xy)
                clouds.append(pts)
                p = V.mask_points(pts, rng, ego_box=first)
                first = False
                per_type[t].append(V.voxelize(p, rng, preprocess["args"]["voxel_size"],
                                              preprocess["args"]["max_points_per_voxel"], max_voxels))
                k += 1
    dd = {}
    for t in O.AGENT_TYPES:
        idxs = [b for b, n in enumerate(rl[t]) if n > 0]
        if per_type[t]:
            col = V.collate(per_type[t])
            feats = {k2: torch.from_numpy(v) for k2, v in col.items()}
        else:
            feats = None
        dd[t] = {"batch_merged_lidar_features_torch": feats, "record_len": torch.tensor(rl[t], dtype=torch.int32),
                 "batch_idxs": idxs}
    B = len(scenes)
    dd["img_pairwise_t_matrix_collab"] = torch.eye(4).view(1, 1, 1, 4, 4).repeat(B, 15, 15, 1, 1)
    dd["record_len"] = torch.tensor([len(a) for a in scenes], dtype=torch.int32)
    offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    raw = {"raw_points": {"points": torch.from_numpy(np.concatenate(clouds, 0)), "offsets": torch.from_numpy(offs),
                          "preprocess": preprocess, "filter": True}}
    for t in O.AGENT_TYPES:
        raw[t] = {"record_len": rl[t], "batch_idxs": dd[t]["batch_idxs"]}
    return dd, raw
