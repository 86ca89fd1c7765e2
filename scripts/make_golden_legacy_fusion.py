"""Generate tests/golden/pp{cobevt,v2xvit}_small.npz: the REAL reference `point_pillar_cobevt` / `point_pillar_v2xvit`
(V2XR_cobevt.yaml / V2XR_v2xvit.yaml args; 3 agents, 8k points, 128 x 128 pillars) on the CPU, eval mode, checked against
the oracle restatements. One neighbour gets a non-identity pairwise pose so the V2X-ViT ego-warp is exercised.

    python scripts/make_golden_legacy_fusion.py
"""
import json
import math
import os
import re
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cobevt_oracle as CO, ref_import, v2xvit_oracle as VO, w2c_oracle as O  # noqa: E402

RANGE = [-25.6, -25.6, -3, 25.6, 25.6, 1]      # 128 x 128 pillars at 0.4 m
N_AGENTS, N_POINTS, SCENE_SEED, PARAM_SEED = 3, 8000, 31, 2468


def pairwise(L):
    """agent 1 sits 4.8 m ahead / 1.6 m left of the ego with a 0.15 rad heading difference; the rest identity"""
    t = torch.eye(4).view(1, 1, 1, 4, 4).repeat(1, L, L, 1, 1)
    a = 0.15
    t[0, 0, 1, :2, :2] = torch.tensor([[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]])
    t[0, 0, 1, 0, 3], t[0, 0, 1, 1, 3] = 4.8, -1.6
    t[0, 0, 2, 0, 3], t[0, 0, 2, 1, 3] = -3.2, 2.4
    return t


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


def run(name, yaml_rel, oracle_forward):
    src = open(os.path.join(ref_import.REF_ROOT, "opencood", "hypes_yaml", yaml_rel)).read()
    out_y = re.sub(r"(cav_lidar_range: &cav_lidar )\[[^\]]+\]", r"\1" + str(RANGE), src)
    out_y = re.sub(r"(voxel_size: &voxel_size )\[[^\]]+\]", r"\1[0.4, 0.4, 4]", out_y)
    assert out_y != src
    p = os.path.join(tempfile.mkdtemp(), "small.yaml")
    open(p, "w").write(out_y)
    from opencood.hypes_yaml import yaml_utils
    hypes = yaml_utils.load_yaml(p)
    args = hypes["model"]["args"]
    model = ref_import.create_model(hypes)
    print(name, "params", sum(q.numel() for q in model.parameters()))
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if "relative_position_index" not in k}
    sd = O.det_init_state_dict(shapes, seed=PARAM_SEED)
    sd = {k: v for k, v in sd.items() if not k.endswith("rte.emb.emb.weight")}      # keep the analytic sinusoid table
    full = model.state_dict()
    full.update(sd)
    model.load_state_dict(full)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    dd = O.make_scene_legacy(hypes["preprocess"], N_AGENTS, N_POINTS, SCENE_SEED, hypes["preprocess"]["args"]["max_voxel_test"])
    dd["pairwise_t_matrix"] = pairwise(args["max_cav"])
    model.eval()
    with torch.no_grad():
        ref = model({k: (v.clone() if torch.is_tensor(v) else v) for k, v in dd.items()})
        ora, _ = oracle_forward(sd, args, dd, training=False)
    out = {"n_agents": N_AGENTS, "n_points": N_POINTS, "scene_seed": SCENE_SEED, "param_seed": PARAM_SEED}
    for k in ("psm", "rm"):
        err = float((ref[k] - ora[k]).abs().max())
        print("  eval %s %s: ref-vs-oracle %.3e (|ref| max %.3f)" % (k, tuple(ref[k].shape), err, float(ref[k].abs().max())))
        assert err < 2e-5
        out["eval_" + k] = ref[k].numpy()
    assert ref["comm_rate"] == ora["comm_rate"]
    out["eval_comm_rate"] = int(ref["comm_rate"])
    cfg = {"model_args": jsonable(args), "preprocess": jsonable(hypes["preprocess"]), "postprocess": jsonable(hypes["postprocess"]),
           "source": "opencood/hypes_yaml/%s (cav_lidar_range %s, voxel_size [0.4, 0.4, 4])" % (yaml_rel, RANGE)}
    json.dump(cfg, open(os.path.join(ROOT, "tests", "golden", name + "_small_config.json"), "w"), indent=1)
    dst = os.path.join(ROOT, "tests", "golden", name + "_small.npz")
    np.savez_compressed(dst, **out)
    print("  wrote", dst, "%.1f KB" % (os.path.getsize(dst) / 1024))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs("/tmp/a2x_golden/debug", exist_ok=True)
    os.chdir("/tmp/a2x_golden")
    ref_import.install()
    run("ppcobevt", "V2X-R/LiDAR/V2XR_cobevt.yaml", CO.pp_cobevt_forward)
    run("ppv2xvit", "V2X-R/LiDAR/V2XR_v2xvit.yaml", VO.pp_v2xvit_forward)


if __name__ == "__main__":
    main()
