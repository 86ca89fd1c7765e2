"""Generate tests/golden/pp{cobevt,v2xvit}_small.npz: the REAL reference `point_pillar_cobevt` / `point_pillar_v2xvit`
(V2XR_cobevt.yaml / V2XR_v2xvit.yaml args; 3 agents, 8k points, 128 x 128 pillars) on the CPU, eval mode AND train mode
(forward, PointPillarLoss, every parameter gradient, running statistics; dropout p = 0), checked against the oracle
restatements. One neighbour gets a non-identity pairwise pose so the V2X-ViT ego-warp is exercised.

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


def train_labels(H, W, A, seed=99):
    """20 planted positives with regression targets (the legacy collate's label_dict: loss/point_pillar_loss.py:77-100)"""
    g = torch.Generator().manual_seed(seed)
    pos = torch.zeros(1, H, W, A, dtype=torch.float64)
    pos.view(-1)[torch.randperm(H * W * A, generator=g)[:20]] = 1.0
    tg = 0.3 * torch.randn(1, H, W, 7 * A, generator=g, dtype=torch.float64) * pos.repeat_interleave(7, -1)
    return {"pos_equal_one": pos, "targets": tg}


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
    # ---- train mode (batch-statistic BatchNorm; every nn.Dropout of the real model set to p = 0, which is what the
    # oracle's legacy forwards restate): forward, the reference's own PointPillarLoss, all parameter gradients and the
    # updated running statistics, real reference vs oracle
    from opencood.loss.point_pillar_loss import PointPillarLoss
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.train()
    model.zero_grad()
    tr = model({k: (v.clone() if torch.is_tensor(v) else v) for k, v in dd.items()})
    lab = train_labels(tr["psm"].shape[2], tr["psm"].shape[3], args["anchor_number"])
    crit = PointPillarLoss({"cls_weight": 1.0, "reg": 2.0})
    loss_ref = crit({"psm": tr["psm"], "rm": tr["rm"]}, {k: v.float() for k, v in lab.items()})
    loss_ref.backward()
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    to, bufs = oracle_forward(p, args, dd, training=True)
    loss_ora = O.point_pillar_loss(to, lab, 1.0, 2.0)[0]
    loss_ora.backward()
    print("  train: psm ref-vs-oracle %.3e, loss %.6f vs %.6f" % (float((tr["psm"] - to["psm"]).abs().max()), float(loss_ref), float(loss_ora)))
    assert float((tr["psm"] - to["psm"]).abs().max()) < 5e-5 and abs(float(loss_ref) - float(loss_ora)) < 1e-5 * abs(float(loss_ref))
    worst, n_grads, names, norms = 0.0, 0, [], []
    for n, q in model.named_parameters():
        if q.grad is None:
            assert p[n].grad is None or float(p[n].grad.abs().max()) == 0.0, n
            continue
        e = float((q.grad - p[n].grad).norm() / (q.grad.norm() + 1e-30))
        if q.grad.norm() > 1e-6:
            worst = max(worst, e)
        n_grads += 1
        names.append(n)
        norms.append(float(q.grad.norm()))
    rs = max(float((model.state_dict()[k] - v).abs().max()) for k, v in bufs.items() if "num_batches" not in k)
    print("  train: %d parameter gradients, worst norm-wise difference %.3e; running statistics %.3e" % (n_grads, worst, rs))
    assert worst < 2e-3 and rs < 1e-5
    out["train_loss"] = float(loss_ref)
    out["train_psm"] = tr["psm"].detach().numpy()
    out["train_grad_names"] = np.array(names)
    out["train_grad_norms"] = np.array(norms)
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
