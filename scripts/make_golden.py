"""Generate tests/golden/w2c_small.npz by running the REAL reference (imported from /root/reference, CPU) and
checking the oracle restatement against it. Container-side only (the reference tree does not travel).

    python scripts/make_golden.py

Geometry is shrunk through the yaml's own keys (lidar ranges) so fixtures stay small; everything else is the
unmodified airv2x_intermediate_where2com.yaml. Parameters come from oracle.w2c_oracle.det_init_state_dict (seeded),
so the fixture stores inputs' seeds and outputs only.
"""
import os
import random
import re
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, voxelize as V, w2c_oracle as O  # noqa: E402

SMALL_RANGE_XY = (-25.6, -12.8, 25.6, 12.8)   # 128 x 64 pillars at 0.4 m
YAML = "airv2x/lidar/det/airv2x_intermediate_where2com.yaml"


def small_hypes():
    src = open(os.path.join(ref_import.REF_ROOT, "opencood", "hypes_yaml", YAML)).read()
    x0, y0, x1, y1 = SMALL_RANGE_XY

    def patch(m):
        vals = [v.strip() for v in m.group(2).split(",")]
        vals[0], vals[1], vals[3], vals[4] = str(x0), str(y0), str(x1), str(y1)
        return m.group(1) + "[" + ", ".join(vals) + "]"

    out = re.sub(r"((?:cav_lidar_range|lidar_range): &\w+ )\[([^\]]+)\]", patch, src)
    assert out != src
    d = tempfile.mkdtemp()
    p = os.path.join(d, "small.yaml")
    open(p, "w").write(out)
    ref_import.install()
    from opencood.hypes_yaml import yaml_utils
    return yaml_utils.load_yaml(p)


make_scene = lambda hypes, agents, n_points, seed, max_voxels: O.make_scene(hypes['preprocess'], agents, n_points, seed, max_voxels)
make_labels = O.make_labels


def sample(t, n=4096):
    """deterministic strided sample of a tensor (keeps fixtures small)."""
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].double().numpy()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs("/tmp/a2x_golden/debug", exist_ok=True)
    os.chdir("/tmp/a2x_golden")  # the reference writes debug/debug_image_bevfeat.png
    hypes = small_hypes()
    args = hypes["model"]["args"]
    model = ref_import.create_model(hypes)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = O.det_init_state_dict(shapes, seed=1234)
    full = model.state_dict()
    full.update(sd)
    # a strongly negative classification bias makes the communication mask non-trivial
    full["cls_head.bias"] = full["cls_head.bias"] - 5.0
    model.load_state_dict(full)
    sd = {k: v.clone() for k, v in model.state_dict().items()}

    agents = ["vehicle", "vehicle", "rsu", "drone"]
    dd = make_scene(hypes, agents, n_points=6000, seed=7, max_voxels=hypes["preprocess"]["args"]["max_voxel_train"])
    out = {"agents": np.array(agents), "n_points": 6000, "scene_seed": 7, "param_seed": 1234, "cls_bias_shift": -5.0,
           "range_xy": np.array(SMALL_RANGE_XY)}

    # ---------------- eval forward: reference vs oracle
    model.eval()
    with torch.no_grad():
        ref_out = model(dd)
    keep = {}
    with torch.no_grad():
        ora_out, _ = O.where2com_forward(sd, args, dd, training=False, keep=keep)
    for k in ("psm", "rm", "obj"):
        err = float((ref_out[k] - ora_out[k]).abs().max())
        print("eval %s: ref-vs-oracle max abs err %.3e (max |ref| %.3f)" % (k, err, float(ref_out[k].abs().max())))
        assert err < 1e-5, k
        out["eval_" + k] = ref_out[k].numpy()
    print("eval com", float(ref_out["com"]), float(ora_out["com"]), "comm_rate", ref_out["comm_rate"], ora_out["comm_rate"])
    assert abs(float(ref_out["com"]) - float(ora_out["com"])) < 1e-7 and ref_out["comm_rate"] == ora_out["comm_rate"]
    out["eval_com"] = float(ref_out["com"])
    out["eval_comm_rate"] = int(ref_out["comm_rate"])
    for k in ("spatial_features", "spatial_features_2d", "psm_single", "mask", "fused_l0", "fused_l1", "fused_l2",
              "fused_feature"):
        out["eval_keep_" + k] = sample(keep[k])
    for t in O.AGENT_TYPES:
        out["eval_keep_pillar_features_" + t] = sample(keep["pillar_features_" + t])

    # ---------------- train forward + loss + backward
    from opencood.loss.point_pillar_loss_multiclass import PointPillarLossMultiClass
    crit = PointPillarLossMultiClass(hypes["loss"]["det"]["args"])
    H, W = ref_out["psm"].shape[2:]
    labels = make_labels(11, 1, H, W, args["anchor_number"])
    out["label_seed"] = 11
    model.train()
    model.load_state_dict(sd)
    random.seed(5)
    tr_out = model(dd)
    loss = crit(tr_out, labels)
    model.zero_grad()
    loss.backward()
    ref_grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    ref_buffers = {k: v.clone() for k, v in model.state_dict().items() if "running" in k}

    sd_t = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
            for k, v in sd.items()}
    random.seed(5)
    o_out, o_buf = O.where2com_forward(sd_t, args, dd, training=True)
    o_loss, o_reg, o_cls, o_obj = O.point_pillar_loss_multiclass(o_out, labels, args["num_class"],
                                                                 hypes["loss"]["det"]["args"]["cls_weight"],
                                                                 hypes["loss"]["det"]["args"]["reg"])
    o_loss.backward()
    print("train loss ref %.8f oracle %.8f" % (float(loss), float(o_loss)))
    assert abs(float(loss) - float(o_loss)) < 1e-5 * max(1.0, abs(float(loss)))
    for k in ("psm", "rm", "obj"):
        err = float((tr_out[k] - o_out[k]).abs().max())
        print("train %s: max abs err %.3e" % (k, err))
        assert err < 1e-4, k
        out["train_" + k] = tr_out[k].detach().numpy()
    out["train_loss"] = float(loss)
    out["train_com"] = float(tr_out["com"])
    out["train_K_seed"] = 5
    worst = 0.0
    for k, g in ref_grads.items():
        og = sd_t[k].grad
        if og is None:
            assert float(g.abs().max()) == 0.0, "oracle has no grad for %s" % k
            continue
        denom = float(g.abs().max()) + 1e-12
        worst = max(worst, float((g - og).abs().max()) / denom)
        out["grad_" + k] = sample(g, 512)
        out["gradnorm_" + k] = float(g.double().norm())
    print("worst relative grad error ref-vs-oracle: %.3e over %d params" % (worst, len(ref_grads)))
    assert worst < 1e-3
    for k, v in ref_buffers.items():
        err = float((v - o_buf[k]).abs().max())
        assert err < 1e-5, (k, err)
        out["buf_" + k] = sample(v, 64)
    print("running stats after one train step match (triple update)")

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

    import json
    cfg = {"model_args": jsonable(args), "preprocess": jsonable(hypes["preprocess"]),
           "loss_args": jsonable(hypes["loss"]["det"]["args"]), "postprocess": jsonable(hypes["postprocess"]),
           "source": "opencood/hypes_yaml/" + YAML + " (lidar ranges shrunk to %s)" % (SMALL_RANGE_XY,)}
    json.dump(cfg, open(os.path.join(ROOT, "tests", "golden", "w2c_small_config.json"), "w"), indent=1)
    full = ref_import.load_hypes(YAML)
    cfg = {"model_args": jsonable(full["model"]["args"]), "preprocess": jsonable(full["preprocess"]),
           "loss_args": jsonable(full["loss"]["det"]["args"]), "postprocess": jsonable(full["postprocess"]),
           "source": "opencood/hypes_yaml/" + YAML + " as loaded by yaml_utils.load_yaml (grid_size etc. injected)"}
    os.makedirs(os.path.join(ROOT, "configs"), exist_ok=True)
    json.dump(cfg, open(os.path.join(ROOT, "configs", "airv2x_intermediate_where2com.json"), "w"), indent=1)

    dst = os.path.join(ROOT, "tests", "golden", "w2c_small.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, "%.1f KB" % (os.path.getsize(dst) / 1024))


if __name__ == "__main__":
    main()
