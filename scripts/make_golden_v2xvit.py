"""Generate tests/golden/v2xvit_small.npz by running the REAL reference Airv2xV2XVit (imported from /root/reference,
CPU, eval mode, L = 15 padded as shipped) and checking oracle/v2xvit_oracle.py against it. Container-side only.

    python scripts/make_golden_v2xvit.py
"""
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from oracle import ref_import, v2xvit_oracle as VO, w2c_oracle as O  # noqa: E402
import make_golden as MG  # noqa: E402

YAML = "airv2x/lidar/det/airv2x_intermediate_v2xvit.yaml"


def scene_extras(agents, L):
    """prior_encoding [v/30, time_delay, infra] and spatial_correction_matrix (one neighbour mis-aligned by
    0.2 rad / (6, -3) m, another by a pure translation) — SURVEY App. A-5"""
    prior = torch.zeros(1, L, 3)
    scm = torch.eye(4, dtype=torch.float64).repeat(1, L, 1, 1)
    for i, t in enumerate(agents):
        prior[0, i] = torch.tensor([0.1 * i, float(i % 3), 1.0 if t == "rsu" else 0.0])
    a = 0.2
    scm[0, 1, :2, :2] = torch.tensor([[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]], dtype=torch.float64)
    scm[0, 1, 0, 3], scm[0, 1, 1, 3] = 6.0, -3.0
    scm[0, 2, 0, 3], scm[0, 2, 1, 3] = -4.8, 1.6
    return prior, scm


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs("/tmp/a2x_golden/debug", exist_ok=True)
    os.chdir("/tmp/a2x_golden")
    MG.YAML = YAML
    hypes = MG.small_hypes()
    args = hypes["model"]["args"]
    model = ref_import.create_model(hypes)
    print("params", sum(p.numel() for p in model.parameters()))
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = O.det_init_state_dict(shapes, seed=2468)
    # keep the analytic sinusoid table of the RTE embedding (v2xvit_basic.py:47-53); pos_embedding / relation tensors
    # take the seeded values (2-D / 4-D "weights" in det_init_state_dict's rule)
    sd = {k: v for k, v in sd.items() if not k.endswith("rte.emb.emb.weight")}
    full = model.state_dict()
    full.update(sd)
    model.load_state_dict(full)
    sd = {k: v.clone() for k, v in model.state_dict().items()}

    agents = ["vehicle", "vehicle", "rsu", "drone"]
    L = sum(args["max_cav"].values())
    dd = O.make_scene(hypes["preprocess"], agents, 6000, 13, hypes["preprocess"]["args"]["max_voxel_train"])
    dd["prior_encoding"], dd["spatial_correction_matrix"] = scene_extras(agents, L)
    out = {"agents": np.array(agents), "n_points": 6000, "scene_seed": 13, "param_seed": 2468,
           "range_xy": np.array(MG.SMALL_RANGE_XY), "max_cav_num": L}
    model.eval()
    with torch.no_grad():
        ref_out = model(dd)
        keep = {}
        ora_out, _ = VO.v2xvit_forward(sd, args, dd, training=False, keep=keep)
    for k in ("psm", "rm", "obj"):
        err = float((ref_out[k] - ora_out[k]).abs().max())
        print("eval %s: ref-vs-oracle max abs err %.3e (max |ref| %.3f)" % (k, err, float(ref_out[k].abs().max())))
        assert err < 2e-5, k
        out["eval_" + k] = ref_out[k].numpy()
    assert int(ref_out["comm_rate"]) == ora_out["comm_rate"]
    out["eval_comm_rate"] = int(ref_out["comm_rate"])
    for k in ("sttf", "com_mask", "layer0", "layer1", "layer2", "fused_feature"):
        out["eval_keep_" + k] = MG.sample(keep[k].float())
    out["rte_table"] = sd["fusion_net.encoder.rte.emb.emb.weight"].numpy()

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

    cfg = {"model_args": jsonable(args), "preprocess": jsonable(hypes["preprocess"]),
           "loss_args": jsonable(hypes["loss"]["det"]["args"]), "postprocess": jsonable(hypes["postprocess"]),
           "source": "opencood/hypes_yaml/" + YAML + " (lidar ranges shrunk to %s)" % (MG.SMALL_RANGE_XY,)}
    json.dump(cfg, open(os.path.join(ROOT, "tests", "golden", "v2xvit_small_config.json"), "w"), indent=1)
    full_h = ref_import.load_hypes(YAML)
    cfg = {"model_args": jsonable(full_h["model"]["args"]), "preprocess": jsonable(full_h["preprocess"]),
           "loss_args": jsonable(full_h["loss"]["det"]["args"]), "postprocess": jsonable(full_h["postprocess"]),
           "source": "opencood/hypes_yaml/" + YAML + " as loaded by yaml_utils.load_yaml"}
    json.dump(cfg, open(os.path.join(ROOT, "configs", "airv2x_intermediate_v2xvit.json"), "w"), indent=1)
    dst = os.path.join(ROOT, "tests", "golden", "v2xvit_small.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, "%.1f KB" % (os.path.getsize(dst) / 1024))


if __name__ == "__main__":
    main()
