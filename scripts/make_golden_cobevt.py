"""Generate tests/golden/cobevt_small.npz by running the REAL reference Airv2xCoBEVT (imported from /root/reference,
CPU, eval mode) and checking oracle/cobevt_oracle.py against it. Container-side only.

    python scripts/make_golden_cobevt.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from oracle import cobevt_oracle as CO, ref_import, w2c_oracle as O  # noqa: E402
import make_golden as MG  # noqa: E402

YAML = "airv2x/lidar/det/airv2x_intermediate_cobevt.yaml"


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs("/tmp/a2x_golden/debug", exist_ok=True)
    os.chdir("/tmp/a2x_golden")
    MG.YAML = YAML
    hypes = MG.small_hypes()
    args = hypes["model"]["args"]
    model = ref_import.create_model(hypes)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if "relative_position_index" not in k}
    sd = O.det_init_state_dict(shapes, seed=4321)
    full = model.state_dict()
    full.update(sd)
    model.load_state_dict(full)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    print("params", sum(p.numel() for p in model.parameters()))

    agents = ["vehicle", "vehicle", "rsu", "drone"]           # 4 real agents, padded to L = 3 + 2 + 2 = 7
    dd = O.make_scene(hypes["preprocess"], agents, 6000, 9, hypes["preprocess"]["args"]["max_voxel_train"])
    out = {"agents": np.array(agents), "n_points": 6000, "scene_seed": 9, "param_seed": 4321,
           "range_xy": np.array(MG.SMALL_RANGE_XY)}
    model.eval()
    with torch.no_grad():
        ref_out = model(dd)
        keep = {}
        ora_out, _ = CO.cobevt_forward(sd, args, dd, training=False, keep=keep)
    for k in ("psm", "rm", "obj"):
        err = float((ref_out[k] - ora_out[k]).abs().max())
        print("eval %s: ref-vs-oracle max abs err %.3e (max |ref| %.3f)" % (k, err, float(ref_out[k].abs().max())))
        assert err < 1e-5, k
        out["eval_" + k] = ref_out[k].numpy()
    for k in ("regroup", "block0", "block1", "block2", "fused_feature"):
        out["eval_keep_" + k] = MG.sample(keep[k])
    # NaiveCompressor (a10): same scene with `compression: 2`, the compressor's parameters seeded separately
    import copy
    hy2 = copy.deepcopy(hypes)
    hy2["model"]["args"]["compression"] = 2
    model2 = ref_import.create_model(hy2)
    sh2 = {k: tuple(v.shape) for k, v in model2.state_dict().items() if k.startswith("naive_compressor")}
    sd2 = dict(sd)
    sd2.update(O.det_init_state_dict(sh2, seed=97))
    full2 = model2.state_dict()
    full2.update(sd2)
    model2.load_state_dict(full2)
    sd2 = {k: v.clone() for k, v in model2.state_dict().items()}
    model2.eval()
    with torch.no_grad():
        r2 = model2(dd)
        o2, _ = CO.cobevt_forward(sd2, hy2["model"]["args"], dd, training=False)
    for k in ("psm", "rm", "obj"):
        err = float((r2[k] - o2[k]).abs().max())
        print("compression=2 eval %s: ref-vs-oracle max abs err %.3e" % (k, err))
        assert err < 1e-5, k
        out["cmp2_eval_" + k] = r2[k].numpy()
    out["cmp_param_seed"] = 97
    # the fusion network alone on a seeded random input with a ragged mask (2 scenes: 3 and 5 agents of L = 7)
    fa = dict(args["fax_fusion"])
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 7, 256, 8, 16, generator=g)
    mask = torch.tensor([[1, 1, 1, 0, 0, 0, 0], [1, 1, 1, 1, 1, 0, 0]])
    x = x * mask[:, :, None, None, None]
    com = mask[:, None, None, None, :].expand(-1, 8, 16, 1, -1)
    with torch.no_grad():
        r = model.fusion_net(x, com)
        o = CO.swap_fusion_encoder(sd, fa, x, com)
    err = float((r - o).abs().max())
    print("fusion_net alone: ref-vs-oracle %.3e" % err)
    assert err < 1e-5
    out["fusion_seed"] = 3
    out["fusion_out"] = r.numpy()

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
    json.dump(cfg, open(os.path.join(ROOT, "tests", "golden", "cobevt_small_config.json"), "w"), indent=1)
    full_h = ref_import.load_hypes(YAML)
    cfg = {"model_args": jsonable(full_h["model"]["args"]), "preprocess": jsonable(full_h["preprocess"]),
           "loss_args": jsonable(full_h["loss"]["det"]["args"]), "postprocess": jsonable(full_h["postprocess"]),
           "source": "opencood/hypes_yaml/" + YAML + " as loaded by yaml_utils.load_yaml"}
    json.dump(cfg, open(os.path.join(ROOT, "configs", "airv2x_intermediate_cobevt.json"), "w"), indent=1)
    dst = os.path.join(ROOT, "tests", "golden", "cobevt_small.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, "%.1f KB" % (os.path.getsize(dst) / 1024))


if __name__ == "__main__":
    main()
