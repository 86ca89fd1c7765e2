"""Eval-forward timing of the transformer-fusion paths at full geometry on one GPU — BASELINE config 3 (V2X-ViT) and
config 4 (CoBEVT), 5 agents x 60k points — with a per-C-ABI-call breakdown (CUDA events). One JSON line per model.
Not the headline bench (bench.py measures config 2).

    python scripts/bench_fusion_models.py [cobevt|v2xvit] [n_agents]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import a2x_import
import bench


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cobevt"
    n_agents = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    cfg = json.load(open(os.path.join(ROOT, "configs", "airv2x_intermediate_%s.json" % which)))
    M = a2x_import.pkg("opencood.models.airv2x_" + which)
    libmod = a2x_import.pkg("_lib")
    torch.manual_seed(0)
    model = (M.Airv2xCoBEVT if which == "cobevt" else M.Airv2xV2XVit)(cfg["model_args"]).cuda().eval()
    types = ["vehicle", "vehicle", "rsu", "rsu", "drone"][:n_agents] if n_agents <= 5 else \
        ["vehicle"] * 3 + ["rsu"] * 2 + ["drone"] * (n_agents - 5)
    rng = cfg["preprocess"]["cav_lidar_range"]
    clouds = [bench.synth_cloud(k, bench.N_POINTS, rng) for k in range(len(types))]
    offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    dd = {"raw_points": {"points": torch.from_numpy(np.concatenate(clouds, 0)).cuda(), "offsets": torch.from_numpy(offs).cuda(),
                         "preprocess": cfg["preprocess"], "filter": True}}
    for t in ("vehicle", "rsu", "drone"):
        n = sum(1 for a in types if a == t)
        dd[t] = {"record_len": [n], "batch_idxs": [0] if n else []}
    if which == "v2xvit":
        L = sum(cfg["model_args"]["max_cav"].values())
        prior = torch.zeros(1, L, 3)
        scm = torch.eye(4, dtype=torch.float64).repeat(1, L, 1, 1)
        for i, t in enumerate(types):
            prior[0, i] = torch.tensor([0.1 * i, float(i % 2), 1.0 if t == "rsu" else 0.0])
        scm[0, 1, 0, 3], scm[0, 1, 1, 3] = 6.0, -3.0
        scm[0, 1, :2, :2] = torch.tensor([[0.9801, -0.1987], [0.1987, 0.9801]], dtype=torch.float64)
        dd["prior_encoding"], dd["spatial_correction_matrix"] = prior, scm
    with torch.no_grad():
        for _ in range(3):
            model(dd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            model(dd)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        libmod.PROFILE = []
        model(dd)
        torch.cuda.synchronize()
        prof, libmod.PROFILE = libmod.PROFILE, None
    train = None
    if "--train" in sys.argv:
        lab_np = bench.synth_labels(3, 100, 352, cfg["model_args"]["anchor_number"])
        lab = {k: torch.from_numpy(v).cuda() for k, v in lab_np.items()}
        model.train()
        for _ in range(2):
            model.train_step(dd, lab, 1.0, 2.0)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            loss3 = model.train_step(dd, lab, 1.0, 2.0)
        e1.record()
        torch.cuda.synchronize()
        train = {"ms_per_step": e0.elapsed_time(e1) / 5, "loss": float(loss3.sum()),
                 "note": "fwd (train-mode BN, nn.Dropout on as in the shipped yaml) + PointPillarLossMultiClass + bwd, eager launches"}
        libmod.PROFILE = []
        model.train_step(dd, lab, 1.0, 2.0)
        torch.cuda.synchronize()
        prof, libmod.PROFILE = libmod.PROFILE, None
    groups = {}
    for name, _, a, b in prof:
        g = groups.setdefault(name, [0.0, 0])
        g[0] += a.elapsed_time(b)
        g[1] += 1
    top = sorted(((k, round(v[0], 3), v[1]) for k, v in groups.items()), key=lambda x: -x[1])[:24]
    print(json.dumps({"metric": "scenes/sec (eval fwd) %s %d-agent 60k-pt" % (which, len(types)), "value": 1000.0 / ms,
                      "ms_per_scene": ms, "agents": types, "train_step": train, "top_calls_ms": top}))


if __name__ == "__main__":
    main()
