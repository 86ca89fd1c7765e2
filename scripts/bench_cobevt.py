"""Eval-forward timing of the CoBEVT path (BASELINE config 4 geometry, N agents on one GPU) with a per-C-ABI-call
breakdown (CUDA events). Prints one JSON line. Not the headline bench (bench.py measures config 2)."""
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
    n_agents = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    cfg = json.load(open(os.path.join(ROOT, "configs", "airv2x_intermediate_cobevt.json")))
    M = a2x_import.pkg("opencood.models.airv2x_cobevt")
    libmod = a2x_import.pkg("_lib")
    torch.manual_seed(0)
    model = M.Airv2xCoBEVT(cfg["model_args"]).cuda().eval()
    types = (["vehicle"] * 3 + ["rsu"] * 2 + ["drone"] * 2)
    types = sorted(types[:2] + types[3:4] + types[5:6] + types[2:3] + types[4:5] + types[6:7][:0], key=lambda t: {"vehicle": 0, "rsu": 1, "drone": 2}[t])[:n_agents] \
        if n_agents != 5 else ["vehicle", "vehicle", "rsu", "rsu", "drone"]
    rng = cfg["preprocess"]["cav_lidar_range"]
    clouds = [bench.synth_cloud(k, bench.N_POINTS, rng) for k in range(len(types))]
    offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    dd = {"raw_points": {"points": torch.from_numpy(np.concatenate(clouds, 0)).cuda(), "offsets": torch.from_numpy(offs).cuda(),
                         "preprocess": cfg["preprocess"], "filter": True}}
    for t in ("vehicle", "rsu", "drone"):
        n = sum(1 for a in types if a == t)
        dd[t] = {"record_len": [n], "batch_idxs": [0] if n else []}
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
    groups = {}
    for name, _, a, b in prof:
        g = groups.setdefault(name, [0.0, 0])
        g[0] += a.elapsed_time(b)
        g[1] += 1
    top = sorted(((k, round(v[0], 3), v[1]) for k, v in groups.items()), key=lambda x: -x[1])[:10]
    print(json.dumps({"metric": "scenes/sec (eval fwd) CoBEVT %d-agent 60k-pt, L=7" % len(types), "value": 1000.0 / ms,
                      "ms_per_scene": ms, "agents": types, "top_calls_ms": top}))


if __name__ == "__main__":
    main()
