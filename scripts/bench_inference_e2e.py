"""Serving-path latency of BASELINE config 2 on one GPU: pinned host clouds (5 agents x 60k points, sensor data as it
arrives) -> H2D -> voxelise + Where2comm eval forward -> GPU decode + rotated NMS -> D2H of the kept boxes. One JSON line.
Not the headline bench (bench.py measures the training step).

    python scripts/bench_inference_e2e.py [n_agents] [--iters K]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import a2x_import
import bench


def main():
    n_agents = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 5
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 30
    cfg = bench.load_config()
    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    PP = a2x_import.pkg("postprocess")
    torch.manual_seed(0)
    model = M.Airv2xWhere2com(cfg["model_args"]).cuda().eval()
    post = PP.DetPostprocessor(cfg["postprocess"], "cuda")
    types = ["vehicle", "vehicle", "rsu", "rsu", "drone"][:n_agents]
    rng = cfg["preprocess"]["cav_lidar_range"]
    clouds = [bench.synth_cloud(k, bench.N_POINTS, rng) for k in range(n_agents)]
    offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    host_pts = torch.from_numpy(np.concatenate(clouds, 0)).pin_memory()
    host_offs = torch.from_numpy(offs).pin_memory()
    dd = {"raw_points": {"points": host_pts, "offsets": host_offs, "preprocess": cfg["preprocess"], "filter": True}}
    for t in ("vehicle", "rsu", "drone"):
        n = sum(1 for a in types if a == t)
        dd[t] = {"record_len": [n], "batch_idxs": [0] if n else []}

    def once():
        with torch.no_grad():
            out = model(dd)                                   # H2D of the clouds happens inside (pinned -> device)
            boxes, scores, labels, _ = post(out)              # one D2H read (count), results stay on the device
            return None if boxes is None else (boxes.cpu(), scores.cpu(), labels.cpu())

    # random-init heads: shift the objectness bias so that ~500 anchors pass the 0.2 threshold (a realistic NMS load)
    with torch.no_grad():
        obj = model(dd)["obj"].flatten()
        kth = torch.topk(obj, 500).values[-1]
        model.obj_head.bias += float(np.log(0.2 / 0.8)) - kth
    for _ in range(5):
        res = once()
    torch.cuda.synchronize()
    lat = []
    for _ in range(iters):
        t0 = time.perf_counter()
        res = once()
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = np.array(lat)
    print(json.dumps({"metric": "scene latency, points -> NMS'd boxes (Where2comm eval, %d agents x 60k pts)" % n_agents,
                      "median_ms": float(np.median(lat)), "p90_ms": float(np.percentile(lat, 90)), "scenes_per_s": 1000.0 / float(np.median(lat)),
                      "iters": iters, "boxes_kept": 0 if res is None else int(res[0].shape[0]),
                      "h2d_bytes": int(host_pts.numel() * 4 + host_offs.numel() * 4),
                      "note": "wall clock per scene incl. H2D of the raw clouds, voxelisation, forward, decode + rotated NMS, D2H of the boxes"}))


if __name__ == "__main__":
    main()
