"""BASELINE.md bar (b): the reference's algorithm through STOCK PyTorch on one B200 — the torch-functional restatement of
the reference modules (oracle/w2c_oracle.py: F.conv2d / F.batch_norm / F.conv_transpose2d / bmm ..., i.e. the cuDNN /
cuBLAS kernels the reference's nn.Modules dispatch to) moved to the GPU, train-mode forward + PointPillarLossMultiClass +
autograd backward on BASELINE config 2 (5 agents x 60k points), same synthetic scene and labels as bench.py. Voxelisation
(CPU in the reference) is excluded from this timing. Comparison bar only: nothing here is the product path.

    python scripts/bench_stock_pytorch_gpu.py [--tf32] [--steps K]
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import a2x_import
import bench
from oracle import voxelize as V, w2c_oracle as O


def to_dev(d, dev):
    if isinstance(d, dict):
        return {k: to_dev(v, dev) for k, v in d.items()}
    return d.to(dev) if torch.is_tensor(d) else d


def main():
    tf32 = "--tf32" in sys.argv
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 5
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    cfg = bench.load_config()
    args, pre = cfg["model_args"], cfg["preprocess"]
    rng = pre["cav_lidar_range"]
    types = ["vehicle", "vehicle", "rsu", "rsu", "drone"]
    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    torch.manual_seed(0)
    model = M.Airv2xWhere2com(args)                      # parameter container: reference-shaped state_dict
    sd = {k: v.detach().clone().cuda() for k, v in model.state_dict().items()}
    p = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k and "gaussian" not in k else v)
         for k, v in sd.items()}
    dd, k = {}, 0
    for t in O.AGENT_TYPES:
        ids = [i for i, a in enumerate(types) if a == t]
        per = []
        for i in ids:
            pts = V.mask_points(bench.synth_cloud(i, bench.N_POINTS, rng), rng, ego_box=(i == 0))
            per.append(V.voxelize(pts, rng, pre["args"]["voxel_size"], pre["args"]["max_points_per_voxel"], pre["args"]["max_voxel_train"]))
        dd[t] = {"batch_merged_lidar_features_torch": {k2: torch.from_numpy(v) for k2, v in V.collate(per).items()},
                 "record_len": torch.tensor([len(ids)], dtype=torch.int32), "batch_idxs": [0]}
    dd["record_len"] = torch.tensor([len(types)], dtype=torch.int32)
    dd = to_dev(dd, "cuda")
    dd["record_len"] = dd["record_len"].cpu()
    for t in O.AGENT_TYPES:
        dd[t]["record_len"] = dd[t]["record_len"].cpu()
    lab = {k2: torch.from_numpy(v).cuda() for k2, v in bench.synth_labels(3, 100, 352, args["anchor_number"]).items()}
    lab = {"targets": lab["targets"].double(), "pos_equal_one": lab["pos_equal_one"].double(), "class_ids": lab["class_ids"].long(),
           "neg_equal_one": 1.0 - lab["pos_equal_one"].double()}

    def step():
        for v in p.values():
            if torch.is_tensor(v) and v.grad is not None:
                v.grad = None
        out, _ = O.where2com_forward(p, args, dd, training=True)
        loss = O.point_pillar_loss_multiclass(out, lab, args["num_class"], 1.0, 2.0)[0]
        loss.backward()
        return loss

    random.seed(0)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"metric": "scenes/sec (fwd+bwd) Where2Comm 5-agent 60k-pt", "impl": "stock PyTorch on B200 (cuDNN/cuBLAS, %s)"
                      % ("TF32 allowed" if tf32 else "fp32"), "value": 1000.0 / ms, "ms_per_step": ms, "steps": steps,
                      "loss": float(loss.detach()), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                      "note": "voxelisation excluded (CPU in the reference); eager launches, cudnn.benchmark on"}))


if __name__ == "__main__":
    main()
