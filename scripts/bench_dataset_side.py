"""Worker-side cost of one training sample, reference vs this repo (container side: needs /root/reference; CPU only).

The REAL `IntermediateFusionDatasetAirv2x.__getitem__` + `collate_batch_train` (spconv replaced by the sequential C
restatement, which does the same per-point work) against `intermediate_fusion_dataset.IntermediateFusionDatasetAirv2x` of
this repo on the same BASELINE-config-2-sized synthetic scene (5 agents x 60 000 points, 20 boxes per agent's list):
milliseconds per sample on one core (what a DataLoader worker spends) and the bytes the collated batch ships to the GPU.

    python scripts/bench_dataset_side.py            # -> prints a markdown table (profiles/r2_dataset_side.md)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import a2x_import  # noqa: E402
import dataset_common as DC  # noqa: E402
import make_golden_dataset as MGD  # noqa: E402
from oracle import ref_import  # noqa: E402


def tensor_bytes(x):
    if torch.is_tensor(x):
        return x.numel() * x.element_size()
    if isinstance(x, np.ndarray):
        return x.nbytes
    if isinstance(x, dict):
        return sum(tensor_bytes(v) for v in x.values())
    if isinstance(x, (list, tuple)):
        return sum(tensor_bytes(v) for v in x)
    return 0


def main():
    torch.set_num_threads(1)
    os.makedirs("debug", exist_ok=True)
    IFD = MGD.reference_env()
    DS = a2x_import.pkg("intermediate_fusion_dataset")
    hypes = ref_import.load_hypes(MGD.YAML)
    scenes = [DC.synth_scene(DS, seed=100 + i, n_veh=2, n_rsu=2, n_drone=1, n_obj=40, n_pts=60000, far=False, cameras=False)
              for i in range(3)]
    # the reference cannot run without camera images (torch.stack of an empty list): give every agent one small frame
    from PIL import Image
    for sc in scenes:
        for rec in sc.values():
            rec["cameras"] = [Image.fromarray(np.zeros((72, 128, 3), np.uint8))]
            rec["params"]["delay_extrinsic"] = np.eye(4, dtype=np.float32)[None]
            rec["params"]["delay_intrinsic"] = np.eye(3, dtype=np.float32)[None]
    ref = MGD.reference_dataset(IFD, hypes, True)
    rows = []
    for name in ("reference", "this repo"):
        t_item, t_coll, nbytes = [], [], 0
        for rep in range(3):
            if name == "reference":
                np.random.seed(rep)
                t0 = time.perf_counter()
                items, batch = MGD.run_reference(ref, scenes[rep:rep + 1], seed=rep)
                t1 = time.perf_counter()
                t_item.append(t1 - t0)
                ego = batch["ego"]
                nbytes = tensor_bytes({k: ego[k] for k in ("vehicle", "rsu", "drone")}) + tensor_bytes(ego["label_dict"])
            else:
                lidar_only = [{cid: {k: v for k, v in rec.items() if k != "cameras"} for cid, rec in scenes[rep].items()}]
                for rec in lidar_only[0].values():
                    rec["cameras"] = []
                t0 = time.perf_counter()
                _, items, batch = MGD.run_ours(DS, hypes, True, lidar_only, seed=rep)
                t1 = time.perf_counter()
                t_item.append(t1 - t0)
                ego = batch["ego"]
                nbytes = tensor_bytes(ego["raw_points"]) + tensor_bytes({k: ego[k] for k in ("object_bbx_center", "object_bbx_mask",
                                                                                               "object_class_ids")})
        rows.append((name, 1e3 * float(np.median(t_item)), nbytes / 1e6))
    print("| dataset side, one sample (5 agents x 60k points, 1 core) | __getitem__ + collate, ms | lidar + label bytes to the GPU, MB |")
    print("|---|---|---|")
    for name, ms, mb in rows:
        print("| %s | %.0f | %.1f |" % (name, ms, mb))


if __name__ == "__main__":
    main()
