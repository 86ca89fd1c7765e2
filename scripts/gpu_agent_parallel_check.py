"""Agent-parallel CoBEVT (one agent per GPU) == single-GPU CoBEVT, for both transports (NCCL all-gather; peer-memory
pull fused into the regroup kernel). Run under torchrun with world_size = number of agents:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/gpu_agent_parallel_check.py [--full]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import a2x_import
from oracle import w2c_oracle as O  # synthetic clouds + seeded parameters only (test infrastructure)


def main():
    full = "--full" in sys.argv
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    M = a2x_import.pkg("opencood.models.airv2x_cobevt")
    D = a2x_import.pkg("dist")
    if full:
        cfg = json.load(open(os.path.join(ROOT, "configs", "airv2x_intermediate_cobevt.json")))
        npts = 60000
    else:
        cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "cobevt_small_config.json")))
        npts = 6000
    args = cfg["model_args"]
    types = (["vehicle", "vehicle", "vehicle", "rsu", "rsu", "drone", "drone"] + ["drone"])[:world] if world > 2 else ["vehicle", "rsu"]
    types = sorted(types, key=lambda t: {"vehicle": 0, "rsu": 1, "drone": 2}[t])
    if world > sum(args["max_cav"].values()):
        args["max_cav"] = {"vehicle": 3, "rsu": 3, "drone": 2}          # the benchmark overrides max_cav
    model = M.Airv2xCoBEVT(args)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if "relative_position_index" not in k}
    sd = {k: v for k, v in model.state_dict().items()}
    sd.update(O.det_init_state_dict(shapes, seed=77))
    model.load_state_dict(sd)
    model.to(dev).eval()
    rng = cfg["preprocess"]["cav_lidar_range"]
    clouds = [O.synth_points(500 + k, npts, rng, (35.0, 15.0) if full else (10.0, 5.0)) for k in range(world)]
    offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    raw = {"raw_points": {"points": torch.from_numpy(np.concatenate(clouds, 0)), "offsets": torch.from_numpy(offs),
                          "preprocess": cfg["preprocess"], "filter": True}}
    for t in ("vehicle", "rsu", "drone"):
        n = sum(1 for a in types if a == t)
        raw[t] = {"record_len": [n], "batch_idxs": [0] if n else []}

    def timed(fn, n=5):
        for _ in range(2):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t)

    with torch.no_grad():
        single, ms_single = timed(lambda: model(raw))
        single = {k: v.clone() for k, v in single.items()}
        res = {"world": world, "full": full, "ms_single_gpu": ms_single}
        for transport in ("nccl", "peer"):
            try:
                ap = D.AgentParallelCoBEVT(model, types, transport=transport)
                out, ms = timed(lambda: ap(torch.from_numpy(clouds[rank]), cfg["preprocess"]))
                err = max(float((out[k] - single[k]).abs().max()) for k in ("psm", "rm", "obj"))
                exact = all(torch.equal(out[k], single[k]) for k in ("psm", "rm", "obj"))
                e = torch.tensor([err], device=dev)
                dist.all_reduce(e, op=dist.ReduceOp.MAX)
                res[transport] = {"ms": ms, "max_abs_err_vs_single_gpu": float(e), "bit_exact_rank%d" % rank: exact}
            except Exception as ex:  # report, do not hide
                import traceback

                traceback.print_exc()
                res[transport] = {"error": repr(ex)[:300]}
    if rank == 0:
        print("AGENT_PARALLEL " + json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
