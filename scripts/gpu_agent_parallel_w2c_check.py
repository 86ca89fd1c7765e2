"""Agent-parallel Where2comm inference (one agent per GPU; sparse level-0 payload by warp-ballot compaction) == single-GPU
Where2comm, for both transports. Run under torchrun with world_size = number of agents:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        scripts/gpu_agent_parallel_w2c_check.py [--full]
"""
import json
import os
import sys

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
    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    D = a2x_import.pkg("dist")
    cfg = json.load(open(os.path.join(ROOT, "configs", "airv2x_intermediate_where2com.json") if full else
                         os.path.join(ROOT, "tests", "golden", "w2c_small_config.json")))
    npts = 60000 if full else 6000
    types = sorted((["vehicle", "rsu", "vehicle", "rsu", "drone", "vehicle", "rsu", "drone"])[:world],
                   key=lambda t: {"vehicle": 0, "rsu": 1, "drone": 2}[t])
    model = M.Airv2xWhere2com(cfg["model_args"])
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = {k: v for k, v in model.state_dict().items()}
    sd.update(O.det_init_state_dict(shapes, seed=99))
    sd["cls_head.bias"] = sd["cls_head.bias"] - 4.6          # makes the communication mask non-trivial
    model.load_state_dict(sd)
    model.to(dev).eval()
    rng = cfg["preprocess"]["cav_lidar_range"]
    clouds = [O.synth_points(700 + k, npts, rng, (35.0, 15.0) if full else (10.0, 5.0)) for k in range(world)]
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
        single = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in single.items()}
        res = {"world": world, "full": full, "ms_single_gpu": ms_single, "com_single": float(single["com"]),
               "comm_rate_single": single["comm_rate"]}
        for transport in ("nccl", "peer"):
            try:
                ap = D.AgentParallelWhere2comm(model, types, transport=transport)
                out, ms = timed(lambda: ap(torch.from_numpy(clouds[rank]), cfg["preprocess"]))
                err = max(float((out[k] - single[k]).abs().max()) for k in ("psm", "rm", "obj"))
                e = torch.tensor([err], device=dev)
                dist.all_reduce(e, op=dist.ReduceOp.MAX)
                sent = int(ap._state[1][:64].view(torch.int32)[0])
                res[transport] = {"ms": ms, "max_abs_err_vs_single_gpu": float(e), "com": float(out["com"]),
                                  "comm_rate": out["comm_rate"], "level0_cells_sent_rank%d" % rank: sent}
            except Exception as ex:
                import traceback

                traceback.print_exc()
                res[transport] = {"error": repr(ex)[:300]}
    if rank == 0:
        print("AGENT_PARALLEL_W2C " + json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
