"""CPU: the N > 1 path (gradient averaging across data-parallel replicas) with world_size 2 over gloo."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import a2x_import

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = a2x_import.pkg("dist")
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
    params[2].requires_grad = False
    params[0].grad = torch.full((5, 3), float(rank + 1))
    params[1].grad = torch.arange(7, dtype=torch.float32) * (rank + 1)
    avg = D.GradAverager(params)
    assert params[0].grad.data_ptr() == avg.flat.data_ptr()          # p.grad are views of ONE flat buffer: no staging copies
    avg()
    avg()  # idempotent on already-averaged gradients
    # two buckets (the in-step protocol of W2CEngine.backward): start() the early one, finish() reduces the rest and joins
    q2 = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2))]
    b = D.GradAverager(q2, late=lambda n: n == "b", names=["a", "b", "c"])
    assert b.split == 6 and b.flat.numel() == 9 and q2[1].grad.data_ptr() == b.flat[6:].data_ptr()   # late params at the tail
    for i, p in enumerate(q2):
        p.grad.fill_(float((rank + 1) * (i + 1)))
    b.start()
    b.finish()
    assert all(torch.allclose(p.grad, torch.full_like(p.grad, 1.5 * (i + 1))) for i, p in enumerate(q2))
    q.put((rank, params[0].grad.clone(), params[1].grad.clone()))
    dist.destroy_process_group()


def test_grad_averager_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, g0, g1 in res:
        assert torch.allclose(g0, torch.full((5, 3), 1.5))
        assert torch.allclose(g1, torch.arange(7, dtype=torch.float32) * 1.5)


def _gather_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import a2x_import

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = a2x_import.pkg("dist")
    local = torch.full((1, 2, 3, 4), float(rank + 1))
    out = D.gather_agent_maps(local)
    q.put((rank, out.clone()))
    dist.destroy_process_group()


def test_agent_parallel_plan_and_gather_world2():
    """agents one per rank: the rank plan follows the reference's scene-major agent order and the gather
    returns the maps in agent order on every rank"""
    sys.path.insert(0, ROOT)
    import a2x_import

    D = a2x_import.pkg("dist")
    per_rank, glob = D.agent_rank_plan(["vehicle", "vehicle", "rsu", "drone"])
    assert glob["record_len"] == [4] and glob["counts"] == {"vehicle": 2, "rsu": 1, "drone": 1}
    assert per_rank[0]["vehicle"] == {"record_len": [1], "batch_idxs": [0]} and per_rank[0]["rsu"]["batch_idxs"] == []
    assert per_rank[2]["rsu"]["record_len"] == [1] and per_rank[3]["drone"]["batch_idxs"] == [0]
    try:
        D.agent_rank_plan(["rsu", "vehicle"])
    except AssertionError:
        pass
    else:
        raise AssertionError("out-of-order agents must be rejected (ego = first vehicle)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, out in res:
        assert out.shape == (2, 2, 3, 4)
        assert torch.equal(out[0], torch.full((2, 3, 4), 1.0)) and torch.equal(out[1], torch.full((2, 3, 4), 2.0))


def _dataset_worker(rank, world, port, q, tree):
    """scene-parallel input side: every rank builds the dataset on the same directory, a DistributedSampler hands each rank
    its shard, the ranks' batches are collated independently and the union covers every sample exactly once"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import json

    import a2x_import

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    DS = a2x_import.pkg("intermediate_fusion_dataset")
    hypes = json.load(open(os.path.join(ROOT, "tests", "golden", "dataset_config.json")))
    hypes.update(root_dir=tree, validate_dir=tree, task="det")
    ds = DS.IntermediateFusionDatasetAirv2x(hypes, False, train=False)
    sampler = torch.utils.data.distributed.DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=False)
    loader = torch.utils.data.DataLoader(ds, batch_size=1, sampler=sampler, collate_fn=ds.collate_batch_train, num_workers=0)
    seen, agents = [], 0
    for batch in loader:
        ego = batch["ego"]
        seen.append((ego["scenario_index_list"][0], ego["timestamp_key_list"][0]))
        agents += int(ego["record_len"].sum())
        assert ego["raw_points"]["offsets"][-1] == 600 * int(ego["record_len"].sum())
    total = torch.tensor([agents], dtype=torch.int64)
    dist.all_reduce(total)
    q.put((rank, seen, int(total)))
    dist.destroy_process_group()


def test_dataset_shards_over_two_ranks(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenes_common as SC

    tree = SC.write_tree(str(tmp_path / "tree"), seed=9, late_agent=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dataset_worker, args=(r, 2, port, q, tree)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, t0), (_, s1, t1) = res
    assert len(s0) == len(s1) == 3 and not set(s0) & set(s1)
    assert sorted(s0 + s1) == [(s, t) for s in (0, 1) for t in (0, 5, 10)]
    assert t0 == t1 == 6 * 6            # six agents in every one of the six samples, summed over the ranks
