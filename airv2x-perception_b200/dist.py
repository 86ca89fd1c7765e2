"""Multi-GPU plumbing: scenes shard data-parallel (one process per GPU, the reference's DDP —
opencood/tools/train.py:161-163, tools/multi_gpu_utils.py:38-48); the only exchange is the gradient average."""
import torch
import torch.distributed as dist


class GradAverager:
    """Gradient average for scene-parallel training (the reference's DDP, tools/train.py:161-163) without staging copies:
    every `p.grad` is a VIEW of one flat fp32 buffer (the fused step overwrites the views in place), so the exchange is one
    NCCL all-reduce (op = AVG) per bucket on the buffer itself.

    Two buckets, ordered by when the fused backward finalises them (see W2CEngine.backward): bucket 0 = everything except
    the level-0 block and the PillarVFE encoders (97 % of the bytes) is complete while the level-0 backward still runs, so
    `start(0)` launches its all-reduce asynchronously there (ProcessGroupNCCL's own stream; inside a CUDA-graph capture it
    becomes a forked branch of the graph) and `finish()` reduces the small rest and joins. `late(name)` decides the split;
    with no predicate there is one bucket and `__call__()` = start + finish after the step."""

    def __init__(self, params, late=None, names=None):
        params = list(params)
        names = list(names) if names is not None else [None] * len(params)
        pairs = [(n, p) for n, p in zip(names, params) if p.requires_grad]
        early = [(n, p) for n, p in pairs if late is None or not late(n)]
        tail = [(n, p) for n, p in pairs if late is not None and late(n)]
        self.params = [p for _, p in early + tail]
        self.flat = None
        self.split = sum(p.numel() for _, p in early)
        self._work = None
        if self.params:
            self._attach()

    def _attach(self):
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        pos = 0
        for p in self.params:
            view = self.flat[pos:pos + p.numel()].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            pos += p.numel()

    @staticmethod
    def active():
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    @property
    def nbytes(self):
        return 0 if self.flat is None else self.flat.numel() * self.flat.element_size()

    def _check_views(self):
        """an optimizer's zero_grad(set_to_none=True) drops the views: re-attach (their content is rewritten by the step)"""
        if self.flat is None or any(p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or
                                    p.grad.data_ptr() >= self.flat.data_ptr() + self.nbytes for p in self.params[:1] + self.params[-1:]):
            self._attach()

    def _reduce(self, t, async_op):
        if dist.get_backend() == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
        w = dist.all_reduce(t, async_op=False)      # gloo (CPU tests) has no AVG
        t.div_(dist.get_world_size())
        return w

    def start(self, bucket=0):
        """launch the all-reduce of the early bucket (asynchronous w.r.t. the current stream)"""
        if not self.active() or self.split == 0:
            return
        self._work = self._reduce(self.flat[:self.split], True)

    def finish(self):
        if not self.active():
            return
        if self.split < self.flat.numel():
            self._reduce(self.flat[self.split:], False)
        if self._work is not None:
            self._work.wait()          # the current stream waits for the early bucket (no host block)
        self._work = None

    def __call__(self):
        if not self.active():
            return
        self._check_views()
        if self._work is None:
            self.start()
        self.finish()


# ---------------------------------------------------------------------------------------------------------------------
# Agent-parallel inference (BASELINE config 4): one agent per GPU. Everything up to the fusion boundary
# (voxelise, PillarVFE, scatter, backbone, shrink) is per-agent independent; the path's only exchange is the gather of
# the shrunk BEV maps. Two transports:
#   "nccl": one all_gather_into_tensor of the fp32 maps, then the ordinary regroup;
#   "peer": the maps live in symmetric memory and every rank's regroup kernel pulls each peer's map over NVLink
#           directly into its slot of the padded token tensor (collective fused into the first consumer).
# Both are exact: the fused output equals the single-GPU output bit for bit (same kernels, same operands).
def agent_rank_plan(agent_types):
    """agent_types: the scene's agents in the reference's scene-major order (vehicles, RSUs, drones; ego first —
    common_modules/airv2x_base_model.py:212-236). Rank r owns agent r. Returns per-rank local dict skeletons and the
    global layout lists (record_len, per-type counts)."""
    order = {"vehicle": 0, "rsu": 1, "drone": 2}
    assert list(agent_types) == sorted(agent_types, key=lambda t: order[t]), \
        "agents must be ordered vehicles, RSUs, drones (ego = agent 0)"
    per_rank = []
    for r, t in enumerate(agent_types):
        per_rank.append({ty: {"record_len": [1 if ty == t else 0], "batch_idxs": [0] if ty == t else []}
                         for ty in order})
    return per_rank, {"record_len": [len(agent_types)],
                      "counts": {ty: sum(1 for t in agent_types if t == ty) for ty in order}}


def gather_agent_maps(local, group=None):
    """all-gather of one [1, ...] map per rank into [world, ...] (rank order = agent order); gloo or nccl"""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:
        dist.all_gather(list(out.split(1)), local.contiguous(), group=group)
    return out


class AgentParallelCoBEVT:
    """Airv2xCoBEVT with the agents of ONE scene sharded one per rank. Every rank returns the fused output."""

    def __init__(self, model, agent_types, transport="nccl", group=None):
        assert transport in ("nccl", "peer")
        self.model, self.group, self.transport = model, group, transport
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert len(agent_types) == self.world, "one agent per rank"
        self.agent_types = list(agent_types)
        self.per_rank, self.glob = agent_rank_plan(agent_types)
        self._symm = None

    def _layouts(self, dev):
        m = self.model
        skeleton = dict(self.per_rank[self.rank])
        skeleton["raw_points"] = True  # layout key only
        lay = dict(m._layout(skeleton, dev))
        # only agent 0 of the scene is the ego (mask_ego_points applies to it alone)
        lay["ego_flags"] = torch.tensor([1 if self.rank == 0 else 0], dtype=torch.uint8, device=dev)
        L, n = m.max_cav_num, self.world
        assert n <= L, "more agents than max_cav allows"
        glob = {"record_len": [n], "n_total": n,
                "scene_start": torch.zeros(1, dtype=torch.int32, device=dev),
                "scene_len": torch.tensor([n], dtype=torch.int32, device=dev),
                "key_mask": torch.tensor([[1] * n + [0] * (L - n)], dtype=torch.int32, device=dev)}
        return lay, glob

    def __call__(self, points, preprocess):
        """points: [P, 4] f32 cloud of THIS rank's agent (host or device). Returns {"psm","rm","obj"}."""
        m = self.model
        assert not m.training, "agent-parallel mode is inference only"
        dev = next(m.parameters()).device
        if not hasattr(self, "_lay"):
            self._lay = self._layouts(dev)
        lay, glob = self._lay
        dd = dict(self.per_rank[self.rank])
        pts = points.to(device=dev, dtype=torch.float32)
        dd["raw_points"] = {"points": pts, "offsets": torch.tensor([0, pts.shape[0]], dtype=torch.int32),
                            "preprocess": preprocess, "filter": True}
        lidar = m._lidar(dd, dev, lay)
        lidar["raw"]["ego_flags"] = lay["ego_flags"]
        eng, P = m.engine, m._param_dict()
        eng._begin_step()
        W = eng._pack_weights(P)
        if self.transport == "nccl":
            local = eng.encode(P, W, lidar, lay)
            feat = gather_agent_maps(local, self.group)
            heads = eng.fuse_heads(P, W, feat, glob)
        else:
            import torch.distributed._symmetric_memory as symm_mem

            if self._symm is None:
                la = m.args[self.agent_types[self.rank]]["lidar"]["point_pillar_scatter"]["grid_size"]
                shape = (1, int(la[1]) // 2, int(la[0]) // 2, eng.c_shrink)
                buf = symm_mem.empty(shape, dtype=torch.float32, device=dev)
                hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
                table = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=dev)
                self._symm = (buf, hdl, table)
            buf, hdl, table = self._symm
            eng.encode(P, W, lidar, lay, out=buf)
            hdl.barrier(channel=0)   # every rank's map is complete before anyone pulls it
            heads = eng.fuse_heads(P, W, buf, glob, peer_ptrs=table)
            hdl.barrier(channel=1)   # nobody overwrites its map (next call) while a peer may still be reading
        A, K = m.args["anchor_number"], m.args["num_class"]
        nc, nr = A * K, 7 * A
        nchw = heads.permute(0, 3, 1, 2)
        # bytes this rank puts on the wire per scene: its shrunk fp32 map, read once by every peer
        self.wire_bytes = int(heads.shape[1] * heads.shape[2] * eng.c_shrink * 4) * (self.world - 1)
        return {"psm": nchw[:, :nc], "rm": nchw[:, nc:nc + nr], "obj": nchw[:, nc + nr:nc + nr + A]}


class AgentParallelV2XVit(AgentParallelCoBEVT):
    """Airv2xV2XVit (BASELINE config 3) with the agents of ONE scene sharded one per rank: per-agent encoders in parallel,
    one all-gather of the shrunk maps (the path's only exchange: v2xvit fuses after `regroup`, airv2x_v2xvit.py:117-140),
    replicated V2XTransformer. Every rank returns the fused output, bit-equal to the single-GPU one. NCCL transport (the
    peer-memory pull is fused into CoBEVT's regroup kernel; V2X-ViT's first consumer is the RTE / STTF copy)."""

    def __init__(self, model, agent_types, transport="nccl", group=None):
        assert transport == "nccl", "AgentParallelV2XVit: NCCL transport only"
        super().__init__(model, agent_types, transport, group)

    def __call__(self, points, preprocess, prior, scm):
        """points: [P, 4] cloud of THIS rank's agent; prior [1, L, 3] / scm [1, L, 4, 4]: the scene's prior encoding and
        spatial correction matrices (replicated on every rank, as data_dict carries them)."""
        from .ops import Act
        from .w2c_engine import HEAD_PAD
        from . import ops

        m = self.model
        assert not m.training, "agent-parallel mode is inference only"
        dev = next(m.parameters()).device
        if not hasattr(self, "_lay"):
            self._lay = self._layouts(dev)
        lay, glob = self._lay
        dd = dict(self.per_rank[self.rank])
        pts = points.to(device=dev, dtype=torch.float32)
        dd["raw_points"] = {"points": pts, "offsets": torch.tensor([0, pts.shape[0]], dtype=torch.int32),
                            "preprocess": preprocess, "filter": True}
        lidar = m._lidar(dd, dev, lay)
        lidar["raw"]["ego_flags"] = lay["ego_flags"]
        eng, P = m.engine, m._param_dict()
        eng._begin_step()
        W = eng._pack_weights(P)
        local = eng.encode(P, W, lidar, lay)
        feat = gather_agent_maps(local, self.group)
        fused = eng.fusion(P, W, feat, glob, prior, scm)
        heads = eng._buf("heads.out", (fused.shape[0], feat.shape[1], feat.shape[2], HEAD_PAD))
        ops.linear_fwd(fused, W["heads"], Act(heads), bias=W["heads.bias"])
        A, K = m.args["anchor_number"], m.args["num_class"]
        nc, nr = A * K, 7 * A
        nchw = heads.permute(0, 3, 1, 2)
        self.wire_bytes = int(heads.shape[1] * heads.shape[2] * eng.c_shrink * 4) * (self.world - 1)
        return {"psm": nchw[:, :nc], "rm": nchw[:, nc:nc + nr], "obj": nchw[:, nc + nr:nc + nr + A]}


class AgentParallelWhere2comm:
    """Airv2xWhere2com inference with the agents of ONE scene sharded one per rank (rank 0 = ego). Each rank publishes
    the level-0 cells its communication mask selected (warp-ballot compaction, `a2x_mask_compact`) and its dense deeper
    levels in ONE buffer; transports: "nccl" (one all_gather_into_tensor of the buffers) or "peer" (symmetric memory:
    the receivers' decompaction / copy kernels pull the records over NVLink, bytes on the wire ~ mask rate)."""

    def __init__(self, model, agent_types, transport="nccl", group=None):
        assert transport in ("nccl", "peer")
        self.model, self.group, self.transport = model, group, transport
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert len(agent_types) == self.world, "one agent per rank"
        self.agent_types = list(agent_types)
        self.per_rank, _ = agent_rank_plan(agent_types)
        self._state = None

    def _setup(self, dev):
        m = self.model
        skeleton = dict(self.per_rank[self.rank])
        skeleton["raw_points"] = True
        lay = dict(m._layout(skeleton, dev))
        lay["ego_flags"] = torch.tensor([1 if self.rank == 0 else 0], dtype=torch.uint8, device=dev)
        h2, w2 = lay["ny"] // 2, lay["nx"] // 2
        total = m.engine.ap_regions(h2, w2)["total"]
        if self.transport == "peer":
            import torch.distributed._symmetric_memory as symm_mem

            buf = symm_mem.empty(total, dtype=torch.float32, device=dev)
            hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            table = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=dev)

            def exchange():
                hdl.barrier(channel=0)
                return table

            def done():
                hdl.barrier(channel=1)
        else:
            buf = torch.empty(total, dtype=torch.float32, device=dev)
            gathered = torch.empty(self.world, total, dtype=torch.float32, device=dev)
            table = torch.tensor([gathered.data_ptr() + r * total * 4 for r in range(self.world)], dtype=torch.int64,
                                 device=dev)

            def exchange():
                dist.all_gather_into_tensor(gathered, buf, group=self.group)
                return table

            def done():
                pass
        self._state = (lay, buf, exchange, done)

    def __call__(self, points, preprocess):
        """points: [P, 4] f32 cloud of THIS rank's agent. Returns the reference's output dict (same on every rank)."""
        m = self.model
        assert not m.training, "agent-parallel mode is inference only"
        dev = next(m.parameters()).device
        if self._state is None:
            self._setup(dev)
        lay, buf, exchange, done = self._state
        dd = dict(self.per_rank[self.rank])
        pts = points.to(device=dev, dtype=torch.float32)
        dd["raw_points"] = {"points": pts, "offsets": torch.tensor([0, pts.shape[0]], dtype=torch.int32),
                            "preprocess": preprocess, "filter": True}
        lidar = m._lidar(dd, dev, lay)
        lidar["raw"]["ego_flags"] = lay["ego_flags"]
        heads, aux = m.engine.forward_agent_parallel(m._param_dict(), lidar, lay, self.rank, self.world, buf, exchange, done)
        m._last_aux = aux
        self._buf = buf
        return m._output_dict(heads, {"record_len": [self.world]})

    @property
    def wire_bytes(self):
        """bytes this rank put on the wire in the last call, per peer x (world - 1): the NCCL transport gathers the whole
        fixed-capacity buffer; the peer transport reads only the records the communication mask selected (header count x
        (index + 64-float row)) plus the dense deeper levels. Reads one header word (host sync): call it outside timed loops."""
        m = self.model
        lay = self._state[0]
        reg = m.engine.ap_regions(lay["ny"] // 2, lay["nx"] // 2)
        if self.transport == "nccl":
            return int(reg["total"] * 4) * (self.world - 1)
        count = int(self._buf[reg["hdr"][0]:reg["hdr"][0] + 1].view(torch.int32).item())
        dense = sum(v[1] for k, v in reg.items() if k.startswith("lvl"))
        return int(64 * 4 + count * (4 + m.engine.num_filters[0] * 4) + dense * 4) * (self.world - 1)
