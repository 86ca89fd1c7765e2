"""`opencood.models.airv2x_where2com.Airv2xWhere2com` on the B200 kernels.

Same registry name / class name / constructor (`cls(hypes["model"]["args"])`), same `hypes_yaml` keys, same
`state_dict` key names and shapes, same `forward(data_dict) -> {"psm","rm","obj","mask","com","comm_rate"}` as
opencood/models/airv2x_where2com.py:20-179 of the reference, so `train_utils.create_model`
(opencood/tools/train_utils.py:288-325) dispatches here unchanged once `install()` has registered the module.
The torch.nn layers below are parameter containers only (names + default init); their forward is never called.
There is no CPU / eager fallback: without a CUDA device or the built library the forward raises.
"""
import math
import random

import numpy as np
import torch
import torch.nn as nn

from ...w2c_engine import AGENT_TYPES, HEAD_PAD, TYPE_PREFIX, W2CEngine


class _PFNParams(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear = nn.Linear(cin, cout, bias=False)
        self.norm = nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01)


class _PillarVFEParams(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        assert cfg["use_norm"] and cfg["use_absolute_xyz"] and not cfg["with_distance"] and list(cfg["num_filters"]) == [64], \
            "only the airv2x PillarVFE configuration (10 -> 64, BN, absolute xyz) is implemented"
        self.pfn_layers = nn.ModuleList([_PFNParams(10, 64)])


def _backbone_params(cfg, cin):
    blocks, deblocks = nn.ModuleList(), nn.ModuleList()
    chans = [cin] + list(cfg["num_filters"])[:-1]
    for i, (ln, st, cf) in enumerate(zip(cfg["layer_nums"], cfg["layer_strides"], cfg["num_filters"])):
        layers = [nn.ZeroPad2d(1), nn.Conv2d(chans[i], cf, 3, stride=st, padding=0, bias=False),
                  nn.BatchNorm2d(cf, eps=1e-3, momentum=0.01), nn.ReLU()]
        for _ in range(ln):
            layers += [nn.Conv2d(cf, cf, 3, padding=1, bias=False), nn.BatchNorm2d(cf, eps=1e-3, momentum=0.01), nn.ReLU()]
        blocks.append(nn.Sequential(*layers))
        us, uf = cfg["upsample_strides"][i], cfg["num_upsample_filter"][i]
        deblocks.append(nn.Sequential(nn.ConvTranspose2d(cf, uf, us, stride=us, bias=False),
                                      nn.BatchNorm2d(uf, eps=1e-3, momentum=0.01), nn.ReLU()))
    m = nn.Module()
    m.blocks, m.deblocks = blocks, deblocks
    return m


def _shrink_params(cfg):
    m = nn.Module()
    m.layers = nn.ModuleList()
    cin = cfg["input_dim"]
    for k, d, s, p in zip(cfg["kernal_size"], cfg["dim"], cfg["stride"], cfg["padding"]):
        dc = nn.Module()
        dc.double_conv = nn.Sequential(nn.Conv2d(cin, d, k, stride=s, padding=p), nn.ReLU(inplace=True),
                                       nn.Conv2d(d, d, 3, padding=1), nn.ReLU(inplace=True))
        m.layers.append(dc)
        cin = d
    return m


def _comm_params(cfg):
    fusion = nn.Module()
    comm = nn.Module()
    if "gaussian_smooth" in cfg["communication"]:
        k = cfg["communication"]["gaussian_smooth"]["k_size"]
        sigma = cfg["communication"]["gaussian_smooth"]["c_sigma"]
        comm.gaussian_filter = nn.Conv2d(1, 1, k, stride=1, padding=(k - 1) // 2)
        c = k // 2
        ys, xs = np.mgrid[0 - c:k - c, 0 - c:k - c]
        g = 1 / (2 * np.pi * sigma) * np.exp(-(np.square(ys) + np.square(xs)) / (2 * np.square(sigma)))
        comm.gaussian_filter.weight.data = torch.Tensor(g).unsqueeze(0).unsqueeze(0)
        comm.gaussian_filter.bias.data.zero_()
    fusion.naive_communication = comm
    return fusion


class Airv2xWhere2com(nn.Module):
    def __init__(self, args, precision="split3"):
        super().__init__()
        self.args = args
        self.collaborators = args["collaborators"]
        self.active_sensors = args["active_sensors"]
        self.veh_models, self.rsu_models, self.drone_models = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for t in AGENT_TYPES:
            if t not in self.collaborators:
                continue
            for m in args[t]["modalities"]:
                if m != "lidar":
                    raise NotImplementedError("modality %r is outside the B200 hot path (lidar only)" % m)
                enc = nn.Sequential(_PillarVFEParams(args[t]["lidar"]["pillar_vfe"]), nn.Identity())
                getattr(self, TYPE_PREFIX[t]).append(enc)
        mf = args["modality_fusion"]
        self.backbone = _backbone_params(mf["base_bev_backbone"], 64)
        self.shrink_flag = bool(mf.get("shrink_header", {}).get("use", False))
        if self.shrink_flag:
            self.shrink_conv = _shrink_params(mf["shrink_header"])
        self.compression = mf["compression"] > 0
        self.fusion_net = _comm_params(args["where2com_fusion"])
        self.multi_scale = args["where2com_fusion"]["multi_scale"]
        self.outC = args["outC"]
        if args["task"] != "det":
            raise NotImplementedError("task %r is outside the B200 hot path (det only)" % args["task"])
        self.cls_head = nn.Conv2d(self.outC, args["anchor_number"] * args["num_class"], kernel_size=1)
        self.reg_head = nn.Conv2d(self.outC, 7 * args["anchor_number"], kernel_size=1)
        if args["obj_head"]:
            self.obj_head = nn.Conv2d(self.outC, args["anchor_number"], kernel_size=1)
        self.precision = precision
        self._engine = None
        self._last_aux = None
        if args.get("backbone_fix", False):
            self.backbone_fix()

    def backbone_fix(self):
        for n, p in self.named_parameters():
            if not n.startswith("fusion_net"):
                p.requires_grad = False

    # ------------------------------------------------------------------ plumbing
    @property
    def engine(self):
        if self._engine is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("Airv2xWhere2com (B200) needs its parameters on a CUDA device; there is no CPU path")
            self._engine = W2CEngine(self.args, dev, self.precision)
        return self._engine

    def _param_dict(self):
        d = {n: p.data for n, p in self.named_parameters()}
        d.update({n: b for n, b in self.named_buffers()})
        return d

    def _layout(self, data_dict, device):
        """Scene-major agent order (vehicles, RSUs, drones per scene): airv2x_base_model.py:179-248.
        Cached per (record_len, batch_idxs) signature so steady-state steps create no new device tensors."""
        raw0 = data_dict.get("raw_points")
        key = []
        for t in AGENT_TYPES:
            d = data_dict.get(t)
            if d is None:
                continue
            r = d["record_len"]
            key.append((t, tuple(int(v) for v in (r.tolist() if torch.is_tensor(r) else r)), tuple(d["batch_idxs"]),
                        raw0 is not None or d.get("batch_merged_lidar_features_torch") is not None))
        key = (tuple(key), str(device))
        cache = self.__dict__.setdefault("_layout_cache", {})
        if key not in cache:
            cache[key] = self._build_layout(data_dict, device)
        return cache[key]

    def _build_layout(self, data_dict, device):
        rl, idxs = {}, {}
        raw = data_dict.get("raw_points")
        for t in AGENT_TYPES:
            if t in self.collaborators and len(data_dict[t]["batch_idxs"]) > 0 and \
                    (raw is not None or data_dict[t].get("batch_merged_lidar_features_torch") is not None):
                r = data_dict[t]["record_len"]
                rl[t] = [int(v) for v in (r.tolist() if torch.is_tensor(r) else r)]
                idxs[t] = list(data_dict[t]["batch_idxs"])
        B = max(len(v) for v in idxs.values())
        starts = {}
        for t in rl:
            pos, st = 0, {}
            for b in idxs[t]:
                st[b] = pos
                pos += rl[t][b]
            starts[t] = st
        amap = {t: [0] * sum(rl[t][b] for b in idxs[t]) for t in rl}
        record_len, row = [], 0
        for b in range(B):
            n_b = 0
            for t in AGENT_TYPES:
                if t in rl and b in idxs[t]:
                    for j in range(rl[t][b]):
                        amap[t][starts[t][b] + j] = row
                        row += 1
                    n_b += rl[t][b]
            record_len.append(n_b)
        first = next(iter(rl))
        nx, ny, _ = [int(v) for v in self.args[first]["lidar"]["point_pillar_scatter"]["grid_size"]]
        scene_start = np.concatenate([[0], np.cumsum(record_len)[:-1]]).astype(np.int32)
        types = [None] * row
        for t in amap:
            for r in amap[t]:
                types[r] = t
        return dict(n_total=row, nx=nx, ny=ny, record_len=record_len, types=types,
                    identity_map=torch.arange(row, dtype=torch.int32, device=device),
                    ego_flags=torch.tensor([1 if i in set(int(v) for v in scene_start) else 0 for i in range(row)],
                                           dtype=torch.uint8, device=device),
                    agent_map={t: torch.tensor(v, dtype=torch.int32, device=device) for t, v in amap.items()},
                    scene_start=torch.tensor(scene_start, dtype=torch.int32, device=device),
                    scene_len=torch.tensor(record_len, dtype=torch.int32, device=device))

    def _lidar(self, data_dict, device, layout):
        """the engine's input dict: voxels (or raw clouds) per agent type, plus — B200 extension for the camera + lidar
        configs — `data_dict[type]["camera_bev"]` = [n_type, 64, ny, nx], the `spatial_features` of that type's camera encoder
        (lss.CameraBranch = LiftSplatShootEncoder with a pluggable image trunk), averaged with the pillar canvas as the
        reference's fuse_bev does (common_modules/airv2x_base_model.py:167-177). Eval forward only."""
        out = self._lidar_inputs(data_dict, device, layout)
        cam = {}
        for t in layout["agent_map"]:
            v = data_dict.get(t, {}).get("camera_bev") if isinstance(data_dict.get(t), dict) else None
            if v is not None:
                cam[t] = v.to(device=device, dtype=torch.float32).permute(0, 2, 3, 1).contiguous()
        if cam:
            out["camera_bev"] = cam
        return out

    @staticmethod
    def _all_agents_flag(layout):
        """body-box flag of every agent (sensor-frame clouds): one persistent tensor per layout (stable under graph replay)"""
        if "all_flags" not in layout:
            layout["all_flags"] = torch.ones_like(layout["ego_flags"])
        return layout["all_flags"]

    def _lidar_inputs(self, data_dict, device, layout):
        raw = data_dict.get("raw_points")
        if raw is not None:
            # B200 extension of the boundary: raw per-agent clouds instead of CPU-voxelised pillars.
            # raw_points = {"points": [sum P, 4] f32 (all agents, scene-major order), "offsets": int32 [N+1],
            #               optional "preprocess": hypes["preprocess"], optional "filter": True -> apply the dataset's
            #               mask_ego_points + mask_points_by_range on the GPU, optional "transforms": [N,4,4] agent -> ego
            #               poses (the dataset's `transformation_matrix`, intermediate_fusion_dataset.py:592-600)}
            # With "transforms" the clouds are in each agent's SENSOR frame, as the dataset holds them: every agent's own
            # body box is removed, then the points are projected, then range-filtered (the reference's order). Without,
            # the clouds are already in the ego frame and only the ego agent's body box (first agent of a scene) is.
            pre = raw.get("preprocess")
            first = next(iter(layout["agent_map"]))
            la = self.args[first]["lidar"]
            vs = pre["args"]["voxel_size"] if pre else la["voxel_size"]
            rng = pre["cav_lidar_range"] if pre else la["lidar_range"]
            mp = pre["args"]["max_points_per_voxel"] if pre else 32
            mv = (pre["args"]["max_voxel_train" if self.training else "max_voxel_test"] if pre
                  else (32000 if self.training else 70000))
            xf = raw.get("transforms")
            if xf is not None:  # float64 numpy / torch poses -> fp32, like check_numpy_to_torch(...).float()
                xf = torch.as_tensor(xf).to(device=device, dtype=torch.float32).contiguous()
                assert xf.shape == (layout["n_total"], 4, 4), "raw_points['transforms'] must be [N, 4, 4]"
            return {"raw": {"points": raw["points"].to(device=device, dtype=torch.float32, non_blocking=True).contiguous(),
                            "offsets": raw["offsets"].to(device=device, dtype=torch.int32, non_blocking=True).contiguous(),
                            "types": layout["types"], "voxel_size": vs, "lidar_range": rng, "max_points": mp,
                            "max_voxels": mv, "filter": bool(raw.get("filter", False)),
                            "transforms": xf,
                            "ego_flags": (None if not raw.get("filter", False) else
                                          self._all_agents_flag(layout) if xf is not None else layout["ego_flags"])}}
        out = {}
        for t in layout["agent_map"]:
            d = data_dict[t]["batch_merged_lidar_features_torch"]
            out[t] = {"voxel_features": d["voxel_features"].to(device=device, dtype=torch.float32).contiguous(),
                      "voxel_num_points": d["voxel_num_points"].to(device=device, dtype=torch.int32).contiguous(),
                      "voxel_coords": d["voxel_coords"].to(device=device, dtype=torch.int32).contiguous()}
        return out

    # ------------------------------------------------------------------ reference-facing API
    def forward(self, data_dict):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("Airv2xWhere2com (B200) needs its parameters on a CUDA device; there is no CPU path")
        layout = self._layout(data_dict, dev)
        lidar = self._lidar(data_dict, dev, layout)
        B = len(layout["record_len"])
        k_list = None
        if self.training and not self.engine.fully:
            hw = None  # drawn inside the engine from Python `random`, same call order as the reference
        names = [n for n, p in self.named_parameters() if p.requires_grad and not n.startswith("fusion_net")]
        params = [p for n, p in self.named_parameters() if p.requires_grad and not n.startswith("fusion_net")]
        if self.training and torch.is_grad_enabled():
            # autograd boundary = the torch.library ops a2x::fused_forward / a2x::fused_backward (torch_ops.py): inputs
            # are the trainable parameters, the output is the NHWC head-logit tensor
            from ... import torch_ops
            eng = self.engine

            def run_forward():
                h, self._last_aux = eng.forward(self._param_dict(), lidar, layout, True, k_list)
                return h

            def run_backward(dheads):
                P = self._param_dict()
                grads = {n: torch.zeros_like(P[n]) for n in names}
                eng.backward(P, dheads, grads)
                return [grads[n] for n in names]

            key = torch_ops.bind(self, run_forward, run_backward)
            heads = torch.ops.a2x.fused_forward(key, params, self._heads_shape(layout))
        else:
            heads, self._last_aux = self.engine.forward(self._param_dict(), lidar, layout, self.training, k_list)
        return self._output_dict(heads, layout)

    def _heads_shape(self, layout):
        """[B, H/2, W/2, HEAD_PAD] of the fused head-logit tensor (what the ops' fake implementation returns)"""
        return [len(layout["record_len"]), layout["ny"] // 2, layout["nx"] // 2, HEAD_PAD]

    # ------------------------------------------------------------------ fused training step (public fast path)
    def _grad_buffers(self):
        """Persistent .grad tensors (overwritten every step; same effect as zero_grad() + backward())."""
        g = {}
        for n, p in self.named_parameters():
            if n.startswith("fusion_net") or not p.requires_grad:
                continue
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            g[n] = p.grad
        return g

    def attach_grad_sync(self):
        """Scene-parallel training (one process per GPU, the reference wraps the model in DDP: tools/train.py:161-163):
        back every p.grad by one flat buffer and average it over the ranks INSIDE the fused step, the bulk of it overlapped
        with the level-0 backward (dist.GradAverager). Call once, before the first step (it is part of the captured CUDA
        graph); a no-op when torch.distributed is not initialised with world_size > 1."""
        from ... import dist as D
        pairs = [(n, p) for n, p in self.named_parameters() if p.requires_grad and not n.startswith("fusion_net")]
        late = lambda n: n.startswith(("backbone.blocks.0.", "veh_models.", "rsu_models.", "drone_models."))
        self.grad_sync = D.GradAverager([p for _, p in pairs], late=late, names=[n for n, _ in pairs])
        self.__dict__.pop("_graphs", None)       # graphs captured before hold the old gradient pointers
        return self.grad_sync

    def prepare_labels(self, label_dict, device):
        """label tensors of the reference's collate (fp64 targets / pos_equal_one, int64 class_ids:
        data_utils/post_processor/voxel_postprocessor.py:392-430) -> device fp32 / int32, contiguous."""
        return {"targets": label_dict["targets"].to(device=device, dtype=torch.float32, non_blocking=True).contiguous(),
                "pos_equal_one": label_dict["pos_equal_one"].to(device=device, dtype=torch.float32,
                                                                non_blocking=True).contiguous(),
                "class_ids": label_dict["class_ids"].to(device=device, dtype=torch.int32, non_blocking=True).contiguous()}

    def train_step(self, data_dict, label_dict, cls_weight=1.0, reg_coe=2.0, k_list=None):
        """forward (train-mode BN) + PointPillarLossMultiClass + backward in one call, all on the CUDA kernels.
        Inputs may live on the host (pinned memory -> async H2D here) or on the device. Parameter gradients land
        in p.grad; returns the device tensor [reg, cls, obj] loss terms (float64) — total = .sum(). The tensor is the
        engine's persistent buffer (a CUDA-graph replay writes it in place): `.clone()` it to keep a value across steps."""
        assert self.training, "train_step() needs model.train()"
        if self.args.get("backbone_fix", False):
            raise NotImplementedError("backbone_fix: true freezes everything but fusion_net (airv2x_where2com.py:84-92), which "
                                      "has no trainable parameter on this path: there is nothing to train")
        dev = next(self.parameters()).device
        layout = self._layout(data_dict, dev)
        lidar = self._lidar(data_dict, dev, layout)
        labels = self.prepare_labels(label_dict, dev)
        P = self._param_dict()
        heads, self._last_aux = self.engine.forward(P, lidar, layout, True, k_list)
        loss3, dheads = self.engine.loss(heads, labels, cls_weight, reg_coe)
        self.engine.backward(P, dheads, self._grad_buffers(), sync=self.__dict__.get("grad_sync"))
        self._last_layout = layout
        return loss3

    # ------------------------------------------------------------------ CUDA-graph replay of the fused step
    max_graphs = 32   # captured step graphs kept per model (one per agent layout x loss weights x input frame)

    def train_step_graphed(self, data_dict, label_dict, cls_weight=1.0, reg_coe=2.0):
        """train_step() captured once into a CUDA graph (raw-point input only): per step the host copies the clouds /
        labels (and, for sensor-frame clouds, the agent -> ego poses) into static buffers (pinned -> device), draws the top-K
        sizes and replays ~250 kernel launches with one cudaGraphLaunch. Captures again when the agent layout or the cloud
        capacity changes."""
        assert self.training and data_dict.get("raw_points") is not None, "graphed step needs raw_points input"
        import random as _random

        dev = next(self.parameters()).device
        raw = data_dict["raw_points"]
        layout = self._layout(data_dict, dev)
        P = int(raw["points"].shape[0])
        graphs = self.__dict__.setdefault("_graphs", {})
        # sensor-frame clouds + poses (the dataset's batches) replay their own graph: the projection is part of the
        # captured voxeliser launch and reads the poses from a static [N,4,4] buffer refreshed before every replay
        key = (id(layout), float(cls_weight), float(reg_coe), raw.get("transforms") is not None)
        st = graphs.get(key)
        if st is None or st["cap"] < P:
            st = self._capture(data_dict, label_dict, cls_weight, reg_coe, layout, dev, max(P, int(P * 1.1)))
            graphs.pop(key, None)
            graphs[key] = st
            while len(graphs) > self.max_graphs:      # a dataset with many agent layouts: drop the oldest captures
                graphs.pop(next(iter(graphs)))
        st["points"][:P].copy_(raw["points"], non_blocking=True)
        st["offsets"].copy_(raw["offsets"], non_blocking=True)
        if st["transforms"] is not None:
            st["transforms"].copy_(torch.as_tensor(raw["transforms"]).to(torch.float32).reshape(st["transforms"].shape),
                                   non_blocking=True)
        for k in ("targets", "pos_equal_one", "class_ids"):
            st["labels"][k].copy_(label_dict[k].reshape(st["labels"][k].shape), non_blocking=True)
        eng = self.engine
        if not eng.fully:
            hw = st["hw"]
            k_host = eng.set_k([int(hw * _random.uniform(0, 1)) for _ in layout["record_len"]], layout["record_len"])
            eng._buf("k_dev", (k_host.numel(),), torch.int32).copy_(k_host, non_blocking=True)
        st["graph"].replay()
        self._last_aux = st["aux"]
        return st["loss3"]

    # ------------------------------------------------------------------ pipelined input staging (overlap H2D / D2H)
    def stage_inputs(self, data_dict, label_dict, cls_weight=1.0, reg_coe=2.0):
        """Start the host->device copies of the NEXT step's inputs (pinned host tensors) on a copy stream, so they
        overlap the step that is currently running. Follow with `train_step_staged()`. The top-K sizes of the
        communication mask are drawn here from Python's `random`, in call order, exactly like the reference."""
        import random as _random

        assert self.training and data_dict.get("raw_points") is not None
        dev = next(self.parameters()).device
        raw = data_dict["raw_points"]
        if raw.get("transforms") is not None:
            raise NotImplementedError("the pipelined staging takes ego-frame clouds; use train_step_graphed() with raw_points['transforms']")
        layout = self._layout(data_dict, dev)
        P = int(raw["points"].shape[0])
        graphs = self.__dict__.setdefault("_graphs", {})
        key = (id(layout), float(cls_weight), float(reg_coe), False)          # the graph train_step_graphed uses too
        st = graphs.get(key)
        if st is None or st["cap"] < P:
            st = self._capture(data_dict, label_dict, cls_weight, reg_coe, layout, dev, max(P, int(P * 1.1)))
            graphs[key] = st
        if "stage" not in st:
            n_agents = layout["n_total"]
            st["stage"] = dict(points=torch.empty_like(st["points"]), offsets=torch.empty_like(st["offsets"]),
                               labels={k: torch.empty_like(v) for k, v in st["labels"].items()},
                               k=torch.zeros(n_agents, dtype=torch.int32, device=dev),
                               k_host=[torch.zeros(n_agents, dtype=torch.int32).pin_memory() for _ in range(2)], flip=0,
                               stream=torch.cuda.Stream(device=dev), ready=torch.cuda.Event(), consumed=torch.cuda.Event(),
                               loss_host=[torch.zeros(3, dtype=torch.float64).pin_memory() for _ in range(2)],
                               loss_evt=[torch.cuda.Event(), torch.cuda.Event()])
            st["stage"]["consumed"].record()
        sg = st["stage"]
        eng = self.engine
        if not eng.fully:
            kh = sg["k_host"][sg["flip"]]
            pos = 0
            for b, n in enumerate(layout["record_len"]):
                kh[pos:pos + n] = int(st["hw"] * _random.uniform(0, 1))
                pos += n
        with torch.cuda.stream(sg["stream"]):
            sg["stream"].wait_event(sg["consumed"])      # the previous staged batch has been moved to the static buffers
            sg["points"][:P].copy_(raw["points"], non_blocking=True)
            sg["offsets"].copy_(raw["offsets"], non_blocking=True)
            for k in ("targets", "pos_equal_one", "class_ids"):
                sg["labels"][k].copy_(label_dict[k].reshape(sg["labels"][k].shape), non_blocking=True)
            if not eng.fully:
                sg["k"].copy_(sg["k_host"][sg["flip"]], non_blocking=True)
            sg["ready"].record()
        sg["P"] = P
        self._staged = st
        return st

    def train_step_staged(self):
        """Run the step whose inputs `stage_inputs()` copied: device-to-device hand-over into the graph's static buffers,
        one graph replay, and an asynchronous device->host copy of the [reg, cls, obj] loss. Returns a handle whose
        `.result()` waits for that copy only (by then usually complete: call it after launching the next step)."""
        st = self._staged
        sg = st["stage"]
        eng = self.engine
        cur = torch.cuda.current_stream()
        cur.wait_event(sg["ready"])
        st["points"][:sg["P"]].copy_(sg["points"][:sg["P"]], non_blocking=True)
        st["offsets"].copy_(sg["offsets"], non_blocking=True)
        for k in ("targets", "pos_equal_one", "class_ids"):
            st["labels"][k].copy_(sg["labels"][k], non_blocking=True)
        if not eng.fully:
            eng._buf("k_dev", (sg["k"].numel(),), torch.int32).copy_(sg["k"], non_blocking=True)
        sg["consumed"].record()
        if not st.get("k_free_graph", False):
            raise RuntimeError("stage_inputs() must capture the graph (call it before train_step_graphed on this layout)")
        st["graph"].replay()
        i = sg["flip"]
        sg["flip"] ^= 1
        sg["loss_host"][i].copy_(st["loss3"], non_blocking=True)
        sg["loss_evt"][i].record()
        self._last_aux = st["aux"]
        host, evt = sg["loss_host"][i], sg["loss_evt"][i]

        class _Handle:
            def result(self_inner):
                evt.synchronize()
                return host.clone()

        return _Handle()

    def _capture(self, data_dict, label_dict, cw, rc, layout, dev, cap):
        raw = data_dict["raw_points"]
        P = int(raw["points"].shape[0])
        pts = torch.zeros(cap, 4, device=dev)
        pts[:P].copy_(raw["points"])
        offs = raw["offsets"].to(device=dev, dtype=torch.int32).clone()
        labels = {k: v.clone() for k, v in self.prepare_labels(label_dict, dev).items()}
        dd = {k: v for k, v in data_dict.items() if k != "raw_points"}
        dd["raw_points"] = dict(raw)
        dd["raw_points"]["points"] = pts
        dd["raw_points"]["offsets"] = offs
        xf = None
        if raw.get("transforms") is not None:  # static pose buffer: _lidar_inputs passes a device fp32 tensor through as is
            xf = torch.as_tensor(raw["transforms"]).to(device=dev, dtype=torch.float32).reshape(layout["n_total"], 4, 4).contiguous().clone()
            dd["raw_points"]["transforms"] = xf
        rng_state = random.getstate()  # warm-up / capture must not consume the caller's top-K random stream
        # ... nor move the BatchNorm running statistics: the eager warm-up steps below really run (twice) on this batch
        bn_state = {n: b.clone() for n, b in self.named_buffers()}
        for _ in range(2):  # eager warm-up: every buffer / stream / attribute exists before capture
            self.train_step(dd, labels, cw, rc)
        torch.cuda.synchronize()
        with torch.no_grad():
            for n, b in self.named_buffers():
                b.copy_(bn_state[n])
        from ... import _lib

        lib = _lib.load()
        g = torch.cuda.CUDAGraph()
        l0 = lib.a2x_launch_count()
        # the top-K sizes reach `k_dev` by an explicit copy issued before every replay (from the pinned `k_host`, or from
        # the staged device copy in pipelined mode), not from inside the graph
        self.engine.k_on_device = True
        try:
            with torch.cuda.graph(g):
                loss3 = self.train_step(dd, labels, cw, rc)
        finally:
            self.engine.k_on_device = False
        self.launches_per_step = int(lib.a2x_launch_count() - l0)  # kernels of this library inside one replay
        random.setstate(rng_state)
        return dict(graph=g, points=pts, offsets=offs, transforms=xf, labels=labels, loss3=loss3, aux=self._last_aux, cap=cap,
                    hw=self._last_aux["hw"], k_free_graph=True)

    def _output_dict(self, heads, layout):
        A, K = self.args["anchor_number"], self.args["num_class"]
        nc, nr = A * K, 7 * A
        nchw = heads.permute(0, 3, 1, 2)  # logical NCHW views of the NHWC logits (values identical)
        out = {"psm": nchw[:, :nc], "rm": nchw[:, nc:nc + nr], "obj": nchw[:, nc + nr:nc + nr + A]}
        aux = self._last_aux
        rl = torch.tensor(layout["record_len"], dtype=torch.float32, device=heads.device)
        if self.engine.fully:
            com = torch.tensor(1, device=heads.device)
        else:
            com = (aux["ones"] / (rl * aux["hw"])).sum() / len(layout["record_len"])
        out.update({"mask": 0, "com": com, "comm_rate": int(aux["comm_rate"].item())})
        return out
