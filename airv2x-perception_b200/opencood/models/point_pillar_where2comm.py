"""`opencood.models.point_pillar_where2comm.PointPillarWhere2comm` on the B200 kernels (BASELINE config 1).

Same registry name / class name / constructor, same `hypes_yaml` keys (`pillar_vfe`, `point_pillar_scatter`,
`base_bev_backbone`, `shrink_header`, `where2comm_fusion`, `head_dim`, `anchor_number`, `compression`), same
`state_dict` keys and shapes (8 057 386 parameters for V2XR_where2comm.yaml), same
`forward(data_dict) -> {"psm","rm","com","mask","each_mask","comm_rate"}` as opencood/models/point_pillar_where2comm.py
of the reference (input: `data_dict["processed_lidar"]`, `record_len`, `pairwise_t_matrix`). Parameter containers only.
Eval forward, train-mode forward wired to autograd (torch.library ops, so the reference's loop `model(batch)` ->
PointPillarLoss -> `loss.backward()` works) and the fused `train_step` (forward + PointPillarLoss kernel + backward); no
CPU fallback.
"""
import numpy as np
import torch
import torch.nn as nn

from ...ppw2c_engine import LegacyW2CEngine
from ...w2c_engine import HEAD_PAD
from .airv2x_where2com import _backbone_params, _comm_params, _PillarVFEParams


def _legacy_shrink_params(cfg):
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


class PointPillarWhere2comm(nn.Module):
    def __init__(self, args, precision="split3"):
        super().__init__()
        self.args = args
        self.modality = args.get("use_modality", "processed_lidar")
        self.max_cav = args["max_cav"]
        self.pillar_vfe = _PillarVFEParams(args["pillar_vfe"])
        self.backbone = _backbone_params(args["base_bev_backbone"], 64)
        self.shrink_flag = "shrink_header" in args
        if not self.shrink_flag:
            raise NotImplementedError("point_pillar_where2comm without a shrink header is not implemented")
        self.shrink_conv = _legacy_shrink_params(args["shrink_header"])
        if args["compression"]:
            raise NotImplementedError("NaiveCompressor (compression > 0) is not implemented for this model")
        self.compression = False
        self.fusion_net = _comm_params(args["where2comm_fusion"])
        self.multi_scale = args["where2comm_fusion"]["multi_scale"]
        self.cls_head = nn.Conv2d(args["head_dim"], args["anchor_number"], kernel_size=1)
        self.reg_head = nn.Conv2d(args["head_dim"], 7 * args["anchor_number"], kernel_size=1)
        self.precision = precision
        self._engine = None
        if args.get("backbone_fix", False):
            for n, p in self.named_parameters():
                if not n.startswith("fusion_net"):
                    p.requires_grad = False

    @property
    def engine(self):
        if self._engine is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("PointPillarWhere2comm (B200) needs its parameters on a CUDA device; there is no CPU path")
            self._engine = LegacyW2CEngine(self.args, dev, self.precision)
        return self._engine

    def _param_dict(self):
        d = {n: p.data for n, p in self.named_parameters()}
        d.update({n: b for n, b in self.named_buffers()})
        return d

    def _inputs(self, data_dict, dev):
        """-> (lidar dict, layout). B200 extension of the boundary (like the airv2x models): data_dict["raw_points"] =
        {"points" [sum P, 4] f32 ego-frame clouds in agent order, "offsets" int32 [N+1], "preprocess": hypes["preprocess"],
        "filter": True -> the dataset's mask_ego_points (ego = first agent of a scene) + mask_points_by_range on the GPU}
        instead of the CPU-voxelised `processed_lidar`; voxelisation then runs on the GPU, bit-exact."""
        raw = data_dict.get("raw_points")
        lid = None if raw is not None else data_dict[self.modality]
        r = data_dict["record_len"]
        record_len = [int(v) for v in (r.tolist() if torch.is_tensor(r) else r)]
        key = tuple(record_len)
        cache = self.__dict__.setdefault("_layout_cache", {})
        if key not in cache:
            nx, ny, _ = [int(v) for v in self.args["point_pillar_scatter"]["grid_size"]]
            n = sum(record_len)
            start = np.concatenate([[0], np.cumsum(record_len)[:-1]]).astype(np.int32)
            cache[key] = dict(n_total=n, nx=nx, ny=ny, record_len=record_len,
                              identity_map=torch.arange(n, dtype=torch.int32, device=dev),
                              scene_start=torch.tensor(start, dtype=torch.int32, device=dev),
                              scene_len=torch.tensor(record_len, dtype=torch.int32, device=dev))
        layout = cache[key]
        if raw is not None:
            pre = raw["preprocess"]
            starts = set(int(v) for v in np.concatenate([[0], np.cumsum(record_len)[:-1]]))
            ego = torch.tensor([1 if i in starts else 0 for i in range(layout["n_total"])], dtype=torch.uint8, device=dev)
            return {"raw": {"points": raw["points"].to(device=dev, dtype=torch.float32, non_blocking=True).contiguous(),
                            "offsets": raw["offsets"].to(device=dev, dtype=torch.int32, non_blocking=True).contiguous(),
                            "voxel_size": pre["args"]["voxel_size"], "lidar_range": pre["cav_lidar_range"],
                            "max_points": pre["args"]["max_points_per_voxel"],
                            "max_voxels": pre["args"]["max_voxel_train" if self.training else "max_voxel_test"],
                            "filter": bool(raw.get("filter", False)), "transforms": None,
                            "ego_flags": ego if raw.get("filter", False) else None}}, layout
        lidar = {"voxel_features": lid["voxel_features"].to(device=dev, dtype=torch.float32).contiguous(),
                 "voxel_num_points": lid["voxel_num_points"].to(device=dev, dtype=torch.int32).contiguous(),
                 "voxel_coords": lid["voxel_coords"].to(device=dev, dtype=torch.int32).contiguous()}
        return lidar, layout

    def forward(self, data_dict):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("PointPillarWhere2comm (B200) needs its parameters on a CUDA device; there is no CPU path")
        lidar, layout = self._inputs(data_dict, dev)
        record_len = layout["record_len"]
        eng = self.engine
        if self.training and torch.is_grad_enabled():
            # reference training loop: model(batch) -> PointPillarLoss -> loss.backward(); autograd boundary = torch_ops
            from ... import torch_ops
            names = [n for n, p in self.named_parameters() if p.requires_grad and not n.startswith("fusion_net")]
            params = [p for n, p in self.named_parameters() if p.requires_grad and not n.startswith("fusion_net")]
            st = {}

            def run_forward():
                h, st["aux"] = eng.forward(self._param_dict(), lidar, layout, True)
                return h

            def run_backward(dheads):
                P = self._param_dict()
                grads = {n: torch.zeros_like(P[n]) for n in names}
                eng.backward(P, dheads, grads)
                return [grads[n] for n in names]

            s = self.engine.shrink_stride
            hh, ww = layout["ny"] // 2, layout["nx"] // 2
            shape = [len(record_len), (hh - 1) // s + 1, (ww - 1) // s + 1, HEAD_PAD]
            heads = torch.ops.a2x.fused_forward(torch_ops.bind(self, run_forward, run_backward), params, shape)
            aux = st["aux"]
        else:
            heads, aux = eng.forward(self._param_dict(), lidar, layout, self.training)
        A = self.args["anchor_number"]
        nchw = heads.permute(0, 3, 1, 2)
        rl = torch.tensor(record_len, dtype=torch.float32, device=dev)
        if self.engine.fully:
            com = torch.tensor(1, device=dev)
        else:
            com = (aux["ones"] / (rl * aux["hw"])).sum() / len(record_len)
        return {"psm": nchw[:, :A], "rm": nchw[:, A:8 * A], "com": com, "mask": 0, "each_mask": 0,
                "comm_rate": int(aux["comm_rate"].item())}

    def _grad_buffers(self):
        g = {}
        for n, p in self.named_parameters():
            if n.startswith("fusion_net") or not p.requires_grad:
                continue
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            g[n] = p.grad
        return g

    def train_step(self, data_dict, label_dict, cls_weight=1.0, reg_coe=2.0, k_list=None):
        """forward (train-mode BatchNorm, top-K communication mask) + PointPillarLoss (loss/point_pillar_loss.py:77-215) +
        backward in one call on the CUDA kernels; label_dict = the legacy collate's {"targets" [B,H,W,7A], "pos_equal_one"
        [B,H,W,A]}. Parameter gradients land in p.grad; returns the device tensor [reg, conf, 0] (float64), total = .sum()."""
        assert self.training, "train_step() needs model.train()"
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("PointPillarWhere2comm (B200) needs its parameters on a CUDA device; there is no CPU path")
        lidar, layout = self._inputs(data_dict, dev)
        labels = {"targets": label_dict["targets"].to(device=dev, dtype=torch.float32, non_blocking=True).contiguous(),
                  "pos_equal_one": label_dict["pos_equal_one"].to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()}
        P = self._param_dict()
        eng = self.engine
        heads, self._last_aux = eng.forward(P, lidar, layout, True, k_list)
        loss3, dheads = eng.loss(heads, labels, cls_weight, reg_coe)
        eng.backward(P, dheads, self._grad_buffers())
        return loss3
