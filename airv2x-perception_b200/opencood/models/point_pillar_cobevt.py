"""`opencood.models.point_pillar_cobevt.PointPillarCoBEVT` on the B200 kernels.

Same registry name / class name / constructor, same `hypes_yaml` keys (`pillar_vfe`, `point_pillar_scatter`,
`base_bev_backbone`, `shrink_header`, `compression`, `fax_fusion`, `max_cav`, `anchor_number`), same `state_dict` keys
and shapes (10 513 344 parameters for V2XR_cobevt.yaml), same
`forward(data_dict) -> {"psm","rm","mask","each_mask","comm_rate"}` as opencood/models/point_pillar_cobevt.py:14-128 of
the reference (input: `data_dict["processed_lidar"]`, `record_len`). Parameter containers only; eval forward, the
reference-style training loop (`model(batch)` with grad enabled -> PointPillarLoss -> `loss.backward()`) and the fused
`train_step()`; no CPU fallback.
"""
import numpy as np
import torch
import torch.nn as nn

from ...pplegacy_engine import LegacyCoBEVTEngine
from ...w2c_engine import HEAD_PAD
from .airv2x_cobevt import Airv2xCoBEVT, _SwapFusionEncoderParams, fusion_step
from .airv2x_where2com import _backbone_params, _PillarVFEParams
from .point_pillar_where2comm import _legacy_shrink_params


def _compressor_params(c, r):
    """NaiveCompressor(256, r), common_modules/naive_compress.py:5-36"""
    bn = lambda ch: nn.BatchNorm2d(ch, eps=1e-3, momentum=0.01)
    nc = nn.Module()
    nc.encoder = nn.Sequential(nn.Conv2d(c, c // r, 3, padding=1), bn(c // r), nn.ReLU())
    nc.decoder = nn.Sequential(nn.Conv2d(c // r, c, 3, padding=1), bn(c), nn.ReLU(), nn.Conv2d(c, c, 3, padding=1),
                               bn(c), nn.ReLU())
    return nc


class _LegacyFusionModel(nn.Module):
    """what point_pillar_cobevt / point_pillar_v2xvit share: the legacy encoder's parameters and the input plumbing"""
    ENGINE = None

    def _init_encoder(self, args, precision):
        self.args = args
        self.modality = args.get("use_modality", "processed_lidar")
        self.max_cav = args["max_cav"]
        self.pillar_vfe = _PillarVFEParams(args["pillar_vfe"])
        self.backbone = _backbone_params(args["base_bev_backbone"], 64)
        self.shrink_flag = "shrink_header" in args
        if not self.shrink_flag:
            raise NotImplementedError("%s without a shrink header is not implemented" % type(self).__name__)
        self.shrink_conv = _legacy_shrink_params(args["shrink_header"])
        self.compression = args["compression"] > 0
        if self.compression:
            self.naive_compressor = _compressor_params(256, args["compression"])
        self.precision = precision
        self._engine = None

    def _init_heads(self, args):
        self.cls_head = nn.Conv2d(128 * 2, args["anchor_number"], kernel_size=1)
        self.reg_head = nn.Conv2d(128 * 2, 7 * args["anchor_number"], kernel_size=1)
        if args["backbone_fix"]:
            self.backbone_fix()

    def backbone_fix(self):
        for n, p in self.named_parameters():
            if not n.startswith("fusion_net"):
                p.requires_grad = False

    @property
    def engine(self):
        if self._engine is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("%s (B200) needs its parameters on a CUDA device; there is no CPU path" % type(self).__name__)
            self._engine = self.ENGINE(self.args, dev, self.precision)
        return self._engine

    def _param_dict(self):
        d = {n: p.data for n, p in self.named_parameters()}
        d.update({n: b for n, b in self.named_buffers()})
        return d

    def _inputs(self, data_dict):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("%s (B200) needs its parameters on a CUDA device; there is no CPU path" % type(self).__name__)
        lid = data_dict[self.modality]
        r = data_dict["record_len"]
        record_len = [int(v) for v in (r.tolist() if torch.is_tensor(r) else r)]
        assert max(record_len) <= self.max_cav, "a scene has more agents than max_cav allows"
        key = tuple(record_len)
        cache = self.__dict__.setdefault("_layout_cache", {})
        if key not in cache:
            nx, ny, _ = [int(v) for v in self.args["point_pillar_scatter"]["grid_size"]]
            n = sum(record_len)
            start = np.concatenate([[0], np.cumsum(record_len)[:-1]]).astype(np.int32)
            L = self.max_cav
            cache[key] = dict(n_total=n, nx=nx, ny=ny, record_len=record_len,
                              identity_map=torch.arange(n, dtype=torch.int32, device=dev),
                              scene_start=torch.tensor(start, dtype=torch.int32, device=dev),
                              scene_len=torch.tensor(record_len, dtype=torch.int32, device=dev),
                              key_mask=torch.tensor([[1] * k + [0] * (L - k) for k in record_len], dtype=torch.int32, device=dev))
        lidar = {"voxel_features": lid["voxel_features"].to(device=dev, dtype=torch.float32).contiguous(),
                 "voxel_num_points": lid["voxel_num_points"].to(device=dev, dtype=torch.int32).contiguous(),
                 "voxel_coords": lid["voxel_coords"].to(device=dev, dtype=torch.int32).contiguous()}
        return lidar, cache[key]

    # ------------------------------------------------------------------ training
    dropout = "on"   # nn.Dropout of the fusion network in train mode: counter-based masks ("off" disables; see _dropout_state)

    def _heads_shape(self, layout):
        s = int(self.args["shrink_header"]["stride"][0])
        h, w = layout["ny"] // 2, layout["nx"] // 2
        return [len(layout["record_len"]), (h - 1) // s + 1, (w - 1) // s + 1, HEAD_PAD]

    def _grad_buffers(self):
        g = {}
        for n, p in self.named_parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            g[n] = p.grad
        return g

    def _check_trainable(self):
        if self.args.get("backbone_fix", False):
            raise NotImplementedError("%s: training with backbone_fix (frozen encoder) is not implemented" % type(self).__name__)

    def _forward_train(self, P, lidar, layout, data_dict, drop):
        raise NotImplementedError

    def _train_forward_autograd(self, data_dict):
        """reference training loop (tools/train.py:216-221): model(batch) -> criterion -> loss.backward(); the autograd
        boundary is the torch.library op pair of torch_ops.py"""
        self._check_trainable()
        lidar, layout = self._inputs(data_dict)
        names = [n for n, p in self.named_parameters() if p.requires_grad]
        params = [p for n, p in self.named_parameters() if p.requires_grad]
        drop = self._dropout_state(None)
        heads = fusion_step(self, lambda: self._forward_train(self._param_dict(), lidar, layout, data_dict, drop), names, params,
                            self._heads_shape(layout))
        return self._outputs(heads, {"comm_rate": self.engine._canvas_nz})

    def train_step(self, data_dict, label_dict, cls_weight=1.0, reg_coe=2.0, dropout=None):
        """forward (train-mode BatchNorm) + PointPillarLoss (loss/point_pillar_loss.py:77-215) + backward in one call on the
        CUDA kernels; label_dict = the legacy collate's {"targets" [B,H,W,7A], "pos_equal_one" [B,H,W,A]}. Parameter
        gradients land in p.grad; returns the device tensor [reg, conf, 0] (float64), total = .sum(). nn.Dropout of the
        fusion network runs as counter-based masks (dropout: None -> self.dropout; "on", "off" or an int seed)."""
        assert self.training, "train_step() needs model.train()"
        self._check_trainable()
        drop = self.last_dropout = self._dropout_state(dropout)
        lidar, layout = self._inputs(data_dict)
        dev = next(self.parameters()).device
        labels = {"targets": label_dict["targets"].to(device=dev, dtype=torch.float32, non_blocking=True).contiguous(),
                  "pos_equal_one": label_dict["pos_equal_one"].to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()}
        P = self._param_dict()
        eng = self.engine
        heads = self._forward_train(P, lidar, layout, data_dict, drop)
        loss3, dheads = eng.loss(heads, labels, cls_weight, reg_coe)
        eng.backward_train(P, dheads, self._grad_buffers())
        return loss3

    def _outputs(self, heads, aux):
        A = self.args["anchor_number"]
        nchw = heads.permute(0, 3, 1, 2)
        return {"psm": nchw[:, :A], "rm": nchw[:, A:8 * A], "mask": 0, "each_mask": 0,
                "comm_rate": int(aux["comm_rate"].item())}


class PointPillarCoBEVT(_LegacyFusionModel):
    ENGINE = LegacyCoBEVTEngine

    def __init__(self, args, precision="split3"):
        super().__init__()
        self._init_encoder(args, precision)
        self.fusion_net = _SwapFusionEncoderParams(args["fax_fusion"])
        self._init_heads(args)

    _dropout_state = Airv2xCoBEVT._dropout_state     # fax_fusion.drop_out, same keys as the airv2x yaml

    def _forward_train(self, P, lidar, layout, data_dict, drop):
        return self.engine.forward_train(P, lidar, layout, drop)

    def forward(self, data_dict):
        if self.training and torch.is_grad_enabled():
            return self._train_forward_autograd(data_dict)
        lidar, layout = self._inputs(data_dict)
        heads, aux = self.engine.forward(self._param_dict(), lidar, layout, self.training)
        return self._outputs(heads, aux)
