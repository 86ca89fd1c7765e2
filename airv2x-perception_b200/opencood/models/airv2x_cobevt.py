"""`opencood.models.airv2x_cobevt.Airv2xCoBEVT` on the B200 kernels (BASELINE config 4).

Same registry name / class name / constructor (`cls(hypes["model"]["args"])`), same `hypes_yaml` keys
(`base_bev_backbone`, `shrink_header`, `compression`, `fax_fusion.*`, `max_cav`), same `state_dict` key names and shapes
(9 741 454 parameters for the shipped yaml), same `forward(data_dict) -> {"psm","rm","obj"}` as
opencood/models/airv2x_cobevt.py:15-156 of the reference. The torch.nn layers below are parameter containers only
(names + default init); their forward is never called. forward() in eval mode is the inference path; in train mode
(grad enabled) it is wired to autograd through the fused forward_train / backward_train of the engine, and train_step()
is the fused fast path (nn.Dropout as counter-based Philox masks); no CPU fallback.
"""
import torch
import torch.nn as nn

from ...cobevt_engine import CoBEVTEngine
from ...w2c_engine import AGENT_TYPES, TYPE_PREFIX
from .airv2x_where2com import Airv2xWhere2com, _backbone_params, _PillarVFEParams, _shrink_params


def _relative_position_index(L, w):
    """pair-wise relative (agent, row, col) offsets of the L*w*w window tokens flattened to one table index
    (cobevt_modules/swap_fusion_modules.py:53-76)"""
    d, h, ww = torch.meshgrid(torch.arange(L), torch.arange(w), torch.arange(w), indexing="ij")
    c = torch.stack([d.reshape(-1), h.reshape(-1), ww.reshape(-1)])
    rel = c[:, :, None] - c[:, None, :]
    s = 2 * w - 1
    return (rel[0] + L - 1) * s * s + (rel[1] + w - 1) * s + (rel[2] + w - 1)


class _AttentionParams(nn.Module):
    def __init__(self, dim, dim_head, drop, L, w):
        super().__init__()
        heads = dim // dim_head
        self.to_qkv = nn.Linear(dim, dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(dim, dim, bias=False), nn.Dropout(drop))
        self.relative_position_bias_table = nn.Embedding((2 * L - 1) * (2 * w - 1) * (2 * w - 1), heads)
        self.register_buffer("relative_position_index", _relative_position_index(L, w))


class _FeedForwardParams(nn.Module):
    def __init__(self, dim, hidden, drop):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Dropout(drop), nn.Linear(hidden, dim),
                                 nn.Dropout(drop))


class _PreNormResidualParams(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class _SwapFusionBlockParams(nn.Module):
    def __init__(self, dim, mlp_dim, dim_head, w, L, drop):
        super().__init__()
        self.window_attention = _PreNormResidualParams(dim, _AttentionParams(dim, dim_head, drop, L, w))
        self.window_ffd = _PreNormResidualParams(dim, _FeedForwardParams(dim, mlp_dim, drop))
        self.grid_attention = _PreNormResidualParams(dim, _AttentionParams(dim, dim_head, drop, L, w))
        self.grid_ffd = _PreNormResidualParams(dim, _FeedForwardParams(dim, mlp_dim, drop))


class _SwapFusionEncoderParams(nn.Module):
    def __init__(self, fa):
        super().__init__()
        if not fa.get("mask", False):
            raise NotImplementedError("fax_fusion.mask: false (SwapFusionBlock without agent mask) is not implemented")
        self.layers = nn.ModuleList([_SwapFusionBlockParams(fa["input_dim"], fa["mlp_dim"], fa["dim_head"],
                                                            fa["window_size"], fa["agent_size"], fa["drop_out"])
                                     for _ in range(fa["depth"])])
        self.mlp_head = nn.Sequential(nn.Identity(), nn.Identity(), nn.LayerNorm(fa["input_dim"]),
                                      nn.Linear(fa["input_dim"], fa["input_dim"]), nn.Identity())


def fusion_step(model, run_forward, names, params, out_shape):
    """Autograd boundary of the transformer-fusion models = the torch.library ops of torch_ops.py: inputs are the
    trainable parameters, the output is the NHWC head-logit tensor; the backward op runs the engine's backward_train and
    hands one gradient per parameter to autograd."""
    from ... import torch_ops

    def run_backward(dheads):
        P = model._param_dict()
        grads = {n: torch.zeros_like(P[n]) for n in names}
        model.engine.backward_train(P, dheads, grads)
        return [grads[n] for n in names]

    key = torch_ops.bind(model, run_forward, run_backward)
    return torch.ops.a2x.fused_forward(key, params, out_shape)


class Airv2xCoBEVT(Airv2xWhere2com):
    def __init__(self, args, precision="split3"):
        nn.Module.__init__(self)
        self.args = args
        self.collaborators = args["collaborators"]
        self.active_sensors = args["active_sensors"]
        self.max_cav_num = sum(args["max_cav"].values())
        self.veh_models, self.rsu_models, self.drone_models = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for t in AGENT_TYPES:
            if t not in self.collaborators:
                continue
            for m in args[t]["modalities"]:
                if m != "lidar":
                    raise NotImplementedError("modality %r is outside the B200 hot path (lidar only)" % m)
                getattr(self, TYPE_PREFIX[t]).append(nn.Sequential(_PillarVFEParams(args[t]["lidar"]["pillar_vfe"]),
                                                                   nn.Identity()))
        self.backbone = _backbone_params(args["base_bev_backbone"], 64)
        self.shrink_flag = bool(args.get("shrink_header", {}).get("use", False))
        if self.shrink_flag:
            self.shrink_conv = _shrink_params(args["shrink_header"])
        self.compression = args["compression"] > 0
        if self.compression:  # NaiveCompressor(256, ratio), naive_compress.py:5-42 (parameter containers only)
            c, r = 256, args["compression"]
            bn = lambda ch: nn.BatchNorm2d(ch, eps=1e-3, momentum=0.01)
            nc = nn.Module()
            nc.encoder = nn.Sequential(nn.Conv2d(c, c // r, 3, padding=1), bn(c // r), nn.ReLU())
            nc.decoder = nn.Sequential(nn.Conv2d(c // r, c, 3, padding=1), bn(c), nn.ReLU(), nn.Conv2d(c, c, 3, padding=1),
                                       bn(c), nn.ReLU())
            self.naive_compressor = nc
        args["fax_fusion"]["agent_size"] = self.max_cav_num  # airv2x_cobevt.py:49
        self.fusion_net = _SwapFusionEncoderParams(args["fax_fusion"])
        self.outC = args["outC"]
        if args["task"] != "det":
            raise NotImplementedError("task %r is outside the B200 hot path (det only)" % args["task"])
        self.cls_head = nn.Conv2d(self.outC, args["anchor_number"] * args["num_class"], kernel_size=1)
        self.reg_head = nn.Conv2d(self.outC, 7 * args["anchor_number"], kernel_size=1)
        if args["obj_head"]:
            self.obj_head = nn.Conv2d(self.outC, args["anchor_number"], kernel_size=1)
        self.precision = precision
        # nn.Dropout(fax_fusion.drop_out) in train mode: "on" = counter-based Philox masks (csrc/dropout.cu), seeded per step
        # from torch's default CPU generator (torch.manual_seed controls the run); "off" = disabled
        self.dropout = "on"
        self._engine = None
        self._last_aux = None

    @property
    def engine(self):
        if self._engine is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("Airv2xCoBEVT (B200) needs its parameters on a CUDA device; there is no CPU path")
            self._engine = CoBEVTEngine(self.args, dev, self.precision)
        return self._engine

    def forward(self, data_dict):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("Airv2xCoBEVT (B200) needs its parameters on a CUDA device; there is no CPU path")
        layout = self._layout(data_dict, dev)
        if "key_mask" not in layout:
            L = self.max_cav_num
            assert max(layout["record_len"]) <= L, "a scene has more agents than max_cav allows"
            layout["key_mask"] = torch.tensor([[1] * n + [0] * (L - n) for n in layout["record_len"]], dtype=torch.int32,
                                              device=dev)
        lidar = self._lidar(data_dict, dev, layout)
        if self.training and torch.is_grad_enabled():
            # reference training loop (tools/train.py:216-221): model(batch) -> criterion -> loss.backward()
            names = [n for n, p in self.named_parameters() if p.requires_grad]
            params = [p for n, p in self.named_parameters() if p.requires_grad]
            eng = self.engine
            drop = self._dropout_state(None)
            heads = fusion_step(self, lambda: eng.forward_train(self._param_dict(), lidar, layout, drop), names, params,
                                self._heads_shape(layout))
        else:
            heads, _ = self.engine.forward(self._param_dict(), lidar, layout, self.training)
        A, K = self.args["anchor_number"], self.args["num_class"]
        nc, nr = A * K, 7 * A
        nchw = heads.permute(0, 3, 1, 2)
        return {"psm": nchw[:, :nc], "rm": nchw[:, nc:nc + nr], "obj": nchw[:, nc + nr:nc + nr + A]}

    def _grad_buffers(self):
        g = {}
        for n, p in self.named_parameters():
            if not p.requires_grad:
                continue
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            g[n] = p.grad
        return g

    def _dropout_state(self, dropout):
        """dropout: None -> self.dropout; "on" (a fresh seed from torch's default CPU generator), "off", an int seed, or a
        ready ops.Dropout. Returns ops.Dropout or None."""
        from ... import ops
        p = float(self.args["fax_fusion"].get("drop_out", 0.0))
        mode = self.dropout if dropout is None else dropout
        if isinstance(mode, ops.Dropout):
            return mode
        if p <= 0.0 or mode == "off":
            return None
        if mode == "on":
            return ops.Dropout(p, int(torch.randint(0, 2 ** 62, (1,)).item()))
        if isinstance(mode, int):
            return ops.Dropout(p, mode)
        raise ValueError("dropout must be 'on', 'off', an int seed or an ops.Dropout, got %r" % (mode,))

    def train_step(self, data_dict, label_dict, cls_weight=1.0, reg_coe=2.0, dropout=None):
        """forward (train-mode BatchNorm) + PointPillarLossMultiClass + backward of the whole CoBEVT path on the CUDA
        kernels; parameter gradients land in p.grad, returns the device tensor [reg, cls, obj] (float64).
        nn.Dropout(fax_fusion.drop_out) of the fusion network runs as counter-based masks (see _dropout_state); the state
        used is kept in self.last_dropout (tests export its masks to the oracle)."""
        assert self.training, "train_step() needs model.train()"
        drop = self.last_dropout = self._dropout_state(dropout)
        dev = next(self.parameters()).device
        layout = self._layout(data_dict, dev)
        if "key_mask" not in layout:
            L = self.max_cav_num
            layout["key_mask"] = torch.tensor([[1] * n + [0] * (L - n) for n in layout["record_len"]], dtype=torch.int32,
                                              device=dev)
        lidar = self._lidar(data_dict, dev, layout)
        labels = self.prepare_labels(label_dict, dev)
        P = self._param_dict()
        eng = self.engine
        heads = eng.forward_train(P, lidar, layout, drop)
        loss3, dheads = eng.loss(heads, labels, cls_weight, reg_coe)
        eng.backward_train(P, dheads, self._grad_buffers())
        return loss3

    def train_step_graphed(self, *a, **k):
        raise NotImplementedError("Airv2xCoBEVT: CUDA-graph replay of the training step is not implemented")
