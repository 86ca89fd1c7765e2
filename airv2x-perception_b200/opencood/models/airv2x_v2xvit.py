"""`opencood.models.airv2x_v2xvit.Airv2xV2XVit` on the B200 kernels (BASELINE config 3).

Same registry name / class name / constructor (`cls(hypes["model"]["args"])`), same `hypes_yaml` keys
(`modality_fusion.*`, `transformer.encoder.*`, `max_cav`), same `state_dict` key names and shapes (12 758 879 parameters
for the shipped yaml, including the unused `prior_feed` and the frozen sinusoid table of the RTE embedding), same
`forward(data_dict) -> {"psm","rm","obj","comm_rate"}` as opencood/models/airv2x_v2xvit.py:19-167 of the reference.
The torch.nn layers below are parameter containers only; their forward is never called. forward() is the eval path,
train_step() the fused training step (nn.Dropout as counter-based Philox masks); no CPU fallback.
"""
import math

import torch
import torch.nn as nn

from ...v2xvit_engine import V2XViTEngine
from ...w2c_engine import AGENT_TYPES, TYPE_PREFIX
from .airv2x_cobevt import fusion_step
from .airv2x_where2com import Airv2xWhere2com, _backbone_params, _PillarVFEParams, _shrink_params


class _HGTParams(nn.Module):
    """hmsa.py:6-35 (parameter order matters for state_dict order only)"""

    def __init__(self, dim, heads, dim_head, num_types=2, num_relations=4):
        super().__init__()
        inner = heads * dim_head
        self.k_linears, self.q_linears = nn.ModuleList(), nn.ModuleList()
        self.v_linears, self.a_linears = nn.ModuleList(), nn.ModuleList()
        for _ in range(num_types):
            self.k_linears.append(nn.Linear(dim, inner))
            self.q_linears.append(nn.Linear(dim, inner))
            self.v_linears.append(nn.Linear(dim, inner))
            self.a_linears.append(nn.Linear(inner, dim))
        self.relation_att = nn.Parameter(torch.empty(num_relations, heads, dim_head, dim_head))
        self.relation_msg = nn.Parameter(torch.empty(num_relations, heads, dim_head, dim_head))
        nn.init.xavier_uniform_(self.relation_att)
        nn.init.xavier_uniform_(self.relation_msg)


class _WindowAttnParams(nn.Module):
    def __init__(self, dim, heads, dim_head, ws):
        super().__init__()
        inner = heads * dim_head
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.pos_embedding = nn.Parameter(torch.randn(2 * ws - 1, 2 * ws - 1))
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Identity())


class _SplitAttnParams(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, dim, bias=False)
        self.bn1 = nn.LayerNorm(dim)
        self.fc2 = nn.Linear(dim, dim * 3, bias=False)


class _PyramidParams(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.pwmsa = nn.ModuleList([_WindowAttnParams(cfg["dim"], h, d, ws)
                                    for h, d, ws in zip(cfg["heads"], cfg["dim_head"], cfg["window_size"])])
        if cfg["fusion_method"] == "split_attn":
            self.split_attn = _SplitAttnParams(256)


class _PreNormParams(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class _FeedForwardParams(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Identity(), nn.Linear(hidden, dim), nn.Identity())


class _FusionBlockParams(nn.Module):
    def __init__(self, num_blocks, ca, pw):
        super().__init__()
        self.layers = nn.ModuleList([nn.ModuleList([
            _PreNormParams(ca["dim"], _HGTParams(ca["dim"], ca["heads"], ca["dim_head"])),
            _PreNormParams(ca["dim"], _PyramidParams(pw))]) for _ in range(num_blocks)])


class _RTEParams(nn.Module):
    def __init__(self, dim, max_len=100):
        super().__init__()
        emb = nn.Module()
        emb.emb = nn.Embedding(max_len, dim)
        position = torch.arange(0.0, max_len).unsqueeze(1)
        div = torch.exp(torch.arange(0, dim, 2) * -(math.log(10000.0) / dim))
        emb.emb.weight.data[:, 0::2] = torch.sin(position * div) / math.sqrt(dim)
        emb.emb.weight.data[:, 1::2] = torch.cos(position * div) / math.sqrt(dim)
        emb.lin = nn.Linear(dim, dim)
        self.emb = emb


class _EncoderParams(nn.Module):
    def __init__(self, enc):
        super().__init__()
        ca, pw = enc["cav_att_config"], enc["pwindow_att_config"]
        self.prior_feed = nn.Linear(ca["dim"] + 3, ca["dim"])   # present in the reference, never used (v2xvit_basic.py:150)
        self.layers = nn.ModuleList([])
        if ca["use_RTE"]:
            self.rte = _RTEParams(ca["dim"])
        for _ in range(enc["depth"]):
            self.layers.append(nn.ModuleList([_FusionBlockParams(enc["num_blocks"], ca, pw),
                                              _PreNormParams(ca["dim"], _FeedForwardParams(ca["dim"], enc["feed_forward"]["mlp_dim"]))]))


class _TransformerParams(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.encoder = _EncoderParams(args["encoder"])


class Airv2xV2XVit(Airv2xWhere2com):
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
        mf = args["modality_fusion"]
        self.backbone = _backbone_params(mf["base_bev_backbone"], 64)
        self.shrink_flag = bool(mf.get("shrink_header", {}).get("use", False))
        if self.shrink_flag:
            self.shrink_conv = _shrink_params(mf["shrink_header"])
        self.compression = mf["compression"] > 0
        self.fusion_net = _TransformerParams(args["transformer"])
        self.outC = args["outC"]
        if args["task"] != "det":
            raise NotImplementedError("task %r is outside the B200 hot path (det only)" % args["task"])
        self.cls_head = nn.Conv2d(self.outC, args["anchor_number"] * args["num_class"], kernel_size=1)
        self.reg_head = nn.Conv2d(self.outC, 7 * args["anchor_number"], kernel_size=1)
        if args["obj_head"]:
            self.obj_head = nn.Conv2d(self.outC, args["anchor_number"], kernel_size=1)
        self.precision = precision
        # nn.Dropout (cav_att / pwindow_att / feed_forward `dropout`) in train mode: "on" = counter-based Philox masks
        # (csrc/dropout.cu), seeded per step from torch's default CPU generator; "off" = disabled
        self.dropout = "on"
        self._engine = None
        self._last_aux = None

    @property
    def engine(self):
        if self._engine is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("Airv2xV2XVit (B200) needs its parameters on a CUDA device; there is no CPU path")
            self._engine = V2XViTEngine(self.args, dev, self.precision)
        return self._engine

    def forward(self, data_dict):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("Airv2xV2XVit (B200) needs its parameters on a CUDA device; there is no CPU path")
        layout = self._layout(data_dict, dev)
        lidar = self._lidar(data_dict, dev, layout)
        prior, scm = data_dict["prior_encoding"], data_dict["spatial_correction_matrix"]
        eng = self.engine
        if self.training and torch.is_grad_enabled():
            # reference training loop (tools/train.py:216-221): model(batch) -> criterion -> loss.backward()
            names = [n for n, p in self.named_parameters() if p.requires_grad]
            params = [p for n, p in self.named_parameters() if p.requires_grad]
            drops = self._dropout_state(None)
            heads = fusion_step(self, lambda: eng.forward_train(self._param_dict(), lidar, layout, prior, scm, drops),
                                names, params, self._heads_shape(layout))
            aux = eng.last_aux
        else:
            heads, aux = eng.forward(self._param_dict(), lidar, layout, self.training, prior=prior, scm=scm)
        A, K = self.args["anchor_number"], self.args["num_class"]
        nc, nr = A * K, 7 * A
        nchw = heads.permute(0, 3, 1, 2)
        return {"psm": nchw[:, :nc], "rm": nchw[:, nc:nc + nr], "obj": nchw[:, nc + nr:nc + nr + A],
                "comm_rate": int(aux["comm_rate"].item())}

    def _grad_buffers(self):
        g = {}
        for n, p in self.named_parameters():
            if not p.requires_grad:
                continue
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            g[n] = p.grad
        return g

    def _dropouts(self):
        enc = self.args["transformer"]["encoder"]
        return [float(enc["cav_att_config"].get("dropout", 0.0)), float(enc["pwindow_att_config"].get("dropout", 0.0)),
                float(enc["feed_forward"].get("dropout", 0.0))]

    def _dropout_state(self, dropout):
        """dropout: None -> self.dropout; "on" (fresh seed from torch's default CPU generator), "off", or an int seed.
        Returns None or the (cav, pwindow, ffn) ops.Dropout states: one seed, disjoint site-id ranges, own rates."""
        from ... import ops
        mode = self.dropout if dropout is None else dropout
        ps = self._dropouts()
        if isinstance(mode, tuple):
            return mode
        if max(ps) <= 0.0 or mode == "off":
            return None
        if mode == "on":
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        elif isinstance(mode, int):
            seed = mode
        else:
            raise ValueError("dropout must be 'on', 'off' or an int seed, got %r" % (mode,))
        out = []
        for k, p in enumerate(ps):
            d = ops.Dropout(p, seed) if p > 0 else None
            if d is not None:
                d.n_sites = 1000 * k
            out.append(d)
        return tuple(out)

    def train_step(self, data_dict, label_dict, cls_weight=1.0, reg_coe=2.0, dropout=None):
        """forward (train-mode BatchNorm) + PointPillarLossMultiClass + backward of the whole V2X-ViT path on the CUDA
        kernels; parameter gradients land in p.grad, returns the device tensor [reg, cls, obj] (float64).
        nn.Dropout (cav_att / pwindow_att / feed_forward `dropout`) runs as counter-based masks (see _dropout_state; the
        states used are kept in self.last_dropout). The RTE embedding table receives a gradient as it does in the reference
        (`emb.requires_grad = False` at v2xvit_basic.py:53 sets a module attribute, the weight stays trainable)."""
        assert self.training, "train_step() needs model.train()"
        drops = self.last_dropout = self._dropout_state(dropout)
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("Airv2xV2XVit (B200) needs its parameters on a CUDA device; there is no CPU path")
        layout = self._layout(data_dict, dev)
        lidar = self._lidar(data_dict, dev, layout)
        labels = self.prepare_labels(label_dict, dev)
        P = self._param_dict()
        eng = self.engine
        heads = eng.forward_train(P, lidar, layout, data_dict["prior_encoding"], data_dict["spatial_correction_matrix"], drops)
        loss3, dheads = eng.loss(heads, labels, cls_weight, reg_coe)
        eng.backward_train(P, dheads, self._grad_buffers())
        return loss3

    def train_step_graphed(self, *a, **k):
        raise NotImplementedError("Airv2xV2XVit: CUDA-graph replay of the training step is not implemented")
