"""Host-side orchestration of the legacy transformer-fusion models on the C-ABI kernels:
`point_pillar_cobevt` (models/point_pillar_cobevt.py:14-128) and `point_pillar_v2xvit` (models/point_pillar_v2xvit.py:14-185).

Same kernels as the airv2x engines; what differs from them: ONE PillarVFE for all agents (`pillar_vfe.*`), the legacy
3x3 stride-2 shrink header, `max_cav` an int, 1-class heads without objectness, an optional NaiveCompressor; the V2X-ViT
variant warps every agent's map into the ego frame with `pairwise_t_matrix[b, 0]` (warp_affine_simple, align_corners
False) before the transformer and feeds it a zero prior encoding and identity spatial correction. Eval forward and the
training step (forward_train / loss = PointPillarLoss / backward_train, inherited from the airv2x engines: the legacy
shrink header, head rows and pre-warp are parameters of those).
"""
import torch

from . import ops, warp
from .cobevt_engine import CoBEVTEngine
from .ops import Act
from .ppw2c_engine import LegacyW2CEngine
from .v2xvit_engine import V2XViTEngine
from .w2c_engine import HEAD_PAD


def _legacy_common(self, args, device, precision):
    assert precision in ("split3", "tf32"), precision
    self.args = args
    self.device = torch.device(device)
    self.split = precision == "split3"
    self.precision = precision
    bb = args["base_bev_backbone"]
    self.layer_nums = list(bb["layer_nums"])
    self.layer_strides = list(bb["layer_strides"])
    self.num_filters = list(bb["num_filters"])
    self.up_strides = list(bb["upsample_strides"])
    self.up_filters = list(bb["num_upsample_filter"])
    assert all(s == 2 for s in self.layer_strides), "backbone blocks must have stride 2"
    sh = args["shrink_header"]
    assert list(sh["kernal_size"]) == [3] and list(sh["stride"]) in ([2], [1]) and list(sh["padding"]) == [1], \
        "legacy shrink header: one 3x3 (stride 1 or 2, padding 1) + 3x3 double conv"
    self.shrink_k0 = 3
    self.shrink_stride = int(sh["stride"][0])
    self.compression = int(args.get("compression", 0) or 0)
    self.c_cat = sum(self.up_filters)
    self.c_shrink = sh["dim"][0]
    assert sh["input_dim"] == self.c_cat and self.c_shrink == 256, "heads are Conv2d(128 * 2, ...) in the reference"
    assert self.compression == 0 or (256 % self.compression == 0 and (256 // self.compression) % 64 == 0), \
        "NaiveCompressor: 256 / compression must be a multiple of 64 channels (compression in {1, 2, 4})"
    self.A = args["anchor_number"]
    self.K = 1
    self.n_head = self.A + 7 * self.A
    self.L = int(args["max_cav"])
    self.bufs = {}
    self.saved = None
    self.side = None
    self.use_side_stream = False
    self.legacy_loss = True     # PointPillarLoss (loss/point_pillar_loss.py:77-215), not the multi-class loss


class LegacyCoBEVTEngine(CoBEVTEngine):
    def __init__(self, args, device, precision="split3"):  # noqa
        _legacy_common(self, args, device, precision)
        fa = args["fax_fusion"]
        self.fa = fa
        self.dim = fa["input_dim"]
        assert self.dim == self.c_shrink and self.dim % fa["dim_head"] == 0 and fa["agent_size"] == self.L
        self.heads = self.dim // fa["dim_head"]

    _encode = LegacyW2CEngine._encode

    def _head_rows(self):
        return (("cls_head", 0), ("reg_head", self.A))

    def forward(self, P, lidar, layout, training, k_list=None):
        if training:
            raise NotImplementedError("PointPillarCoBEVT: train-mode forward under torch.no_grad() is not implemented "
                                      "(model.eval() for inference; model(batch) with grad enabled or train_step() for training)")
        self._begin_step()
        W = self._pack_weights(P)
        feat = self.encode(P, W, lidar, layout)
        nz = self._buf("comm_rate", (1,), torch.int64)
        nz.copy_(self._canvas_nz)
        return self.fuse_heads(P, W, feat, layout), {"comm_rate": nz}


class LegacyV2XViTEngine(V2XViTEngine):
    def __init__(self, args, device, precision="split3"):  # noqa
        _legacy_common(self, args, device, precision)
        enc = args["transformer"]["encoder"]
        self.enc = enc
        ca, pw = enc["cav_att_config"], enc["pwindow_att_config"]
        assert ca["use_hetero"], "CavAttention (use_hetero: false) is not implemented"
        assert pw["fusion_method"] == "split_attn" and len(pw["window_size"]) == 3 and pw["relative_pos_embedding"], \
            "only the 3-branch split_attn pyramid with relative position embedding is implemented"
        assert enc["num_blocks"] == 1, "num_blocks > 1 not implemented"
        self.dim = ca["dim"]
        assert self.dim == self.c_shrink and ca["heads"] * ca["dim_head"] == self.dim
        assert all(h * d == self.dim for h, d in zip(pw["heads"], pw["dim_head"]))
        self.discrete_ratio = args["voxel_size"][0]

    _encode = LegacyW2CEngine._encode

    def _head_rows(self):
        return (("cls_head", 0), ("reg_head", self.A))

    def _ego_theta(self, pairwise, layout):
        """warp_affine_simple(regroup_feature[b], pairwise_t_matrix[b, ego = 0]) (point_pillar_v2xvit.py:140-166). H, W of
        the normalisation are the pillar canvas', downsample_rate is 1 there (:118-119)"""
        record_len = layout["record_len"]
        theta_all = warp.normalize_pairwise(pairwise, layout["ny"], layout["nx"], 1, self.discrete_ratio)   # [B, L, L, 2, 3]
        return torch.stack([theta_all[b, 0, l] for b in range(len(record_len)) for l in range(record_len[b])]).to(self.device)

    def forward_train(self, P, lidar, layout, pairwise, drops=None):
        B = len(layout["record_len"])
        prior = torch.zeros(B, self.L, 3)
        scm = torch.eye(4, dtype=torch.float64).repeat(B, self.L, 1, 1)
        return super().forward_train(P, lidar, layout, prior, scm, drops, pre_warp=self._ego_theta(pairwise, layout))

    def forward(self, P, lidar, layout, training, pairwise=None):
        if training:
            raise NotImplementedError("PointPillarV2XVit: train-mode forward under torch.no_grad() is not implemented "
                                      "(model.eval() for inference; model(batch) with grad enabled or train_step() for training)")
        self._begin_step()
        W = self._pack_weights(P)
        feat = self.encode(P, W, lidar, layout)
        nz = self._buf("comm_rate", (1,), torch.int64)
        nz.copy_(self._canvas_nz)
        record_len = layout["record_len"]
        B, (N, h, w, C) = len(record_len), feat.shape
        theta = self._ego_theta(pairwise, layout)
        warped = self._buf("pp.warped", feat.shape)
        ops.warp_affine_fwd(feat, theta, Act(warped), align_corners=False)
        prior = torch.zeros(B, self.L, 3)
        scm = torch.eye(4, dtype=torch.float64).repeat(B, self.L, 1, 1)
        fused = self.fusion(P, W, warped, layout, prior, scm)
        heads = self._buf("heads.out", (B, h, w, HEAD_PAD))
        ops.linear_fwd(fused, W["heads"], Act(heads), bias=W["heads.bias"])
        return heads, {"comm_rate": nz}
