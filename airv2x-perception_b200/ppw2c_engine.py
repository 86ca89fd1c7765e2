"""Host-side orchestration of the legacy `point_pillar_where2comm` path (BASELINE config 1) on the C-ABI kernels.

Same kernels as the airv2x Where2comm engine; what differs (models/point_pillar_where2comm.py:26-151): ONE PillarVFE
for all agents (`pillar_vfe.*`), the backbone evaluated once before the fusion, a 3x3 stride-2 shrink header — so the
communication mask lives at half the resolution of the level-0 features and is bilinearly resized
(where2comm_fuse.py:230-236) — and 1-class heads without objectness. Eval-mode forward in this round.
"""
import torch

from . import ops
from .ops import Act
from .w2c_engine import HEAD_PAD, W2CEngine


class LegacyW2CEngine(W2CEngine):
    def __init__(self, args, device, precision="split3"):  # noqa
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
        self.shrink_stride = int(sh["stride"][0])
        assert not args.get("compression", 0), "NaiveCompressor (compression > 0) is not implemented for this model"
        self.c_cat = sum(self.up_filters)
        self.c_shrink = sh["dim"][0]
        assert sh["input_dim"] == self.c_cat and args["head_dim"] == self.c_shrink
        self.A = args["anchor_number"]
        self.K = 1
        self.n_head = self.A + 7 * self.A
        fa = args["where2comm_fusion"]
        assert fa["multi_scale"], "single-scale Where2comm not implemented"
        self.fully = bool(fa["fully"])
        self.comm = fa["communication"]
        self.bufs = {}
        self.saved = None
        self.side = None
        self.use_side_stream = True
        self.fuse_bn_bwd_reduce = False

    def _pack_weights(self, P):
        W, jobs = {}, []
        for i, ln in enumerate(self.layer_nums):
            for k in range(ln + 1):
                name = "backbone.blocks.%d.%d.weight" % (i, 1 + 3 * k)
                w = P[name]
                co, ci = w.shape[0], w.shape[1]
                W[name] = self._packed(name, (9, co, ci), (9, ci, co))
                jobs.append(ops.conv_pack_job(w, W[name]))
            name = "backbone.deblocks.%d.0.weight" % i
            w = P[name]
            s = self.up_strides[i]
            ci, co = w.shape[0], w.shape[1]
            W[name] = self._packed(name, (1, s * s * co, ci), (s * s, ci, co))
            jobs.append(ops.deconv_pack_job(w, W[name]))
        for idx in (0, 2):
            name = "shrink_conv.layers.0.double_conv.%d.weight" % idx
            w = P[name]
            co, ci = w.shape[0], w.shape[1]
            W[name] = self._packed(name, (9, co, ci), (9, ci, co))
            jobs.append(ops.conv_pack_job(w, W[name]))
        fresh = ("packed", "heads") not in self.bufs
        hp = self._packed("heads", (1, HEAD_PAD, self.c_shrink), (1, self.c_shrink, HEAD_PAD))
        hb = self._buf("heads.b", (HEAD_PAD,))
        if fresh:
            for t in (hp.f32, hp.f16, hp.d32, hp.d16, hb):
                t.zero_()
        for name, row0 in (("cls_head", 0), ("reg_head", self.A)):
            jobs.append(ops.conv_pack_job(P[name + ".weight"], hp, row0))
            jobs.append(ops.copy_pack_job(P[name + ".bias"], hb, row0))
        W["heads"] = hp
        W["heads.bias"] = hb
        ops.pack_weights_batched(self._job_table("pack", jobs))
        return W

    def _encode(self, P, lidar, layout, training, record):
        n_total, ny, nx = layout["n_total"], layout["ny"], layout["nx"]
        canvas = self._act("canvas", (n_total, ny, nx, 64))
        hi = canvas.b16 is None      # split mode: only the bf16 planes feed the block-0 GEMM (see W2CEngine._encode)
        if hi:
            canvas.hi.zero_()
        else:
            canvas.b16.zero_()
        nzc = self._buf("canvas.nz", (1,), torch.int64)
        nzc.zero_()
        self._canvas_nz = nzc
        geom = ops.pfn_geom(self.args["voxel_size"], self.args["lidar_range"], nx, ny)
        pre = "pillar_vfe.pfn_layers.0"
        scale, shift = self._buf("pfn.scale", (64,)), self._buf("pfn.shift", (64,))
        ops.bn_eval_affine(P[pre + ".norm.weight"], P[pre + ".norm.bias"], P[pre + ".norm.running_mean"],
                           P[pre + ".norm.running_var"], scale, shift)
        ops.pfn_scatter(lidar["voxel_features"], lidar["voxel_num_points"], lidar["voxel_coords"], geom,
                        P[pre + ".linear.weight"], scale, shift, layout["identity_map"], canvas, nz=nzc, write_hi=hi)
        return canvas

    def _shrink_heads(self, P, W, cat, tag, heads_only_cls=False):
        n, h, w, _ = cat.shape
        s = self.shrink_stride
        ho, wo = (h - 1) // s + 1, (w - 1) // s + 1
        y1 = self._act(tag + ".s1", (n, ho, wo, self.c_shrink))
        y2 = self._act(tag + ".s2", (n, ho, wo, self.c_shrink))
        ops.conv_fwd(cat, W["shrink_conv.layers.0.double_conv.0.weight"], 3, s, y1,
                     shift=P["shrink_conv.layers.0.double_conv.0.bias"], relu=True)
        ops.conv_fwd(y1, W["shrink_conv.layers.0.double_conv.2.weight"], 3, 1, y2,
                     shift=P["shrink_conv.layers.0.double_conv.2.bias"], relu=True)
        heads = self._buf(tag + ".heads", (n, ho, wo, HEAD_PAD))
        ops.conv_fwd(y2, W["heads"], 1, 1, Act(heads), shift=W["heads.bias"])
        return y1, y2, heads

    def forward(self, P, lidar, layout, training, k_list=None):
        if training:
            raise NotImplementedError("point_pillar_where2comm on the B200 kernels is eval-only in this round")
        self._begin_step()
        W = self._pack_weights(P)
        record_len = layout["record_len"]
        B, N = len(record_len), layout["n_total"]
        canvas = self._encode(P, lidar, layout, False, None)
        nz = self._buf("comm_rate", (1,), torch.int64)
        nz.copy_(self._canvas_nz)
        x0 = self._block(P, W, 0, canvas, False, 0, "A", None)
        h2, w2 = x0.shape[1], x0.shape[2]
        catA = self._act("A.cat", (N, h2, w2, self.c_cat))
        xa = x0
        for i in range(len(self.layer_nums)):
            if i > 0:
                xa = self._block(P, W, i, xa, False, 0, "A", None)
            c0 = sum(self.up_filters[:i])
            with self._on_side():
                self._deblock(P, W, i, xa, catA.slice_c(c0, c0 + self.up_filters[i]), False, 0, "A", None)
        self._join_side()
        _, _, headsA = self._shrink_heads(P, W, catA, "A")
        hm, wm = headsA.shape[1], headsA.shape[2]
        mask_lo = self._buf("mask.lo", (N, hm, wm))
        ones = self._buf("mask.ones", (B,))
        thr = float(self.comm["threshold"])
        if self.fully:
            mask_lo.fill_(1.0)
            ones.fill_(float("nan"))
        else:
            conf, smooth = self._buf("conf", (N, hm, wm)), self._buf("smooth", (N, hm, wm))
            ops.comm_confidence(headsA, self.A * self.K, conf)
            gs = self.comm.get("gaussian_smooth")
            if thr:
                ops.comm_smooth_mask(conf, P.get("fusion_net.naive_communication.gaussian_filter.weight"),
                                     P.get("fusion_net.naive_communication.gaussian_filter.bias"),
                                     gs["k_size"] if gs else 0, N, hm, wm, thr, True, smooth, mask_lo)
            else:
                mask_lo.fill_(1.0)
            ones.zero_()
            ops.comm_rate_ego(mask_lo, hm * wm, B, layout["scene_start"], layout["scene_len"], ones)
        if (hm, wm) != (h2, w2):
            mask = self._buf("mask", (N, h2, w2))
            ops.resize_bilinear(mask_lo, mask)
        else:
            mask = mask_lo
        x0m = self._act("B.x0m", x0.shape)
        ops.affine_act(x0.hi, None, None, False, x0m, mask=mask)
        catB = self._act("B.cat", (B, h2, w2, self.c_cat))
        xb = x0m
        for i in range(len(self.layer_nums)):
            if i > 0:
                xb = self._block(P, W, i, xb, False, 0, "B", None)
            fused = self._act("B.fuse%d" % i, (B,) + tuple(xb.shape[1:]))
            pos = 0
            for b, n in enumerate(record_len):
                ops.att_fuse_fwd(xb.hi[pos:pos + n], fused.narrow_n(b, 1))
                pos += n
            c0 = sum(self.up_filters[:i])
            with self._on_side():
                self._deblock(P, W, i, fused, catB.slice_c(c0, c0 + self.up_filters[i]), False, 0, "B", None)
        self._join_side()
        _, _, heads = self._shrink_heads(P, W, catB, "B")
        return heads, dict(comm_rate=nz, ones=ones, hw=hm * wm)
