"""Host-side orchestration of the legacy `point_pillar_where2comm` path (BASELINE config 1) on the C-ABI kernels.

Same kernels as the airv2x Where2comm engine; what differs (models/point_pillar_where2comm.py:26-151): ONE PillarVFE
for all agents (`pillar_vfe.*`), the backbone evaluated once before the fusion, a 3x3 stride-2 shrink header — so the
communication mask lives at half the resolution of the level-0 features and is bilinearly resized
(where2comm_fuse.py:230-236) — and 1-class heads without objectness. forward() / backward() / loss() are the parent's:
the differences are class attributes (shrink geometry, BatchNorm update counts, head rows, PointPillarLoss) and the
single-encoder _encode below, so the legacy model trains on the same kernels (fused PointPillarLoss, a2x_det_loss_legacy).
"""
import torch

from . import ops
from .ops import Act
from .w2c_engine import HEAD_PAD, W2CEngine


class LegacyW2CEngine(W2CEngine):
    shrink_k0 = 3
    # the backbone runs once, then its blocks again inside the fusion (point_pillar_where2comm.py:118-146): block 0 (shared
    # here) sees 2 running-stat updates, everything else 1 per pass
    upd_block0, upd_pass_a, upd_pass_b = 2, 1, 1
    legacy_loss = True

    def _head_rows(self):
        return (("cls_head", 0), ("reg_head", self.A))

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
        self.k_on_device = False

    def _pack_weights(self, P):
        W, jobs = {}, []
        for i, ln in enumerate(self.layer_nums):
            for k in range(ln + 1):
                name = "backbone.blocks.%d.%d.weight" % (i, 1 + 3 * k)
                w = P[name]
                co, ci = w.shape[0], w.shape[1]
                W[name] = self._packed(name, (9, co, ci), (9, ci, co))
                jobs.append(ops.conv_pack_job(w, W[name], f32=not self.split))
            name = "backbone.deblocks.%d.0.weight" % i
            w = P[name]
            s = self.up_strides[i]
            ci, co = w.shape[0], w.shape[1]
            W[name] = self._packed(name, (1, s * s * co, ci), (s * s, ci, co))
            jobs.append(ops.deconv_pack_job(w, W[name], f32=not self.split))
        for idx in (0, 2):
            name = "shrink_conv.layers.0.double_conv.%d.weight" % idx
            w = P[name]
            co, ci = w.shape[0], w.shape[1]
            W[name] = self._packed(name, (9, co, ci), (9, ci, co))
            jobs.append(ops.conv_pack_job(w, W[name], f32=not self.split))
        fresh = ("packed", "heads") not in self.bufs
        hp = self._packed("heads", (1, HEAD_PAD, self.c_shrink), (1, self.c_shrink, HEAD_PAD))
        hb = self._buf("heads.b", (HEAD_PAD,))
        if fresh:
            for t in (hp.f32, hp.f16, hp.d32, hp.d16, hb):
                t.zero_()
        for name, row0 in (("cls_head", 0), ("reg_head", self.A)):
            jobs.append(ops.conv_pack_job(P[name + ".weight"], hp, row0, f32=not self.split))
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
        seg = None
        if "raw" in lidar:   # raw clouds: voxelise on the GPU (shared per-agent slabs, one segment list over all agents)
            raw = dict(lidar["raw"])
            raw["types"] = ["vehicle"] * n_total
            lay = dict(layout)
            lay["agent_map"] = {"vehicle": layout["identity_map"]}
            lidar = self._voxelize(raw, lay)["vehicle"]
            seg = lidar["seg"]
        vox, num, coords = lidar["voxel_features"], lidar["voxel_num_points"], lidar["voxel_coords"]
        w = P[pre + ".linear.weight"]
        scale, shift = self._buf("pfn.scale", (64,)), self._buf("pfn.shift", (64,))
        amap = layout["identity_map"]
        if training:   # batch statistics over all M*32 rows from the moments (PFNLayer, airv2x_pillar_vfe.py:27-49 = pillar_vfe.py)
            mean, invstd = self._buf("pfn.mean", (64,)), self._buf("pfn.invstd", (64,))
            moments = self._buf("pfn.moments", (65,), torch.float64)
            ops.pfn_moments(vox, num, coords, geom, moments, seg=seg)
            rows = vox.shape[0] * 32
            ops.pfn_stats_finalize(moments, rows, w, P[pre + ".norm.weight"], P[pre + ".norm.bias"], 1,
                                   P[pre + ".norm.running_mean"], P[pre + ".norm.running_var"], scale, shift, mean, invstd,
                                   seg=seg)
            amax = self._buf("pfn.amax", (vox.shape[0], 64), torch.uint8)
            ops.pfn_scatter(vox, num, coords, geom, w, scale, shift, amap, canvas, amax=amax, nz=nzc, write_hi=hi, seg=seg)
            if record is not None:
                record.append(dict(kind="pfn", type="all", vox=vox, num=num, coords=coords, geom=geom, pre=pre, scale=scale,
                                   shift=shift, mean=mean, invstd=invstd, amap=amap, amax=amax, moments=moments, rows=rows,
                                   seg=seg))
        else:
            ops.bn_eval_affine(P[pre + ".norm.weight"], P[pre + ".norm.bias"], P[pre + ".norm.running_mean"],
                               P[pre + ".norm.running_var"], scale, shift)
            ops.pfn_scatter(vox, num, coords, geom, w, scale, shift, amap, canvas, nz=nzc, write_hi=hi, seg=seg)
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
