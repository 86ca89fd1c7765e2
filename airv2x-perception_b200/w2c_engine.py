"""Host-side orchestration of the Where2comm hot path on the C-ABI kernels.

One `W2CEngine` owns every activation / gradient buffer (allocated once per shape signature, so addresses are
stable and the step can be captured in a CUDA graph) and issues the kernel sequence for

    voxels -> PillarVFE+scatter -> BEV backbone -> shrink -> cls head (single-agent confidence)
           -> communication mask -> multi-scale attention fusion -> deblocks -> shrink -> heads

forward (eval or train-mode BatchNorm) and backward (all parameter gradients). torch is used only for device
memory, streams and tiny parameter concatenations. Mirrors
opencood/models/airv2x_where2com.py:117-179 and where2comm_modules/where2comm_fuse.py:198-263 with the redundant
second backbone evaluation (airv2x_where2com.py:124) folded into a repeated running-stat update.
"""
import os
import random

import torch

from . import ops
from .ops import Act

AGENT_TYPES = ("vehicle", "rsu", "drone")
TYPE_PREFIX = {"vehicle": "veh_models", "rsu": "rsu_models", "drone": "drone_models"}
HEAD_PAD = 64  # fused head GEMM width (cls | reg | obj | zero pad); 64 keeps split-mode K a multiple of 64


class W2CEngine:
    # What differs between the airv2x model and the legacy `point_pillar_where2comm` (LegacyW2CEngine overrides):
    shrink_k0 = 1        # kernel of shrink_conv.layers.0.double_conv.0 (airv2x yaml: 1x1 s1; legacy yamls: 3x3 s2)
    shrink_stride = 1
    # BatchNorm running-stat updates per training step. airv2x evaluates the backbone twice back to back and its blocks a
    # third time inside the fusion (airv2x_where2com.py:119,124, where2comm_fuse.py:218): block 0 (shared here) 3, the other
    # un-masked layers 2, the masked pass 1. The legacy model evaluates the backbone once (point_pillar_where2comm.py:118).
    upd_block0, upd_pass_a, upd_pass_b = 3, 2, 1
    legacy_loss = False  # PointPillarLoss (1 class, no objectness) instead of PointPillarLossMultiClass

    def _head_rows(self):
        nc, nr = self.A * self.K, 7 * self.A
        return (("cls_head", 0), ("reg_head", nc), ("obj_head", nc + nr))

    def __init__(self, args, device, precision="split3"):
        assert precision in ("split3", "tf32"), precision
        self.args = args
        self.device = torch.device(device)
        self.split = precision == "split3"
        self.precision = precision
        mf = args["modality_fusion"]
        bb = mf["base_bev_backbone"]
        self.layer_nums = list(bb["layer_nums"])
        self.layer_strides = list(bb["layer_strides"])
        self.num_filters = list(bb["num_filters"])
        self.up_strides = list(bb["upsample_strides"])
        self.up_filters = list(bb["num_upsample_filter"])
        assert all(s == 2 for s in self.layer_strides), "backbone blocks must have stride 2"
        assert len(self.up_strides) == len(self.layer_nums), "extra deblock not supported"
        sh = mf["shrink_header"]
        assert sh["use"] and list(sh["kernal_size"]) == [1] and list(sh["stride"]) == [1] and list(sh["padding"]) == [0], \
            "only the airv2x shrink header (1x1 s1 + 3x3) is implemented"
        assert not mf.get("compression", 0), "NaiveCompressor (compression > 0) is not implemented"
        self.c_cat = sum(self.up_filters)
        self.c_shrink = sh["dim"][0]
        assert sh["input_dim"] == self.c_cat
        self.A = args["anchor_number"]
        self.K = args["num_class"]
        self.n_head = self.A * self.K + 7 * self.A + (self.A if args["obj_head"] else 0)
        assert args["obj_head"], "obj_head: false not implemented"
        assert self.n_head <= HEAD_PAD
        fa = args["where2com_fusion"]
        assert fa["multi_scale"], "single-scale Where2comm not implemented"
        self.fully = bool(fa["fully"])
        self.comm = fa["communication"]
        self.bufs = {}
        self.saved = None
        self._arena_off = 0
        self.side = None  # second stream: weight gradients run beside the data-gradient chain (fills idle SMs)
        self.use_side_stream = True
        # BN backward pass 1 inside the producing data-gradient epilogue: correct (tests) but measured SLOWER on B200
        # (round 1: 8.79 vs 8.24 ms per step; round 2 with the 8-warp transposed epilogue: 7.21 vs 6.80 ms — the dgrad
        # epilogue is on the critical path, the separate pass overlaps) -> off unless A2X_FUSE_BN_BWD=1
        self.fuse_bn_bwd_reduce = os.environ.get("A2X_FUSE_BN_BWD", "0") == "1"
        self.k_on_device = False

    # ------------------------------------------------------------------ buffers
    def _buf(self, name, shape, dtype=torch.float32):
        key = (name, tuple(shape), dtype)
        t = self.bufs.get(key)
        if t is None:
            t = torch.empty(tuple(shape), device=self.device, dtype=dtype)
            self.bufs[key] = t
        return t

    def _act(self, name, shape, split=None):
        split = self.split if split is None else split
        hi = self._buf(name, shape)
        return Act(hi, self._buf(name + ".b16", (2,) + tuple(shape), torch.bfloat16) if split else None)

    def _zeroed(self, name, nbytes_elems, dtype):
        """A slice of a per-step arena that is zeroed once per step (sums, packed weight gradients)."""
        arena = self.bufs.get(("arena", dtype))
        off = self._arena_off_map.setdefault(dtype, 0)
        need = off + nbytes_elems
        if arena is None or arena.numel() < need:
            # first pass discovers the size; grow and re-zero. Slices already handed out this step (and raw pointers
            # to them in job tables) must stay valid until the step ends: retire, do not free, the old arena.
            new = torch.zeros(max(need * 2, 1 << 16), device=self.device, dtype=dtype)
            if arena is not None:
                self._retired.append(arena)
            self.bufs[("arena", dtype)] = new
            arena = new
        self._arena_off_map[dtype] = need
        return arena[off:need]

    def _begin_step(self):
        self._arena_off_map = {}
        self._retired = []
        for dt in (torch.float64, torch.float32):
            a = self.bufs.get(("arena", dt))
            if a is not None:
                a.zero_()

    # ------------------------------------------------------------------ weights
    def _packed(self, name, f_shape, d_shape):
        pk = self.bufs.get(("packed", name))
        if pk is None:
            bf = torch.bfloat16
            pk = ops.PackedW(self._buf(name + ".f32", f_shape), self._buf(name + ".f16", (2,) + tuple(f_shape), bf),
                             self._buf(name + ".d32", d_shape), self._buf(name + ".d16", (2,) + tuple(d_shape), bf))
            self.bufs[("packed", name)] = pk
        return pk

    def _pack_weights(self, P):
        """Re-lay every conv / deconv weight for this step (OIHW -> tap-major forward and data-gradient operands, bf16
        split planes) with ONE batched launch over a cached device job table; the three 1x1 heads share one fused
        64-column GEMM operand (cls | reg | obj | zero pad)."""
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
        for idx, k in ((0, 1), (2, 3)):
            name = "shrink_conv.layers.0.double_conv.%d.weight" % idx
            w = P[name]
            co, ci = w.shape[0], w.shape[1]
            W[name] = self._packed(name, (k * k, co, ci), (k * k, ci, co))
            jobs.append(ops.conv_pack_job(w, W[name], f32=not self.split))
        nc, nr = self.A * self.K, 7 * self.A
        fresh = ("packed", "heads") not in self.bufs
        hp = self._packed("heads", (1, HEAD_PAD, self.c_shrink), (1, self.c_shrink, HEAD_PAD))
        hb = self._buf("heads.b", (HEAD_PAD,))
        if fresh:  # the pad rows / pad bias stay zero: no job ever writes them
            for t in (hp.f32, hp.f16, hp.d32, hp.d16, hb):
                t.zero_()
        for name, row0 in (("cls_head", 0), ("reg_head", nc), ("obj_head", nc + nr)):
            jobs.append(ops.conv_pack_job(P[name + ".weight"], hp, row0, f32=not self.split))
            jobs.append(ops.copy_pack_job(P[name + ".bias"], hb, row0))
        W["heads"] = hp
        W["heads.bias"] = hb
        ops.pack_weights_batched(self._job_table("pack", jobs))
        return W

    def _job_table(self, name, jobs):
        """device job table cached by the pointers it refers to (stable across steps: one upload)"""
        key = tuple((j.src, j.dst if hasattr(j, "dst") else (j.f32 or j.f16)) for j in jobs)
        ent = self.bufs.get(("jobs", name))
        if ent is None or ent[0] != key:
            ent = (key, ops.JobTable(jobs, self.device))
            self.bufs[("jobs", name)] = ent
        return ent[1]

    # ------------------------------------------------------------------ layers
    def _bn_params(self, P, bn, training, z, n_updates, tag, sums=None):
        C = z.shape[3]
        scale = self._buf(tag + ".scale", (C,))
        shift = self._buf(tag + ".shift", (C,))
        if not training:
            ops.bn_eval_affine(P[bn + ".weight"], P[bn + ".bias"], P[bn + ".running_mean"], P[bn + ".running_var"],
                               scale, shift)
            return scale, shift, None, None
        mean = self._buf(tag + ".mean", (C,))
        invstd = self._buf(tag + ".invstd", (C,))
        if sums is None:
            sums = self._zeroed(tag + ".sums", 2 * C, torch.float64)
            ops.channel_stats(z, sums)
        count = z.shape[0] * z.shape[1] * z.shape[2]
        ops.bn_finalize(sums, count, P[bn + ".weight"], P[bn + ".bias"], n_updates, P[bn + ".running_mean"],
                        P[bn + ".running_var"], scale, shift, mean, invstd)
        return scale, shift, mean, invstd

    def _bn_train_act(self, P, bn, z, sums, n_updates, tag, out, need_hi=True):
        """train-mode BN (finalize fused into the apply pass) + ReLU + operand split; returns the saved statistics.
        need_hi False: every consumer of `out` is a split GEMM, so its fp32 plane is not written (4 of 12 B / element)"""
        C = z.shape[3]
        scale, shift = self._buf(tag + ".scale", (C,)), self._buf(tag + ".shift", (C,))
        mean, invstd = self._buf(tag + ".mean", (C,)), self._buf(tag + ".invstd", (C,))
        count = z.shape[0] * z.shape[1] * z.shape[2]
        ops.bn_train_act(z, sums, count, P[bn + ".weight"], P[bn + ".bias"], n_updates, P[bn + ".running_mean"],
                         P[bn + ".running_var"], scale, shift, mean, invstd, True, out, write_hi=need_hi)
        return scale, shift, mean, invstd

    def _conv_bn_relu(self, P, W, conv, bn, x, stride, training, n_updates, tag, record, need_hi=True):
        n, h, w, _ = x.shape
        cout = P[conv].shape[0]
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        y = self._act(tag + ".y", (n, ho, wo, cout))
        wf = W[conv]
        if not training:
            scale, shift, _, _ = self._bn_params(P, bn, False, y.hi, 0, tag)
            ops.conv_fwd(x, wf, 3, stride, y, scale=scale, shift=shift, relu=True, write_hi=need_hi)
            return y
        z = self._buf(tag + ".z", (n, ho, wo, cout))
        sums = self._zeroed(tag + ".sums", 2 * cout, torch.float64)
        ops.conv_fwd(x, wf, 3, stride, Act(z), stats=sums)  # BN batch statistics come out of the GEMM epilogue
        scale, shift, mean, invstd = self._bn_train_act(P, bn, z, sums, n_updates, tag, y, need_hi)
        if record is not None:
            record.append(dict(kind="conv", conv=conv, bn=bn, x=x, z=z, y=y, stride=stride, scale=scale, shift=shift,
                               mean=mean, invstd=invstd, tag=tag))
        return y

    def _block(self, P, W, i, x, training, n_updates, tag, record, need_hi=True):
        """need_hi: whether anything reads the fp32 value of the block's OUTPUT (mask multiply, attention fusion); the
        layers inside a block only feed the next conv, so their fp32 planes are never written in train mode"""
        p = "backbone.blocks.%d" % i
        last = self.layer_nums[i]
        x = self._conv_bn_relu(P, W, p + ".1.weight", p + ".2", x, 2, training, n_updates, "%s.b%d.0" % (tag, i), record,
                               need_hi and last == 0)
        for k in range(last):
            x = self._conv_bn_relu(P, W, "%s.%d.weight" % (p, 4 + 3 * k), "%s.%d" % (p, 5 + 3 * k), x, 1, training,
                                   n_updates, "%s.b%d.%d" % (tag, i, k + 1), record, need_hi and k == last - 1)
        return x

    def _deblock(self, P, W, i, x, out_slice, training, n_updates, tag, record):
        """ConvTranspose(k = s) + BN + ReLU written into a channel slice of the concat buffer."""
        conv = "backbone.deblocks.%d.0.weight" % i
        bn = "backbone.deblocks.%d.1" % i
        s = self.up_strides[i]
        cout = self.up_filters[i]
        wf = W[conv]
        n, h, w, _ = x.shape
        tg = "%s.d%d" % (tag, i)
        if not training:
            scale, shift, _, _ = self._bn_params(P, bn, False, out_slice.hi, 0, tg)
            ops.deconv_fwd(x, wf, cout, s, out_slice, scale=scale, shift=shift, relu=True, write_hi=False)
            return
        z = self._buf(tg + ".z", (n, h * s, w * s, cout))
        sums = self._zeroed(tg + ".sums", 2 * cout, torch.float64)
        ops.deconv_fwd(x, wf, cout, s, Act(z), stats=sums)
        scale, shift, mean, invstd = self._bn_train_act(P, bn, z, sums, n_updates, tg, out_slice, False)  # feeds the shrink conv
        if record is not None:
            record.append(dict(kind="deconv", conv=conv, bn=bn, x=x, z=z, y=out_slice, stride=s, scale=scale,
                               shift=shift, mean=mean, invstd=invstd, tag=tg, level=i))

    def _shrink_heads(self, P, W, cat, tag, heads_only_cls=False):
        """shrink (1x1+bias+ReLU, 3x3+bias+ReLU) then the fused 1x1 heads (cls | reg | obj | pad) -> [n,h,w,32]."""
        n, h, w, _ = cat.shape
        y1 = self._act(tag + ".s1", (n, h, w, self.c_shrink))
        y2 = self._act(tag + ".s2", (n, h, w, self.c_shrink))
        ops.conv_fwd(cat, W["shrink_conv.layers.0.double_conv.0.weight"], 1, 1, y1,
                     shift=P["shrink_conv.layers.0.double_conv.0.bias"], relu=True)
        ops.conv_fwd(y1, W["shrink_conv.layers.0.double_conv.2.weight"], 3, 1, y2,
                     shift=P["shrink_conv.layers.0.double_conv.2.bias"], relu=True)
        heads = self._buf(tag + ".heads", (n, h, w, HEAD_PAD))
        ops.conv_fwd(y2, W["heads"], 1, 1, Act(heads), shift=W["heads.bias"])
        return y1, y2, heads

    # ------------------------------------------------------------------ voxelisation (raw point clouds)
    def _voxelize(self, raw, layout):
        """raw: dict(points [P,4] f32 device, offsets int32 device [N+1], types [N] (scene-major), voxel_size,
        lidar_range, max_points, max_voxels). Returns the per-type lidar dict over shared per-agent slabs."""
        N = layout["n_total"]
        pts = raw["points"]
        cap = int(raw["max_voxels"])
        nx, ny = layout["nx"], layout["ny"]
        nz = 1
        ws_bytes = ops.voxelize_workspace_bytes(N, pts.shape[0], nx, ny, nz, cap)
        ws = self._buf("vox.ws", (ws_bytes,), torch.uint8)
        vox = self._buf("vox.voxels", (N * cap, 32, 4))
        coords = self._buf("vox.coords", (N * cap, 4), torch.int32)
        num = self._buf("vox.num", (N * cap,), torch.int32)
        counts = self._buf("vox.counts", (N,), torch.int32)
        ops.voxelize(pts, raw["offsets"], N, raw["lidar_range"], raw["voxel_size"], int(raw["max_points"]), cap, cap, ws,
                     vox, coords, num, counts, ego_flags=raw.get("ego_flags"),
                     strict_range=bool(raw.get("filter", False)), transforms=raw.get("transforms"))
        ident = layout.get("identity_map")
        out = {}
        for t in AGENT_TYPES:
            ids = [i for i, tt in enumerate(raw["types"]) if tt == t]
            if not ids:
                continue
            key = ("seg_ids", t, tuple(ids))
            seg_ids = self.bufs.get(key)
            if seg_ids is None:
                seg_ids = torch.tensor(ids, dtype=torch.int32, device=self.device)
                self.bufs[key] = seg_ids
            out[t] = {"voxel_features": vox, "voxel_num_points": num, "voxel_coords": coords,
                      "seg": ops.pfn_segments(seg_ids, counts, cap), "agent_map": ident, "counts": counts}
        return out

    # ------------------------------------------------------------------ encoder (PillarVFE + scatter)
    def _encode(self, P, lidar, layout, training, record):
        n_total, ny, nx = layout["n_total"], layout["ny"], layout["nx"]
        canvas = self._act("canvas", (n_total, ny, nx, 64))
        # split mode: the canvas only feeds the block-0 tap-GEMM, which reads the bf16 planes -> the fp32 plane is neither
        # cleared nor written, and `spatial_features.count_nonzero()` (airv2x_where2com.py:122) is counted by the scatter
        cam = lidar.get("camera_bev")      # {type: [n_type, ny, nx, 64]}: camera-encoder BEV maps to average with the pillars'
        if cam and training:
            raise NotImplementedError("camera + lidar fusion (fuse_bev) is implemented for the eval forward only")
        hi = canvas.b16 is None or bool(cam)
        nzc = self._buf("canvas.nz", (1,), torch.int64)
        with self._on_side():   # the clears (180 MB of canvas planes) run beside the voxeliser / PFN statistics; joined before the scatter
            if hi:
                canvas.hi.zero_()
            if canvas.b16 is not None:
                canvas.b16.zero_()
            nzc.zero_()
        self._canvas_nz = nzc
        joined = False
        if "raw" in lidar:
            lidar = self._voxelize(lidar["raw"], layout)
        for t in AGENT_TYPES:
            if t not in lidar:
                continue
            vox, num, coords = lidar[t]["voxel_features"], lidar[t]["voxel_num_points"], lidar[t]["voxel_coords"]
            seg = lidar[t].get("seg")
            la = self.args[t]["lidar"]
            geom = ops.pfn_geom(la["voxel_size"], la["lidar_range"], nx, ny)
            pre = TYPE_PREFIX[t] + ".0.0.pfn_layers.0"
            w = P[pre + ".linear.weight"]
            scale = self._buf("pfn.%s.scale" % t, (64,))
            shift = self._buf("pfn.%s.shift" % t, (64,))
            amap = lidar[t].get("agent_map")
            if amap is None:
                amap = layout["agent_map"][t]
            if training:
                mean = self._buf("pfn.%s.mean" % t, (64,))
                invstd = self._buf("pfn.%s.invstd" % t, (64,))
                moments = self._buf("pfn.%s.moments" % t, (65,), torch.float64)
                ops.pfn_moments(vox, num, coords, geom, moments, seg=seg)
                rows = vox.shape[0] * 32
                ops.pfn_stats_finalize(moments, rows, w, P[pre + ".norm.weight"], P[pre + ".norm.bias"], 1,
                                       P[pre + ".norm.running_mean"], P[pre + ".norm.running_var"], scale, shift, mean,
                                       invstd, seg=seg)
                amax = self._buf("pfn.%s.amax" % t, (vox.shape[0], 64), torch.uint8)
                if not joined:
                    self._join_side()
                    joined = True
                ops.pfn_scatter(vox, num, coords, geom, w, scale, shift, amap, canvas, amax=amax, seg=seg, nz=nzc, write_hi=hi)
                if record is not None:
                    record.append(dict(kind="pfn", type=t, vox=vox, num=num, coords=coords, geom=geom, pre=pre,
                                       scale=scale, shift=shift, mean=mean, invstd=invstd, amap=amap, amax=amax,
                                       moments=moments, rows=rows, seg=seg))
            else:
                ops.bn_eval_affine(P[pre + ".norm.weight"], P[pre + ".norm.bias"], P[pre + ".norm.running_mean"],
                                   P[pre + ".norm.running_var"], scale, shift)
                if not joined:
                    self._join_side()
                    joined = True
                ops.pfn_scatter(vox, num, coords, geom, w, scale, shift, amap, canvas, seg=seg, nz=nzc, write_hi=hi)
        if not joined:
            self._join_side()
        if cam:
            # fuse_bev (common_modules/airv2x_base_model.py:167-177): the mean over the type's modality encoders of
            # `spatial_features` = 0.5 * (pillar canvas + camera BEV), per agent of that type; comm_rate counts the fused map
            half = self._buf("cam.half", (64,))
            half.fill_(0.5)
            tmp = self._buf("cam.tmp", (1, ny, nx, 64))
            for t, bev in cam.items():
                rows = layout.setdefault("agent_map_host", {}).get(t)
                if rows is None:
                    rows = layout["agent_map_host"][t] = [int(v) for v in layout["agent_map"][t].tolist()]
                assert tuple(bev.shape) == (len(rows), ny, nx, 64), "camera_bev[%s] must be [n_%s, ny, nx, 64] (NHWC)" % (t, t)
                for i, r in enumerate(rows):
                    ops.dropout_apply(bev[i:i + 1], None, 0, Act(tmp), residual=canvas.hi[r:r + 1])    # pillars + camera
                    ops.affine_act(tmp, half, None, False, canvas.narrow_n(r, 1))                      # * 0.5 -> fp32 + split planes
            nzc.zero_()
            ops.count_nonzero(canvas.hi, nzc)
        return canvas

    # ------------------------------------------------------------------ forward
    def forward(self, P, lidar, layout, training, k_list=None):
        """lidar: {type: {voxel_features, voxel_num_points, voxel_coords}} device tensors.
        layout: dict(n_total, ny, nx, record_len=[...], agent_map={type: int32 device tensor}).
        Returns heads [B, H/2, W/2, 32] (NHWC, channels cls|reg|obj|pad) and an aux dict of device scalars."""
        self._begin_step()
        rec = [] if training else None
        with self._on_side():   # the weight re-layout (0.1 ms) runs beside the voxeliser / PFN; _encode joins before its scatter
            W = self._pack_weights(P)
        record_len = layout["record_len"]
        B, N = len(record_len), layout["n_total"]
        canvas = self._encode(P, lidar, layout, training, rec)
        self._join_side()   # the packed weights (a subclass's _encode may not have joined)
        nz = self._buf("comm_rate", (1,), torch.int64)
        nz.copy_(self._canvas_nz)

        # ---- pass A: un-masked backbone on every agent map (single-agent confidence for the mask)
        # block 0 is shared with the fusion pass; in train mode its BNs see 3 identical running-stat updates
        # (airv2x_where2com.py:119,124 + where2comm_fuse.py:218), the other pass-A BNs 2.
        x0 = self._block(P, W, 0, canvas, training, self.upd_block0, "A", rec)
        h2, w2 = x0.shape[1], x0.shape[2]
        catA = self._act("A.cat", (N, h2, w2, self.c_cat))
        xa = x0
        # deblock i only feeds the concat buffer, so it runs on the side stream beside block i+1 (the deep blocks have
        # fewer 128-pixel tiles than the GPU has SMs: the HBM-bound 1x1 deblock GEMMs fill the idle ones)
        for i in range(len(self.layer_nums)):
            if i > 0:
                xa = self._block(P, W, i, xa, training, self.upd_pass_a, "A", None, need_hi=False)
            c0 = sum(self.up_filters[:i])
            with self._on_side():
                self._deblock(P, W, i, xa, catA.slice_c(c0, c0 + self.up_filters[i]), training, self.upd_pass_a, "A", None)
        self._join_side()
        _, _, headsA = self._shrink_heads(P, W, catA, "A")

        # ---- communication mask (where2comm_fuse.py:83-149), at the resolution of the head map: the legacy model's
        # stride-2 shrink header halves it, and the mask is resized bilinearly to the level-0 features (:230-236)
        hm, wm = headsA.shape[1], headsA.shape[2]
        hw = hm * wm
        mask_lo = self._buf("mask" if (hm, wm) == (h2, w2) else "mask.lo", (N, hm, wm))
        ones = self._buf("mask.ones", (B,))
        if self.fully:
            mask_lo.fill_(1.0)
            ones.fill_(float("nan"))
        else:
            conf = self._buf("conf", (N, hm, wm))
            smooth = self._buf("smooth", (N, hm, wm))
            ops.comm_confidence(headsA, self.A * self.K, conf)
            gs = self.comm.get("gaussian_smooth")
            ksz = gs["k_size"] if gs else 0
            gw = P.get("fusion_net.naive_communication.gaussian_filter.weight")
            gb = P.get("fusion_net.naive_communication.gaussian_filter.bias")
            thr = float(self.comm["threshold"])
            if training:
                # K per scene from Python's `random` exactly like the reference (where2comm_fuse.py:106)
                if k_list is None:
                    k_list = [int(hw * random.uniform(0, 1)) for _ in range(B)]
                k_host = self.set_k(k_list, record_len)
                k_dev = self._buf("k_dev", (N,), torch.int32)
                if not self.k_on_device:  # pipelined mode stages K with the other inputs (no H2D inside the graph)
                    k_dev.copy_(k_host, non_blocking=True)
                ops.comm_smooth_mask(conf, gw, gb, ksz, N, hm, wm, thr, False, smooth, mask_lo)
                ops.comm_topk_mask(smooth, N, hw, k_dev, mask_lo)
            elif thr:
                ops.comm_smooth_mask(conf, gw, gb, ksz, N, hm, wm, thr, True, smooth, mask_lo)
            else:
                mask_lo.fill_(1.0)
            ops.comm_rate_ego(mask_lo, hw, B, layout["scene_start"], layout["scene_len"], ones)
        if (hm, wm) != (h2, w2):
            mask = self._buf("mask", (N, h2, w2))
            ops.resize_bilinear(mask_lo, mask)
        else:
            mask = mask_lo

        # ---- pass B: masked multi-scale fusion (where2comm_fuse.py:214-262)
        x0m = self._act("B.x0m", x0.shape)
        ops.affine_act(self._full(x0, "B.x0full"), None, None, False, x0m, mask=mask)
        catB = self._act("B.cat", (B, h2, w2, self.c_cat))
        xb = x0m
        levels = []
        for i in range(len(self.layer_nums)):
            if i > 0:
                xb = self._block(P, W, i, xb, training, self.upd_pass_b, "B", rec)
            n_, hh, ww, cc = xb.shape
            fused = self._act("B.fuse%d" % i, (B, hh, ww, cc))
            xfull = self._full(xb, "B.xfull%d" % i)
            pos = 0
            for b, n in enumerate(record_len):
                ops.att_fuse_fwd(xfull[pos:pos + n], fused.narrow_n(b, 1))
                pos += n
            levels.append(dict(x=xb, xfull=xfull, fused=fused))
            c0 = sum(self.up_filters[:i])
            with self._on_side():
                self._deblock(P, W, i, fused, catB.slice_c(c0, c0 + self.up_filters[i]), training, self.upd_pass_b, "B", rec)
        self._join_side()
        y1, y2, heads = self._shrink_heads(P, W, catB, "B")
        if training:
            self.saved = dict(rec=rec, W=W, levels=levels, mask=mask, x0=x0, x0m=x0m, catB=catB, y1=y1, y2=y2,
                              heads=heads, layout=layout, canvas=canvas)
        aux = dict(comm_rate=nz, ones=ones, hw=hw)
        return heads, aux

    # ------------------------------------------------------------------ agent-parallel inference (one agent per GPU)
    def ap_regions(self, h2, w2):
        """float offsets of one rank's exchange buffer: header | cell indices | selected level-0 rows | dense levels 1.."""
        hw = h2 * w2
        reg, pos = {}, 0
        for name, n in (("hdr", 64), ("idx", (hw + 3) // 4 * 4), ("vals", hw * self.num_filters[0])):
            reg[name] = (pos, n)
            pos += n
        hh, ww = h2, w2
        for i in range(1, len(self.layer_nums)):
            hh, ww = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
            reg["lvl%d" % i] = (pos, hh * ww * self.num_filters[i], (hh, ww, self.num_filters[i]))
            pos += hh * ww * self.num_filters[i]
        reg["total"] = pos
        return reg

    def forward_agent_parallel(self, P, lidar, layout, rank, n_agents, xbuf, exchange, done):
        """Eval-mode Where2comm with the scene's agents sharded one per rank. Every rank runs the
        per-agent half (encoder, un-masked pass for its own confidence map, mask, masked blocks), publishes ONE buffer —
        the warp-ballot-compacted level-0 cells its mask selected plus its dense level-1.. maps — and pulls the peers'
        buffers (`exchange()` -> device table of per-rank buffer pointers: local copies after an all-gather or peer
        memory) straight into the dense per-agent tensors the fusion kernels read. `xbuf`: this rank's flat fp32 buffer
        of ap_regions()["total"] elements. Returns (heads, aux) like forward(); identical on every rank."""
        self._begin_step()
        W = self._pack_weights(P)
        canvas = self._encode(P, lidar, layout, False, None)
        x0 = self._block(P, W, 0, canvas, False, 0, "A", None)
        _, h2, w2, c0 = x0.shape
        hw = h2 * w2
        reg = self.ap_regions(h2, w2)
        assert xbuf.numel() >= reg["total"] and xbuf.dtype == torch.float32
        hdr = xbuf[reg["hdr"][0]:reg["hdr"][0] + 64].view(torch.int32)
        idx = xbuf[reg["idx"][0]:reg["idx"][0] + reg["idx"][1]].view(torch.int32)
        vals = xbuf[reg["vals"][0]:reg["vals"][0] + reg["vals"][1]]
        hdr[2:4].view(torch.int64).copy_(self._canvas_nz)
        # pass A on the own map -> confidence -> mask (where2comm_fuse.py:83-149, eval branch)
        catA = self._act("A.cat", (1, h2, w2, self.c_cat))
        xa = x0
        for i in range(len(self.layer_nums)):
            if i > 0:
                xa = self._block(P, W, i, xa, False, 0, "A", None)
            c_lo = sum(self.up_filters[:i])
            self._deblock(P, W, i, xa, catA.slice_c(c_lo, c_lo + self.up_filters[i]), False, 0, "A", None)
        _, _, headsA = self._shrink_heads(P, W, catA, "A")
        mask = self._buf("mask", (1, h2, w2))
        thr = float(self.comm["threshold"])
        if self.fully or not thr:
            mask.fill_(1.0)
        else:
            conf, smooth = self._buf("conf", (1, h2, w2)), self._buf("smooth", (1, h2, w2))
            ops.comm_confidence(headsA, self.A * self.K, conf)
            gs = self.comm.get("gaussian_smooth")
            ops.comm_smooth_mask(conf, P.get("fusion_net.naive_communication.gaussian_filter.weight"),
                                 P.get("fusion_net.naive_communication.gaussian_filter.bias"), gs["k_size"] if gs else 0,
                                 1, h2, w2, thr, True, smooth, mask)
        # sparse feature select: what this agent transmits (the ego keeps every cell, where2comm_fuse.py:137-143)
        ego = rank == 0
        ops.mask_compact(x0.hi, mask[0], ego or self.fully, hdr, idx, vals)
        if ego:
            mask.fill_(1.0)
        x0m = self._act("B.x0m", x0.shape)
        ops.affine_act(x0.hi, None, None, False, x0m, mask=mask)
        xb = x0m
        for i in range(1, len(self.layer_nums)):
            xb = self._block(P, W, i, xb, False, 0, "B", None)
            o, n, _ = reg["lvl%d" % i]
            xbuf[o:o + n].copy_(xb.hi.reshape(-1))
        # ---- the path's one exchange
        table = exchange()
        xfull0 = self._buf("ap.x0", (n_agents, h2, w2, c0))
        ops.mask_decompact_ptrs(table, reg["idx"][0] * 4, reg["vals"][0] * 4, n_agents, xfull0)
        start = self._buf("ap.start", (1,), torch.int32)
        length = self._buf("ap.len", (1,), torch.int32)
        start.zero_()
        length.fill_(n_agents)
        hdrs = self._buf("ap.hdrs", (n_agents, 1, 1, 64))
        ops.regroup_ptrs(table, (1, 1, 64), start, length, 1, n_agents, Act(hdrs))
        levels = [xfull0]
        for i in range(1, len(self.layer_nums)):
            o, n, shp = reg["lvl%d" % i]
            xf = self._buf("ap.x%d" % i, (n_agents,) + shp)
            ops.regroup_ptrs(table + o * 4, shp, start, length, 1, n_agents, Act(xf))
            levels.append(xf)
        done()
        # ---- fusion (where2comm_fuse.py:214-262), replicated on every rank
        catB = self._act("B.cat", (1, h2, w2, self.c_cat))
        for i, xf in enumerate(levels):
            fused = self._act("B.fuse%d" % i, (1,) + tuple(xf.shape[1:]))
            ops.att_fuse_fwd(xf, fused)
            c_lo = sum(self.up_filters[:i])
            self._deblock(P, W, i, fused, catB.slice_c(c_lo, c_lo + self.up_filters[i]), False, 0, "B", None)
        _, _, heads = self._shrink_heads(P, W, catB, "B")
        hi = hdrs.view(n_agents, 64).view(torch.int32)
        aux = dict(comm_rate=hi[:, 2:4].contiguous().view(torch.int64).sum(), hw=hw,
                   ones=hi[:, 1].sum().to(torch.float32).reshape(1))
        return heads, aux

    def set_k(self, k_list, record_len):
        """write the per-scene top-K sizes into a pinned staging buffer for the asynchronous H2D copy that follows. A ring
        of 64 buffers: the host may run many steps ahead of the device without overwriting values not yet copied."""
        self._k_ring = (getattr(self, "_k_ring", -1) + 1) % 64
        k_host = self._pinned("k_host.%d" % self._k_ring, (sum(record_len),), torch.int32)
        pos = 0
        for b, n in enumerate(record_len):
            k_host[pos:pos + n] = int(k_list[b])
            pos += n
        return k_host

    def _pinned(self, name, shape, dtype):
        key = ("pinned", name, tuple(shape), dtype)
        t = self.bufs.get(key)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
            self.bufs[key] = t
        return t

    def _full(self, act, name):
        """fp32 value of a (possibly split) activation (the fp32 plane always holds the full value)."""
        return act.hi

    # ------------------------------------------------------------------ loss (fused value + gradient)
    def loss(self, heads, labels, cls_weight=1.0, reg_coe=2.0, want_grad=True):
        B = heads.shape[0]
        dheads = self._buf("dheads", heads.shape) if want_grad else None
        loss3 = self._buf("loss3", (3,), torch.float64)
        npos = self._buf("npos", (B,))
        ops.det_loss(heads, self.A, self.K, labels["targets"], labels["pos_equal_one"], labels.get("class_ids"),
                     float(cls_weight), float(reg_coe), npos, dheads, loss3, legacy=self.legacy_loss)
        return loss3, dheads

    # ------------------------------------------------------------------ backward
    class _Side:
        """`with self._on_side():` issues the enclosed launches on the side stream, ordered after everything issued so
        far on the main stream (event fork). backward() joins the side stream back at its end."""

        def __init__(self, eng):
            self.eng = eng

        def __enter__(self):
            e = self.eng
            if not e.use_side_stream:
                return
            if e.side is None:
                e.side = torch.cuda.Stream(device=e.device)
            e.side.wait_stream(torch.cuda.current_stream())
            self.ctx = torch.cuda.stream(e.side)
            self.ctx.__enter__()

        def __exit__(self, *a):
            if self.eng.use_side_stream:
                self.ctx.__exit__(*a)

    def _on_side(self):
        return W2CEngine._Side(self)

    def _join_side(self):
        if self.use_side_stream and self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)

    def backward(self, P, dheads, grads, sync=None):
        """dheads: [B,h,w,32] gradient w.r.t. the head logits. grads: dict name -> fp32 tensor (written).
        sync: dist.GradAverager whose flat buffer backs `grads` (scene-parallel training): the all-reduce of every gradient
        except the level-0 block's and the PillarVFE's starts as soon as those are final and runs beside the level-0
        backward; the rest is reduced at the end."""
        S = self.saved
        assert S is not None, "backward() needs a train-mode forward first"
        W, rec = S["W"], S["rec"]
        nc, nr = self.A * self.K, 7 * self.A
        B, h2, w2, _ = dheads.shape

        unpack = []  # weight / bias gradient re-layouts, done by ONE batched launch at the end of the side stream

        def wgrad_conv(x, dy, name, k, stride):
            cout = dy.shape[3]
            cin = x.shape[3]
            dwp = self._zeroed(name + ".dwp", k * k * cout * cin, torch.float32).view(k * k, cout, cin)
            ops.conv_wgrad(x, dy, k, stride, dwp)
            return dwp

        def bias_grad(g, C, outs):
            """out[c] = sum over pixels of the gradient; outs: [(tensor, first channel)] slices of the C channels"""
            sums = self._zeroed("bias.sums", 2 * C, torch.float64)
            ops.channel_stats(g.hi if isinstance(g, Act) else g, sums)
            for out, c0 in outs:
                unpack.append(ops.sums_unpack_job(sums, out, c0))

        # ---- heads
        dh = self._act("bwd.dheads", dheads.shape)
        ops.affine_act(dheads, None, None, False, dh)
        with self._on_side():
            dwp = wgrad_conv(S["y2"], dh, "heads", 1, 1)
            for name, row0 in self._head_rows():
                unpack.append(ops.conv_unpack_job(dwp, grads[name + ".weight"], row0))
            bias_grad(dheads, HEAD_PAD, [(grads[name + ".bias"], row0) for name, row0 in self._head_rows()])
        d_y2 = self._buf("bwd.d_y2", S["y2"].shape)
        ops.conv_dgrad(dh, W["heads"], 1, 1, d_y2)

        # ---- shrink conv 2 (3x3 + bias + ReLU)
        g2 = self._act("bwd.g2", S["y2"].shape)
        ops.relu_bwd(d_y2, S["y2"].hi, g2)
        n2 = "shrink_conv.layers.0.double_conv.2"
        with self._on_side():
            dwp = wgrad_conv(S["y1"], g2, n2 + ".weight", 3, 1)
            unpack.append(ops.conv_unpack_job(dwp, grads[n2 + ".weight"]))
            bias_grad(g2, self.c_shrink, [(grads[n2 + ".bias"], 0)])
        d_y1 = self._buf("bwd.d_y1", S["y1"].shape)
        ops.conv_dgrad(g2, W[n2 + ".weight"], 3, 1, d_y1)

        # ---- shrink conv 1 (1x1 + bias + ReLU)
        g1 = self._act("bwd.g1", S["y1"].shape)
        ops.relu_bwd(d_y1, S["y1"].hi, g1)
        n1 = "shrink_conv.layers.0.double_conv.0"
        with self._on_side():
            dwp = wgrad_conv(S["catB"], g1, n1 + ".weight", self.shrink_k0, self.shrink_stride)
            unpack.append(ops.conv_unpack_job(dwp, grads[n1 + ".weight"]))
            bias_grad(g1, self.c_shrink, [(grads[n1 + ".bias"], 0)])
        d_cat = self._buf("bwd.d_cat", S["catB"].shape)
        ops.conv_dgrad(g1, W[n1 + ".weight"], self.shrink_k0, self.shrink_stride, d_cat)

        # ---- deblocks (pass B) -> d(fused_i); fusion backward -> d(x_i) for every agent
        levels = S["levels"]
        record_len = S["layout"]["record_len"]
        d_levels = [None] * len(levels)
        deconv_recs = {r["level"]: r for r in rec if r["kind"] == "deconv"}
        for i, lv in enumerate(levels):
            r = deconv_recs[i]
            c0 = sum(self.up_filters[:i])
            dy = d_cat[..., c0:c0 + self.up_filters[i]]
            dz = self._act("bwd.dz.d%d" % i, r["z"].shape)
            sums = self._zeroed(r["tag"] + ".bsums", ops.bn_bwd_sums_len(r["z"].shape[3]), torch.float64)
            ops.bn_relu_bwd(dy, r["z"], r["scale"], r["shift"], r["mean"], r["invstd"], sums, dz,
                            grads[r["bn"] + ".weight"], grads[r["bn"] + ".bias"], write_hi=False)
            s = r["stride"]
            cin, cout = r["x"].shape[3], r["z"].shape[3]
            dwp = self._zeroed(r["conv"] + ".dwp", s * s * cin * cout, torch.float32).view(s * s, cin, cout)
            with self._on_side():
                ops.deconv_wgrad(r["x"], dz, s, dwp)
                unpack.append(ops.deconv_unpack_job(dwp, grads[r["conv"]]))
            d_fused = self._buf("bwd.d_fused%d" % i, lv["fused"].shape)
            ops.deconv_dgrad(dz, W[r["conv"]], s, d_fused)
            dx = self._buf("bwd.dx%d" % i, lv["x"].shape)
            pos = 0
            for b, n in enumerate(record_len):
                ops.att_fuse_bwd(lv["xfull"][pos:pos + n], d_fused[b:b + 1], dx[pos:pos + n])
                pos += n
            d_levels[i] = dx

        # ---- backbone blocks in reverse (pass B for levels >= 1, then block 0 through the mask)
        conv_recs = [r for r in rec if r["kind"] == "conv"]
        by_tag = {r["tag"]: r for r in conv_recs}

        def block_bwd(tag, i, dy, dx_first, accumulate_first, mask=None):
            """dy: gradient w.r.t. the block output; returns nothing, writes the input gradient into dx_first."""
            nl = self.layer_nums[i]
            sums_ready = False
            for k in range(nl, -1, -1):
                r = by_tag["%s.b%d.%d" % (tag, i, k)]
                dz = self._act("bwd.dz." + r["tag"], r["z"].shape)
                if not sums_ready:
                    sums = self._zeroed(r["tag"] + ".bsums", ops.bn_bwd_sums_len(r["z"].shape[3]), torch.float64)
                ops.bn_relu_bwd(dy, r["z"], r["scale"], r["shift"], r["mean"], r["invstd"], sums, dz,
                                grads[r["bn"] + ".weight"], grads[r["bn"] + ".bias"], sums_ready=sums_ready,
                                write_hi=False)
                cout, cin = r["z"].shape[3], r["x"].shape[3]
                dwp = self._zeroed(r["conv"] + ".dwp." + tag, 9 * cout * cin, torch.float32).view(9, cout, cin)
                with self._on_side():
                    ops.conv_wgrad(r["x"], dz, 3, r["stride"], dwp)
                    unpack.append(ops.conv_unpack_job(dwp, grads[r["conv"]]))
                sums_ready = False
                if k > 0:
                    dprev = self._buf("bwd.dprev.%s.b%d.%d" % (tag, i, k), r["x"].shape)
                    bn_stats = None
                    if r["stride"] == 1 and self.fuse_bn_bwd_reduce:
                        # the data gradient of this conv is the dy of layer k-1: accumulate that layer's BN+ReLU backward
                        # reduction (sum g, sum g*zhat) in this GEMM's epilogue instead of a separate pass over dy and z
                        rp = by_tag["%s.b%d.%d" % (tag, i, k - 1)]
                        sums = self._zeroed(rp["tag"] + ".bsums", ops.bn_bwd_sums_len(rp["z"].shape[3]), torch.float64)
                        bn_stats = (rp["z"], rp["scale"], rp["shift"], rp["mean"], rp["invstd"], sums)
                        sums_ready = True
                    ops.conv_dgrad(dz, W[r["conv"]], 3, r["stride"], dprev, bn_stats=bn_stats)
                    dy = dprev
                else:
                    ops.conv_dgrad(dz, W[r["conv"]], 3, r["stride"], dx_first, accumulate=accumulate_first)

        nlev = len(levels)
        for i in range(nlev - 1, 0, -1):
            block_bwd("B", i, d_levels[i], d_levels[i - 1], True)
        n_early = 0
        if sync is not None and sync.active() and self.use_side_stream:
            # everything but blocks.0 / PillarVFE is final once the re-layout of the weight gradients issued so far has run:
            # launch their all-reduce from the side stream (ordered after those GEMMs and, through the fork, after the BN /
            # bias gradients on the main stream), beside the level-0 backward
            with self._on_side():
                ops.unpack_wgrads_batched(self._job_table("unpack.early", unpack))
                sync.start()
            n_early = len(unpack)
        # level 0: d(x0m) -> mask -> d(x0) ; block 0 ran in pass "A" (shared)
        d_x0 = self._buf("bwd.d_x0", S["x0"].shape)
        ops.relu_bwd(d_levels[0], None, Act(d_x0), mask=S["mask"])
        d_canvas = self._buf("bwd.d_canvas", S["canvas"].shape)
        block_bwd("A", 0, d_x0, d_canvas, False)

        # ---- PillarVFE
        # agent types absent from this batch get a ZERO gradient (the persistent .grad buffers are overwritten, never
        # accumulated: without this a type's previous-step gradient would be applied again; the reference's zero_grad()
        # leaves those grads None and the optimizer skips them)
        present = {r["type"] for r in rec if r["kind"] == "pfn"}
        for t in AGENT_TYPES:
            if t in present:
                continue
            for suffix in (".linear.weight", ".norm.weight", ".norm.bias"):
                g = grads.get(TYPE_PREFIX[t] + ".0.0.pfn_layers.0" + suffix)
                if g is not None:
                    g.zero_()
        for r in rec:
            if r["kind"] != "pfn":
                continue
            pre = r["pre"]
            acc = self._buf("pfn.%s.acc" % r["type"], (64 * 12,), torch.float64)
            ops.pfn_bwd(r["vox"], r["num"], r["coords"], r["geom"], P[pre + ".linear.weight"], r["scale"], r["shift"],
                        r["mean"], r["invstd"], r["amap"], d_canvas, r["amax"], r["moments"], r["rows"], acc,
                        grads[pre + ".linear.weight"], grads[pre + ".norm.weight"], grads[pre + ".norm.bias"],
                        seg=r["seg"])
        with self._on_side():  # ordered after every weight-gradient GEMM (and, through the fork, the main stream so far)
            ops.unpack_wgrads_batched(self._job_table("unpack", unpack[n_early:]))
        if self.use_side_stream and self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        if sync is not None and sync.active():
            if n_early == 0:
                sync.start()
            sync.finish()
        return grads
