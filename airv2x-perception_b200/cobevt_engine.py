"""Host-side orchestration of the CoBEVT path (BASELINE config 4) on the C-ABI kernels:

    voxels -> PillarVFE+scatter -> BEV backbone (once) -> shrink -> regroup (pad to L agents)
           -> SwapFusionEncoder: depth x [LN -> window attention -> +res, LN -> FFN -> +res,
                                          LN -> grid attention   -> +res, LN -> FFN -> +res]
           -> mean over agents -> LN -> Linear -> detection heads

Mirrors opencood/models/airv2x_cobevt.py:112-156 and cobevt_modules/swap_fusion_modules.py:130-280. The encoder half
is shared with the Where2comm engine (same kernels); every nn.Linear is the 1x1 tcgen05 tap-GEMM (bf16x3 split, bias /
GELU / residual add fused in the epilogue), LayerNorm and the window / grid attention are token kernels
(csrc/transformer.cu). forward() is the eval path; forward_train() / backward_train() are the training step (dropout =
identity): saved activations per sublayer, 1x1 tap-GEMM dgrad / wgrad, LayerNorm / GELU / window-attention backward
kernels, then the shared encoder backward (_encoder_backward, also used by the V2X-ViT engine).
"""
import torch

from . import ops
from .ops import Act
from .w2c_engine import AGENT_TYPES, HEAD_PAD, TYPE_PREFIX, W2CEngine


class CoBEVTEngine(W2CEngine):
    def __init__(self, args, device, precision="split3"):  # noqa: the Where2comm-specific checks do not apply
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
        assert sh["use"] and list(sh["kernal_size"]) == [1] and list(sh["stride"]) == [1] and list(sh["padding"]) == [0], \
            "only the airv2x shrink header (1x1 s1 + 3x3) is implemented"
        self.compression = int(args.get("compression", 0) or 0)
        self.c_cat = sum(self.up_filters)
        self.c_shrink = sh["dim"][0]
        assert self.compression == 0 or (256 % self.compression == 0 and (256 // self.compression) % 64 == 0), \
            "NaiveCompressor: 256 / compression must be a multiple of 64 channels (compression in {1, 2, 4})"
        self.A = args["anchor_number"]
        self.K = args["num_class"]
        assert args["obj_head"], "obj_head: false not implemented"
        self.n_head = self.A * self.K + 7 * self.A + self.A
        fa = args["fax_fusion"]
        self.fa = fa
        self.L = sum(args["max_cav"].values())
        self.dim = fa["input_dim"]
        assert self.dim == self.c_shrink and self.dim % fa["dim_head"] == 0
        self.heads = self.dim // fa["dim_head"]
        self.bufs = {}
        self.saved = None
        self.side = None
        self.use_side_stream = False

    # shrink header geometry / head rows: overridden by the legacy (`point_pillar_*`) engines
    shrink_k0 = 1        # kernel of shrink_conv.layers.0.double_conv.0 (airv2x yaml: 1x1 s1; legacy yamls: 3x3 s2)
    shrink_stride = 1

    def _head_rows(self):
        nc, nr = self.A * self.K, 7 * self.A
        return (("cls_head", 0), ("reg_head", nc), ("obj_head", nc + nr))

    # ------------------------------------------------------------------ weights (one batched pack per step)
    def _linear_names(self):
        names = []
        for i in range(self.fa["depth"]):
            p = "fusion_net.layers.%d" % i
            for part in ("window", "grid"):
                names += ["%s.%s_attention.fn.to_qkv.weight" % (p, part), "%s.%s_attention.fn.to_out.0.weight" % (p, part),
                          "%s.%s_ffd.fn.net.0.weight" % (p, part), "%s.%s_ffd.fn.net.3.weight" % (p, part)]
        return names + ["fusion_net.mlp_head.3.weight"]

    @staticmethod
    def _compressor_convs():
        """(conv, its BatchNorm) of NaiveCompressor: encoder conv, two decoder convs (naive_compress.py:10-36)"""
        return {"naive_compressor.encoder.0": "naive_compressor.encoder.1",
                "naive_compressor.decoder.0": "naive_compressor.decoder.1",
                "naive_compressor.decoder.3": "naive_compressor.decoder.4"}

    def _compress(self, P, W, x):
        """NaiveCompressor.forward (eval): 3 x [conv3x3 + bias -> BatchNorm (running stats) -> ReLU]; the conv bias is
        folded into the BN shift, so each layer is one tap-GEMM with an affine + ReLU epilogue."""
        convs = self._compressor_convs()
        for li, (conv, bn) in enumerate(convs.items()):
            co = P[conv + ".weight"].shape[0]
            scale, shift = self._buf("cmp.scale%d" % li, (co,)), self._buf("cmp.shift%d" % li, (co,))
            ops.bn_eval_affine(P[bn + ".weight"], P[bn + ".bias"], P[bn + ".running_mean"], P[bn + ".running_var"],
                               scale, shift)
            shift.addcmul_(scale, P[conv + ".bias"])
            last = li == len(convs) - 1
            y = self._act("cmp.y%d" % li, x.shape[:3] + (co,), split=(False if last else None))
            ops.conv_fwd(x, W[conv], 3, 1, y, scale=scale, shift=shift, relu=True)
            x = y
        return x.hi

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
        for idx, k in ((0, self.shrink_k0), (2, 3)):
            name = "shrink_conv.layers.0.double_conv.%d.weight" % idx
            w = P[name]
            co, ci = w.shape[0], w.shape[1]
            W[name] = self._packed(name, (k * k, co, ci), (k * k, ci, co))
            jobs.append(ops.conv_pack_job(w, W[name], f32=not self.split))
        if self.compression:
            for name in self._compressor_convs():
                w = P[name + ".weight"]
                co, ci = w.shape[0], w.shape[1]
                W[name] = self._packed(name, (9, co, ci), (9, ci, co))
                jobs.append(ops.conv_pack_job(w, W[name], f32=not self.split))
        for name in self._linear_names():  # nn.Linear [out, in] == 1x1 conv OIHW [out, in, 1, 1]
            w = P[name]
            co, ci = w.shape
            W[name] = self._packed(name, (1, co, ci), (1, ci, co))
            jobs.append(ops.conv_pack_job(w.view(co, ci, 1, 1), W[name], f32=not self.split))
        fresh = ("packed", "heads") not in self.bufs
        hp = self._packed("heads", (1, HEAD_PAD, self.c_shrink), (1, self.c_shrink, HEAD_PAD))
        hb = self._buf("heads.b", (HEAD_PAD,))
        if fresh:
            for t in (hp.f32, hp.f16, hp.d32, hp.d16, hb):
                t.zero_()
        for name, row0 in self._head_rows():
            jobs.append(ops.conv_pack_job(P[name + ".weight"], hp, row0, f32=not self.split))
            jobs.append(ops.copy_pack_job(P[name + ".bias"], hb, row0))
        W["heads"] = hp
        W["heads.bias"] = hb
        ops.pack_weights_batched(self._job_table("pack", jobs))
        return W

    # ------------------------------------------------------------------ fusion network
    def _attention(self, P, W, pre, X, key_mask, B, grid_mode, tag):
        """x += to_out(attention(LN(x)))    (PreNormResidual(Attention), swap_fusion_modules.py:78-127)"""
        n, h, w, d = X.shape
        ln = self._act("fax.ln", X.shape)
        ops.layernorm_fwd(X, P[pre + ".norm.weight"], P[pre + ".norm.bias"], ln)
        qkv = self._buf("fax.qkv", (n, h, w, 3 * d))
        ops.linear_fwd(ln, W[pre + ".fn.to_qkv.weight"], Act(qkv))
        att = self._act("fax.att", X.shape)
        ops.window_attention_fwd(qkv, P[pre + ".fn.relative_position_bias_table.weight"], key_mask, B, self.L, self.heads,
                                 self.fa["dim_head"], self.fa["window_size"], grid_mode, att)
        ops.linear_fwd(att, W[pre + ".fn.to_out.0.weight"], Act(X), accumulate=True)

    def _ffn(self, P, W, pre, X):
        """x += W2 gelu(W1 LN(x) + b1) + b2   (PreNormResidual(FeedForward), base_transformer.py:16-28)"""
        ln = self._act("fax.ln", X.shape)
        ops.layernorm_fwd(X, P[pre + ".norm.weight"], P[pre + ".norm.bias"], ln)
        hid = self._act("fax.hid", X.shape[:3] + (self.fa["mlp_dim"],))
        ops.linear_fwd(ln, W[pre + ".fn.net.0.weight"], hid, bias=P[pre + ".fn.net.0.bias"], act=2)
        ops.linear_fwd(hid, W[pre + ".fn.net.3.weight"], Act(X), bias=P[pre + ".fn.net.3.bias"], accumulate=True)

    def fusion(self, P, W, feat, layout, peer_ptrs=None):
        """feat: dense [N, h, w, C] shrunk maps (scene-major). Returns Act [B, h, w, C].
        peer_ptrs (agent-parallel mode): int64 device table of per-agent map pointers in (peer) GPU memory; the maps
        are then pulled over NVLink inside the regroup kernel and `feat` only supplies the shape."""
        B = len(layout["record_len"])
        n, h, w, d = feat.shape
        X = self._buf("fax.x", (B * self.L, h, w, d))
        if peer_ptrs is not None:
            ops.regroup_ptrs(peer_ptrs, (h, w, d), layout["scene_start"], layout["scene_len"], B, self.L, Act(X))
        else:
            ops.regroup(feat, layout["scene_start"], layout["scene_len"], B, self.L, Act(X))
        key_mask = layout["key_mask"] if self.fa.get("mask", False) else None
        for i in range(self.fa["depth"]):
            p = "fusion_net.layers.%d" % i
            self._attention(P, W, p + ".window_attention", X, key_mask, B, False, "w%d" % i)
            self._ffn(P, W, p + ".window_ffd", X)
            self._attention(P, W, p + ".grid_attention", X, key_mask, B, True, "g%d" % i)
            self._ffn(P, W, p + ".grid_ffd", X)
        m = self._act("fax.mean", (B, h, w, d))
        ops.agent_mean_layernorm(X, B, self.L, P["fusion_net.mlp_head.2.weight"], P["fusion_net.mlp_head.2.bias"], m)
        fused = self._act("fax.fused", (B, h, w, d))
        ops.linear_fwd(m, W["fusion_net.mlp_head.3.weight"], fused, bias=P["fusion_net.mlp_head.3.bias"])
        return fused

    # ------------------------------------------------------------------ forward
    def encode(self, P, W, lidar, layout, out=None):
        """per-agent half of the path: voxels -> PillarVFE+scatter -> backbone -> shrink. Returns the dense
        [N, h/2, w/2, C] shrunk maps (written into `out` if given, e.g. a symmetric-memory buffer peers read)."""
        N = layout["n_total"]
        canvas = self._encode(P, lidar, layout, False, None)
        self._last_canvas = canvas
        x = canvas
        h2 = w2 = None
        cat = None
        for i in range(len(self.layer_nums)):
            x = self._block(P, W, i, x, False, 0, "E", None, need_hi=False)
            if cat is None:
                h2, w2 = x.shape[1], x.shape[2]
                cat = self._act("E.cat", (N, h2, w2, self.c_cat))
            c0 = sum(self.up_filters[:i])
            self._deblock(P, W, i, x, cat.slice_c(c0, c0 + self.up_filters[i]), False, 0, "E", None)
        s = self.shrink_stride
        h2, w2 = (h2 - 1) // s + 1, (w2 - 1) // s + 1
        y1 = self._act("E.s1", (N, h2, w2, self.c_shrink))
        if self.compression:
            y2a = self._act("E.s2c", (N, h2, w2, self.c_shrink))
        else:
            y2 = out if out is not None else self._buf("E.s2", (N, h2, w2, self.c_shrink))
            assert tuple(y2.shape) == (N, h2, w2, self.c_shrink)
            y2a = Act(y2)
        ops.conv_fwd(cat, W["shrink_conv.layers.0.double_conv.0.weight"], self.shrink_k0, s, y1,
                     shift=P["shrink_conv.layers.0.double_conv.0.bias"], relu=True)
        ops.conv_fwd(y1, W["shrink_conv.layers.0.double_conv.2.weight"], 3, 1, y2a,
                     shift=P["shrink_conv.layers.0.double_conv.2.bias"], relu=True)
        if self.compression:  # airv2x_cobevt.py:121-123
            y2 = self._compress(P, W, y2a)
            if out is not None:
                out.copy_(y2)
                y2 = out
        return y2

    def fuse_heads(self, P, W, feat, layout, peer_ptrs=None):
        fused = self.fusion(P, W, feat, layout, peer_ptrs)
        heads = self._buf("heads.out", (fused.shape[0], feat.shape[1], feat.shape[2], HEAD_PAD))
        ops.linear_fwd(fused, W["heads"], Act(heads), bias=W["heads.bias"])
        return heads

    # ------------------------------------------------------------------ training step (dropout disabled)
    def _sub(self, i, part, kind):
        return "fusion_net.layers.%d.%s_%s" % (i, part, kind)

    def _encode_train(self, P, W, lidar, layout, rec):
        """train-mode encoder (batch-statistic BatchNorm, everything the backward needs recorded in `rec`):
        voxels -> PillarVFE+scatter -> backbone -> shrink. Returns (y1 Act, y2 fp32 [N,h,w,C], cat Act)."""
        N = layout["n_total"]
        canvas = self._encode(P, lidar, layout, True, rec)
        self._last_canvas_shape = tuple(canvas.shape)
        self._last_canvas = canvas
        x = canvas
        cat = None
        for i in range(len(self.layer_nums)):
            x = self._block(P, W, i, x, True, 1, "E", rec, need_hi=False)  # feeds the next block and the deblock only
            if cat is None:
                h2, w2 = x.shape[1], x.shape[2]
                cat = self._act("E.cat", (N, h2, w2, self.c_cat))
            c0 = sum(self.up_filters[:i])
            with self._on_side():
                self._deblock(P, W, i, x, cat.slice_c(c0, c0 + self.up_filters[i]), True, 1, "E", rec)
        self._join_side()
        s = self.shrink_stride      # airv2x yamls: 1x1 stride 1; legacy yamls: 3x3 stride 2 (or 1)
        h2, w2 = (h2 - 1) // s + 1, (w2 - 1) // s + 1
        y1 = self._act("E.s1", (N, h2, w2, self.c_shrink))
        ops.conv_fwd(cat, W["shrink_conv.layers.0.double_conv.0.weight"], self.shrink_k0, s, y1,
                     shift=P["shrink_conv.layers.0.double_conv.0.bias"], relu=True)
        if not self.compression:
            y2 = self._buf("E.s2", (N, h2, w2, self.c_shrink))
            ops.conv_fwd(y1, W["shrink_conv.layers.0.double_conv.2.weight"], 3, 1, Act(y2),
                         shift=P["shrink_conv.layers.0.double_conv.2.bias"], relu=True)
            return y1, y2, cat
        # NaiveCompressor in train mode (naive_compress.py:10-42): 3 x [conv3x3 + bias -> BatchNorm(batch statistics) ->
        # ReLU]. The batch mean absorbs the conv bias, so the layer is the bias-free conv + BN of the backbone blocks (the
        # bias gradient is exactly zero); only the running mean sees the bias: momentum * bias on top of the bias-free update.
        y2a = self._act("E.s2c", (N, h2, w2, self.c_shrink))
        ops.conv_fwd(y1, W["shrink_conv.layers.0.double_conv.2.weight"], 3, 1, y2a,
                     shift=P["shrink_conv.layers.0.double_conv.2.bias"], relu=True)
        x = y2a
        convs = self._compressor_convs()
        for li, (conv, bn) in enumerate(convs.items()):
            x = self._conv_bn_relu(P, {conv + ".weight": W[conv]}, conv + ".weight", bn, x, 1, True, 1, "cmp%d" % li, rec,
                                   need_hi=li == len(convs) - 1)
            P[bn + ".running_mean"].add_(P[conv + ".bias"], alpha=0.01)   # BatchNorm2d(momentum=0.01), naive_compress.py:20
        self._cmp_in = y2a
        return y1, x.hi, cat

    def forward_train(self, P, lidar, layout, drop=None):
        """Train-mode forward (batch-statistic BatchNorm in the encoder) that keeps what the backward needs: per sublayer
        the residual input, the LayerNorm output (GEMM operand of the weight gradients), the qkv tensor / the attention
        output, the FFN pre-activation and hidden activation. drop: ops.Dropout (rate fax_fusion.drop_out, this step's
        seed) or None = nn.Dropout disabled. Sites, in call order per sublayer: Attention.to_out's Dropout
        (swap_fusion_modules.py:43), FeedForward's two Dropouts (base_transformer.py:32,34); masks are regenerated from
        (seed, site) by the backward, never stored."""
        self._begin_step()
        rec = []
        W = self._pack_weights(P)
        y1, y2, cat = self._encode_train(P, W, lidar, layout, rec)
        h2, w2 = y2.shape[1], y2.shape[2]
        B = len(layout["record_len"])
        d = self.dim
        X = self._buf("fax.x", (B * self.L, h2, w2, d))
        ops.regroup(y2, layout["scene_start"], layout["scene_len"], B, self.L, Act(X))
        key_mask = layout["key_mask"] if self.fa.get("mask", False) else None
        subs = []
        for i in range(self.fa["depth"]):
            for part, grid_mode in (("window", False), ("grid", True)):
                pa, pf = self._sub(i, part, "attention"), self._sub(i, part, "ffd")
                tag = "%d%s" % (i, part[0])
                # attention sublayer: x_new = x + dropout(to_out(att)) in the GEMM epilogue, written to a fresh buffer so that
                # the sublayer's input (LayerNorm backward needs it) is kept without a copy
                ln = self._act("sv.ln.a" + tag, X.shape)
                ops.layernorm_fwd(X, P[pa + ".norm.weight"], P[pa + ".norm.bias"], ln)
                qkv = self._buf("sv.qkv." + tag, X.shape[:3] + (3 * d,))
                ops.linear_fwd(ln, W[pa + ".fn.to_qkv.weight"], Act(qkv))
                att = self._act("sv.att." + tag, X.shape)
                ops.window_attention_fwd(qkv, P[pa + ".fn.relative_position_bias_table.weight"], key_mask, B, self.L,
                                         self.heads, self.fa["dim_head"], self.fa["window_size"], grid_mode, att)
                site_o = drop.site() if drop is not None else None
                Xn = self._buf("sv.x.a" + tag, X.shape)
                ops.linear_dropout_residual_fwd(att, W[pa + ".fn.to_out.0.weight"], Xn, residual=X, drop=drop, site=site_o or 0)
                subs.append(dict(kind="att", pre=pa, xin=X, ln=ln, qkv=qkv, att=att, grid=grid_mode, site_o=site_o))
                X = Xn
                # feed-forward sublayer (pre-activation kept in fp32: GELU' needs it): x_new = x + dropout(W2 dropout(gelu(pre)) + b2)
                ln = self._act("sv.ln.f" + tag, X.shape)
                ops.layernorm_fwd(X, P[pf + ".norm.weight"], P[pf + ".norm.bias"], ln)
                pre = self._buf("sv.pre." + tag, X.shape[:3] + (self.fa["mlp_dim"],))
                ops.linear_fwd(ln, W[pf + ".fn.net.0.weight"], Act(pre), bias=P[pf + ".fn.net.0.bias"])
                hid = self._act("sv.hid." + tag, pre.shape)
                site_h = site_o = None
                if drop is None:
                    ops.gelu_fwd(pre, hid)
                else:
                    site_h = drop.site()
                    ops.gelu_dropout_fwd(pre, drop, site_h, hid)
                    site_o = drop.site()
                Xn = self._buf("sv.x.f" + tag, X.shape)
                ops.linear_dropout_residual_fwd(hid, W[pf + ".fn.net.3.weight"], Xn, bias=P[pf + ".fn.net.3.bias"], residual=X,
                                                drop=drop, site=site_o or 0)
                subs.append(dict(kind="ffn", pre=pf, xin=X, ln=ln, hpre=pre, hid=hid, site_h=site_h, site_o=site_o))
                X = Xn
        m = self._act("fax.mean", (B, h2, w2, d))
        ops.agent_mean_layernorm(X, B, self.L, P["fusion_net.mlp_head.2.weight"], P["fusion_net.mlp_head.2.bias"], m)
        fused = self._act("fax.fused", (B, h2, w2, d))
        ops.linear_fwd(m, W["fusion_net.mlp_head.3.weight"], fused, bias=P["fusion_net.mlp_head.3.bias"])
        heads = self._buf("heads.out", (B, h2, w2, HEAD_PAD))
        ops.linear_fwd(fused, W["heads"], Act(heads), bias=W["heads.bias"])
        self.saved = dict(rec=rec, W=W, subs=subs, X=X, m=m, fused=fused, y1=y1, y2=y2, cat=cat, layout=layout,
                          key_mask=key_mask, B=B, drop=drop, cmp_in=self._cmp_in if self.compression else None)
        return heads

    def backward_train(self, P, dheads, grads):
        """dheads: [B,h,w,HEAD_PAD] gradient w.r.t. the head logits; grads: name -> fp32 tensor (written)."""
        S = self.saved
        W, rec, layout, B = S["W"], S["rec"], S["layout"], S["B"]
        nc, nr = self.A * self.K, 7 * self.A
        d = self.dim
        unpack = []

        def zero_f32(name, n):
            return self._zeroed(name, n, torch.float32)

        def lin_wgrad(x_act, dy_act, wname):
            """grad of nn.Linear weight [out, in] = 1x1 conv wgrad"""
            co, ci = dy_act.shape[3], x_act.shape[3]
            dwp = zero_f32(wname + ".dwp", co * ci).view(1, co, ci)
            ops.conv_wgrad(x_act, dy_act, 1, 1, dwp)
            g = grads[wname]
            unpack.append(ops.conv_unpack_job(dwp, g.view(co, ci, 1, 1)))

        def col_sums(t, C, outs):
            sums = self._zeroed("bias.sums", 2 * C, torch.float64)
            ops.channel_stats(t, sums)
            for out, c0 in outs:
                unpack.append(ops.sums_unpack_job(sums, out, c0))

        def split_of(t, name):
            a = self._act(name, t.shape)
            ops.affine_act(t, None, None, False, a, write_hi=False)  # GEMM operand only
            return a

        def ln_bwd(xin, d_ln, pre_norm, dX):
            acc = self._zeroed(pre_norm + ".lnacc", 2 * d, torch.float64)
            ops.layernorm_bwd(xin, d_ln, P[pre_norm + ".weight"], dX, acc[:d], acc[d:])
            unpack.append(ops.sums_unpack_job(acc, grads[pre_norm + ".weight"], 0))
            unpack.append(ops.sums_unpack_job(acc, grads[pre_norm + ".bias"], d))

        # ---- heads, mlp_head
        dh = split_of(dheads, "bwd.dheads")
        dwp = zero_f32("heads.dwp", HEAD_PAD * self.c_shrink).view(1, HEAD_PAD, self.c_shrink)
        ops.conv_wgrad(S["fused"], dh, 1, 1, dwp)
        for name, row0 in self._head_rows():
            unpack.append(ops.conv_unpack_job(dwp, grads[name + ".weight"], row0))
        col_sums(dheads, HEAD_PAD, [(grads[name + ".bias"], row0) for name, row0 in self._head_rows()])
        d_fused = self._buf("bwd.d_fused", S["fused"].shape)
        ops.conv_dgrad(dh, W["heads"], 1, 1, d_fused)
        dfs = split_of(d_fused, "bwd.dfs")
        lin_wgrad(S["m"], dfs, "fusion_net.mlp_head.3.weight")
        col_sums(d_fused, d, [(grads["fusion_net.mlp_head.3.bias"], 0)])
        d_m = self._buf("bwd.d_m", S["m"].shape)
        ops.conv_dgrad(dfs, W["fusion_net.mlp_head.3.weight"], 1, 1, d_m)
        # mean over agents + LayerNorm (swap_fusion_modules.py:269-273)
        X = S["X"]
        xbar = self._buf("bwd.xbar", S["m"].shape)
        ops.agent_mean(X, B, self.L, xbar)
        d_xbar = self._buf("bwd.d_xbar", S["m"].shape)
        d_xbar.zero_()
        ln_bwd(xbar, d_m, "fusion_net.mlp_head.2", d_xbar)
        dX = self._buf("bwd.dX", X.shape)
        ops.agent_broadcast(d_xbar, B, self.L, 1.0 / self.L, dX)
        # ---- sublayers in reverse
        drop = S["drop"]
        for sl in reversed(S["subs"]):
            pre = sl["pre"]
            if drop is None:
                dXs, g_out = split_of(dX, "bwd.dXs"), dX
            else:   # gradient w.r.t. the sublayer's last linear = stream gradient through that site's dropout mask
                dXs = self._act("bwd.dXs", dX.shape)
                # the fp32 plane is only read by the bias gradient of the feed-forward's second linear (to_out has no bias)
                ops.dropout_apply(dX, drop, sl["site_o"], dXs, write_hi=sl["kind"] == "ffn")
                g_out = dXs.hi
            if sl["kind"] == "ffn":
                lin_wgrad(sl["hid"], dXs, pre + ".fn.net.3.weight")
                col_sums(g_out, d, [(grads[pre + ".fn.net.3.bias"], 0)])
                d_hid = self._buf("bwd.d_hid", sl["hid"].shape)
                ops.conv_dgrad(dXs, W[pre + ".fn.net.3.weight"], 1, 1, d_hid)
                d_pre = self._act("bwd.d_pre", sl["hid"].shape)
                if drop is None:
                    ops.gelu_bwd(d_hid, sl["hpre"], d_pre)
                else:
                    ops.gelu_dropout_bwd(d_hid, sl["hpre"], drop, sl["site_h"], d_pre)
                lin_wgrad(sl["ln"], d_pre, pre + ".fn.net.0.weight")
                col_sums(d_pre.hi, d_pre.shape[3], [(grads[pre + ".fn.net.0.bias"], 0)])
                d_ln = self._buf("bwd.d_ln", X.shape)
                ops.conv_dgrad(d_pre, W[pre + ".fn.net.0.weight"], 1, 1, d_ln)
            else:
                lin_wgrad(sl["att"], dXs, pre + ".fn.to_out.0.weight")
                d_att = self._buf("bwd.d_att", X.shape)
                ops.conv_dgrad(dXs, W[pre + ".fn.to_out.0.weight"], 1, 1, d_att)
                dqs = self._act("bwd.dqs", sl["qkv"].shape)   # written as a GEMM operand: no fp32 plane, no conversion pass
                tname = pre + ".fn.relative_position_bias_table.weight"
                grads[tname].zero_()
                ops.window_attention_bwd(sl["qkv"], d_att, P[tname], S["key_mask"], B, self.L, self.heads,
                                         self.fa["dim_head"], self.fa["window_size"], sl["grid"], dqs, grads[tname],
                                         write_hi=False)
                lin_wgrad(sl["ln"], dqs, pre + ".fn.to_qkv.weight")
                d_ln = self._buf("bwd.d_ln", X.shape)
                ops.conv_dgrad(dqs, W[pre + ".fn.to_qkv.weight"], 1, 1, d_ln)
            ln_bwd(sl["xin"], d_ln, pre + ".norm", dX)
        # ---- regroup^T: gradients of the valid agents' shrunk maps (padded slots are dropped)
        d_y2 = self._buf("bwd.d_y2", S["y2"].shape)
        pos = 0
        for b, n in enumerate(layout["record_len"]):
            d_y2[pos:pos + n].copy_(dX[b * self.L:b * self.L + n])
            pos += n
        self._encoder_backward(P, S, d_y2, grads, unpack)
        for lo in range(0, len(unpack), 128):
            ops.unpack_wgrads_batched(self._job_table("unpack%d" % lo, unpack[lo:lo + 128]))
        return grads


    # ------------------------------------------------------------------ shared encoder backward
    def _encoder_backward(self, P, S, d_y2, grads, unpack):
        """d_y2: gradient w.r.t. the shrunk maps [N, h, w, C] -> shrink header, deblocks, blocks, PillarVFE. Appends the
        weight-gradient re-layout jobs to `unpack` (flushed by the caller)."""
        W, rec = S["W"], S["rec"]
        y2, y1, cat = S["y2"], S["y1"], S["cat"]

        def zero_f32(name, n):
            return self._zeroed(name, n, torch.float32)

        def col_sums(t, C, outs):
            sums = self._zeroed("bias.sums", 2 * C, torch.float64)
            ops.channel_stats(t, sums)
            for out, c0 in outs:
                unpack.append(ops.sums_unpack_job(sums, out, c0))

        if S.get("cmp_in") is not None:
            # ---- NaiveCompressor, last layer first: d_y2 is the gradient w.r.t. its output; y2 becomes the shrink output
            dy = d_y2
            for li in (2, 1, 0):
                r = next(q for q in rec if q["kind"] == "conv" and q["tag"] == "cmp%d" % li)
                conv = r["conv"][:-len(".weight")]
                dz = self._act("bwd.dz." + r["tag"], r["z"].shape)
                sums = self._zeroed(r["tag"] + ".bsums", ops.bn_bwd_sums_len(r["z"].shape[3]), torch.float64)
                ops.bn_relu_bwd(dy, r["z"], r["scale"], r["shift"], r["mean"], r["invstd"], sums, dz,
                                grads[r["bn"] + ".weight"], grads[r["bn"] + ".bias"], write_hi=False)
                cout, cin = r["z"].shape[3], r["x"].shape[3]
                dwp = zero_f32(conv + ".dwp", 9 * cout * cin).view(9, cout, cin)
                ops.conv_wgrad(r["x"], dz, 3, 1, dwp)
                unpack.append(ops.conv_unpack_job(dwp, grads[r["conv"]]))
                grads[conv + ".bias"].zero_()   # absorbed by the batch mean
                dprev = self._buf("bwd.dprev." + r["tag"], r["x"].shape)
                ops.conv_dgrad(dz, W[conv], 3, 1, dprev)
                dy = dprev
            d_y2, y2 = dy, S["cmp_in"].hi
        # ---- shrink header
        g2 = self._act("bwd.g2", y2.shape)
        ops.relu_bwd(d_y2, y2, g2)  # y2 is a plain fp32 tensor here
        n2, n1 = "shrink_conv.layers.0.double_conv.2", "shrink_conv.layers.0.double_conv.0"
        dwp = zero_f32(n2 + ".dwp", 9 * self.c_shrink * self.c_shrink).view(9, self.c_shrink, self.c_shrink)
        ops.conv_wgrad(y1, g2, 3, 1, dwp)
        unpack.append(ops.conv_unpack_job(dwp, grads[n2 + ".weight"]))
        col_sums(g2.hi, self.c_shrink, [(grads[n2 + ".bias"], 0)])
        d_y1 = self._buf("bwd.d_y1", y1.shape)
        ops.conv_dgrad(g2, W[n2 + ".weight"], 3, 1, d_y1)
        g1 = self._act("bwd.g1", y1.shape)
        ops.relu_bwd(d_y1, y1.hi, g1)
        k0, s0 = self.shrink_k0, self.shrink_stride
        dwp = zero_f32(n1 + ".dwp", k0 * k0 * self.c_shrink * self.c_cat).view(k0 * k0, self.c_shrink, self.c_cat)
        ops.conv_wgrad(cat, g1, k0, s0, dwp)
        unpack.append(ops.conv_unpack_job(dwp, grads[n1 + ".weight"]))
        col_sums(g1.hi, self.c_shrink, [(grads[n1 + ".bias"], 0)])
        d_cat = self._buf("bwd.d_cat", cat.shape)
        ops.conv_dgrad(g1, W[n1 + ".weight"], k0, s0, d_cat)
        # ---- backbone: deblocks, then blocks deep -> shallow
        by_tag = {r["tag"]: r for r in rec if r["kind"] == "conv"}
        deconvs = {r["level"]: r for r in rec if r["kind"] == "deconv"}
        nlev = len(self.layer_nums)
        d_x = [None] * nlev
        for i in range(nlev):
            r = deconvs[i]
            c0 = sum(self.up_filters[:i])
            dy = d_cat[..., c0:c0 + self.up_filters[i]]
            dz = self._act("bwd.dz.d%d" % i, r["z"].shape)
            sums = self._zeroed(r["tag"] + ".bsums", ops.bn_bwd_sums_len(r["z"].shape[3]), torch.float64)
            ops.bn_relu_bwd(dy, r["z"], r["scale"], r["shift"], r["mean"], r["invstd"], sums, dz,
                            grads[r["bn"] + ".weight"], grads[r["bn"] + ".bias"], write_hi=False)
            s = r["stride"]
            cin, cout = r["x"].shape[3], r["z"].shape[3]
            dwp = zero_f32(r["conv"] + ".dwp", s * s * cin * cout).view(s * s, cin, cout)
            ops.deconv_wgrad(r["x"], dz, s, dwp)
            unpack.append(ops.deconv_unpack_job(dwp, grads[r["conv"]]))
            d_x[i] = self._buf("bwd.dx%d" % i, r["x"].shape)
            ops.deconv_dgrad(dz, W[r["conv"]], s, d_x[i])
        d_canvas = self._buf("bwd.d_canvas", self._last_canvas_shape)
        for i in range(nlev - 1, -1, -1):
            dy = d_x[i]
            for k in range(self.layer_nums[i], -1, -1):
                r = by_tag["E.b%d.%d" % (i, k)]
                dz = self._act("bwd.dz." + r["tag"], r["z"].shape)
                sums = self._zeroed(r["tag"] + ".bsums", ops.bn_bwd_sums_len(r["z"].shape[3]), torch.float64)
                ops.bn_relu_bwd(dy, r["z"], r["scale"], r["shift"], r["mean"], r["invstd"], sums, dz,
                                grads[r["bn"] + ".weight"], grads[r["bn"] + ".bias"], write_hi=False)
                cout, cin = r["z"].shape[3], r["x"].shape[3]
                dwp = zero_f32(r["conv"] + ".dwp", 9 * cout * cin).view(9, cout, cin)
                ops.conv_wgrad(r["x"], dz, 3, r["stride"], dwp)
                unpack.append(ops.conv_unpack_job(dwp, grads[r["conv"]]))
                if k > 0:
                    dprev = self._buf("bwd.dprev.b%d.%d" % (i, k), r["x"].shape)
                    ops.conv_dgrad(dz, W[r["conv"]], 3, r["stride"], dprev)
                    dy = dprev
                elif i > 0:
                    ops.conv_dgrad(dz, W[r["conv"]], 3, r["stride"], d_x[i - 1], accumulate=True)
                else:
                    ops.conv_dgrad(dz, W[r["conv"]], 3, r["stride"], d_canvas)
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
                        grads[pre + ".linear.weight"], grads[pre + ".norm.weight"], grads[pre + ".norm.bias"], seg=r["seg"])

    def forward(self, P, lidar, layout, training, k_list=None):
        if training:
            raise NotImplementedError("Airv2xCoBEVT: train-mode forward under torch.no_grad() is not implemented (model.eval() "
                                      "for inference; model(batch) with grad enabled or train_step() for training)")
        self._begin_step()
        W = self._pack_weights(P)
        feat = self.encode(P, W, lidar, layout)
        return self.fuse_heads(P, W, feat, layout), {}
