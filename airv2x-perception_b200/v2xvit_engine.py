"""Host-side orchestration of the V2X-ViT path (BASELINE config 3) on the C-ABI kernels:

    voxels -> PillarVFE+scatter -> BEV backbone -> shrink -> (valid agents only) -> RTE -> STTF ego-warp -> ROI mask
           -> depth x [ LN -> HGT multi-agent attention -> +res, LN -> pyramid window attention + split-attn -> +res,
                        LN -> FFN -> +res ]
           -> ego map -> detection heads

Mirrors opencood/models/airv2x_v2xvit.py:108-167 and v2xvit_modules/v2xvit_basic.py:135-213. Padded agents are never
attention keys and only agent 0 is returned, so the kernels run on the valid agents only (exact; the reference spends
2/3 of its 4.6 TFLOP on padding). Every nn.Linear is the 1x1 tcgen05 tap-GEMM (bf16x3 split); the HGT relation tensors
are folded into the K / V projections once per step. forward() is the eval path; forward_train() / backward_train()
are the training step (dropout = identity).
"""
import torch

from . import ops, warp
from .cobevt_engine import CoBEVTEngine
from .ops import Act
from .w2c_engine import HEAD_PAD


class V2XViTEngine(CoBEVTEngine):
    def __init__(self, args, device, precision="split3"):  # noqa
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
        sh = mf["shrink_header"]
        assert sh["use"] and list(sh["kernal_size"]) == [1] and list(sh["stride"]) == [1] and list(sh["padding"]) == [0], \
            "only the airv2x shrink header (1x1 s1 + 3x3) is implemented"
        assert not mf.get("compression", 0), "NaiveCompressor (compression > 0) is not implemented"
        self.compression = 0
        self.c_cat = sum(self.up_filters)
        self.c_shrink = sh["dim"][0]
        self.A = args["anchor_number"]
        self.K = args["num_class"]
        assert args["obj_head"], "obj_head: false not implemented"
        self.n_head = self.A * self.K + 7 * self.A + self.A
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
        self.L = sum(args["max_cav"].values())
        self.bufs = {}
        self.saved = None
        self.side = None
        self.use_side_stream = False

    # ------------------------------------------------------------------ weights
    def _names(self, d):
        lp = "fusion_net.encoder.layers.%d" % d
        bp = lp + ".0.layers.0"
        return lp, bp, bp + ".0", bp + ".1"

    def _pack_weights(self, P):
        C = self.dim
        heads = self.enc["cav_att_config"]["heads"]
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

        def lin(name, w):
            co, ci = w.shape
            W[name] = self._packed(name, (1, co, ci), (1, ci, co))
            jobs.append(ops.conv_pack_job(w.view(co, ci, 1, 1), W[name], f32=not self.split))

        for d in range(self.enc["depth"]):
            lp, bp, hg, pwp = self._names(d)
            f = hg + ".fn"
            # fold relation_att / relation_msg into the typed K / V projections (one small launch per layer)
            wf = self._buf("hgt.wfold.%d" % d, (2, 5 * C, C))
            bf = self._buf("hgt.bfold.%d" % d, (2, 5 * C))
            pair = lambda n, s: (P["%s.%s.0.%s" % (f, n, s)], P["%s.%s.1.%s" % (f, n, s)])
            ops.hgt_fold(pair("q_linears", "weight"), pair("q_linears", "bias"), pair("k_linears", "weight"),
                         pair("k_linears", "bias"), pair("v_linears", "weight"), pair("v_linears", "bias"),
                         P[f + ".relation_att"], P[f + ".relation_msg"], heads, wf, bf)
            for t in range(2):
                lin("%s.fold.%d" % (f, t), wf[t])
                W["%s.fold.%d.bias" % (f, t)] = bf[t]
                lin("%s.a_linears.%d.weight" % (f, t), P["%s.a_linears.%d.weight" % (f, t)])
            for lv in range(3):
                lin("%s.fn.pwmsa.%d.to_qkv.weight" % (pwp, lv), P["%s.fn.pwmsa.%d.to_qkv.weight" % (pwp, lv)])
                lin("%s.fn.pwmsa.%d.to_out.0.weight" % (pwp, lv), P["%s.fn.pwmsa.%d.to_out.0.weight" % (pwp, lv)])
            lin(lp + ".1.fn.net.0.weight", P[lp + ".1.fn.net.0.weight"])
            lin(lp + ".1.fn.net.3.weight", P[lp + ".1.fn.net.3.weight"])
        if self.compression:
            for name in self._compressor_convs():
                w = P[name + ".weight"]
                co, ci = w.shape[0], w.shape[1]
                W[name] = self._packed(name, (9, co, ci), (9, ci, co))
                jobs.append(ops.conv_pack_job(w, W[name], f32=not self.split))
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
    @staticmethod
    def _runs(types):
        """maximal runs of consecutive agents of one type: [(start, stop, type)]"""
        runs, s = [], 0
        for i in range(1, len(types) + 1):
            if i == len(types) or types[i] != types[s]:
                runs.append((s, i, types[s]))
                s = i
        return runs

    def fusion(self, P, W, feat, layout, prior, scm):
        """feat: dense [N, h, w, C] shrunk maps of the VALID agents (scene-major, ego first per scene).
        prior: [B, L, 3] host/device (velocity, time delay, infra type); scm: [B, L, 4, 4]. Returns Act [B, h, w, C]."""
        enc = self.enc
        ca, pw = enc["cav_att_config"], enc["pwindow_att_config"]
        record_len = layout["record_len"]
        B, N = len(record_len), feat.shape[0]
        _, h, w, C = feat.shape
        dev = self.device
        prior = prior.detach().cpu().float()
        starts = [sum(record_len[:b]) for b in range(B)]
        valid = [(b, l) for b in range(B) for l in range(record_len[b])]
        types = [int(prior[b, l, 2]) for b, l in valid]
        assert all(t in (0, 1) for t in types), "prior_encoding[..., 2] (agent type) must be 0 or 1 (hmsa.py:8)"
        types_dev = torch.tensor(types, dtype=torch.int32, device=dev)
        # ---- RTE (v2xvit_basic.py:41-80): x[a] += Linear(emb[int(dt) * ratio])
        X = self._buf("vit.x", (N, h, w, C))
        X.copy_(feat)
        if ca["use_RTE"]:
            idx = torch.tensor([int(prior[b, l, 1]) * ca["RTE_ratio"] for b, l in valid], dtype=torch.int32, device=dev)
            rp = "fusion_net.encoder.rte.emb"
            ops.rte_add(X, P[rp + ".emb.weight"], idx, P[rp + ".lin.weight"], P[rp + ".lin.bias"], self._buf("vit.rte", (N, C)))
        # ---- STTF (v2xvit_basic.py:17-38): non-ego maps are resampled by the spatial correction, the ego map is kept
        dr, ds = enc["sttf"]["voxel_size"][0], enc["sttf"]["downsample_rate"]
        theta_all = warp.sttf_theta(scm, dr, ds, h, w)
        theta = torch.stack([theta_all[b, l] for b, l in valid]).to(dev)
        Xw = self._buf("vit.xw", (N, h, w, C))
        ops.warp_affine_fwd(X, theta, Act(Xw), align_corners=True)
        for s in starts:
            Xw[s].copy_(X[s])
        X = Xw
        # ---- ROI + agent mask (torch_transformation_utils.py:15-113): per pixel, per key agent
        kmask = self._buf("vit.kmask", (N, h, w))
        if enc["use_roi_mask"]:
            ops.roi_mask(theta, None, N, h, w, kmask, align_corners=True)
        else:
            kmask.fill_(1.0)
        runs = self._runs(types)
        for d in range(enc["depth"]):
            lp, bp, hg, pwp = self._names(d)
            f = hg + ".fn"
            # HGT multi-agent attention
            ln = self._act("vit.ln", X.shape)
            ops.layernorm_fwd(X, P[hg + ".norm.weight"], P[hg + ".norm.bias"], ln)
            qkv = self._buf("vit.hgt_qkv", (N, h, w, 5 * C))
            for s, e, t in runs:
                ops.linear_fwd(ln.narrow_n(s, e - s), W["%s.fold.%d" % (f, t)], Act(qkv[s:e]),
                               bias=W["%s.fold.%d.bias" % (f, t)])
            att = self._act("vit.att", X.shape)
            for b in range(B):
                s, e = starts[b], starts[b] + record_len[b]
                ops.hgt_attention_fwd(qkv[s:e], types_dev[s:e], kmask[s:e], ca["heads"], ca["dim_head"], att.narrow_n(s, e - s))
            for s, e, t in runs:
                ops.linear_fwd(att.narrow_n(s, e - s), W["%s.a_linears.%d.weight" % (f, t)], Act(X[s:e]),
                               bias=P["%s.a_linears.%d.bias" % (f, t)], accumulate=True)
            # pyramid window attention + split attention
            ops.layernorm_fwd(X, P[pwp + ".norm.weight"], P[pwp + ".norm.bias"], ln)
            wins = []
            for lv, (hh, dh, ws) in enumerate(zip(pw["heads"], pw["dim_head"], pw["window_size"])):
                bp_l = "%s.fn.pwmsa.%d" % (pwp, lv)
                wqkv = self._buf("vit.pw_qkv", (N, h, w, 3 * C))
                ops.linear_fwd(ln, W[bp_l + ".to_qkv.weight"], Act(wqkv))
                # relative offsets are key - query in the reference (mswin.py:15-20), query - key in the kernel: flip
                table = P[bp_l + ".pos_embedding"].flip(0, 1).reshape(-1, 1).expand(-1, hh).contiguous()
                ops.window_attention_fwd(wqkv, table, None, N, 1, hh, dh, ws, False, att)
                win = self._buf("vit.win%d" % lv, (N, h, w, C))
                ops.linear_fwd(att, W[bp_l + ".to_out.0.weight"], Act(win), bias=P[bp_l + ".to_out.0.bias"])
                wins.append(win)
            sp = pwp + ".fn.split_attn"
            ops.split_attn_fuse(wins[0], wins[1], wins[2], P[sp + ".fc1.weight"], P[sp + ".bn1.weight"], P[sp + ".bn1.bias"],
                                P[sp + ".fc2.weight"], self._buf("vit.sa_sums", (N, C)), self._buf("vit.sa_w", (N, 3, C)), X)
            # feed forward
            ops.layernorm_fwd(X, P[lp + ".1.norm.weight"], P[lp + ".1.norm.bias"], ln)
            hid = self._act("vit.hid", (N, h, w, enc["feed_forward"]["mlp_dim"]))
            ops.linear_fwd(ln, W[lp + ".1.fn.net.0.weight"], hid, bias=P[lp + ".1.fn.net.0.bias"], act=2)
            ops.linear_fwd(hid, W[lp + ".1.fn.net.3.weight"], Act(X), bias=P[lp + ".1.fn.net.3.bias"], accumulate=True)
        # ---- ego maps (V2XTransformer returns output[:, 0])
        fused = self._act("vit.fused", (B, h, w, C))
        ones = self._buf("vit.ones", (B,), torch.int32)
        ones.fill_(1)
        ops.regroup(X, layout["scene_start"], ones, B, 1, fused)
        return fused

    # ------------------------------------------------------------------ training step
    def forward_train(self, P, lidar, layout, prior, scm, drops=None, pre_warp=None):
        """Train-mode forward (batch-statistic BatchNorm in the encoder) keeping what the backward needs: per sublayer the
        residual input, the LayerNorm output, the projected q|k|v tensors, the attention outputs, the three window
        branches with the split-attention statistics, the FFN pre-activation and hidden activation.
        drops: None (nn.Dropout disabled) or (cav, pwindow, ffn) ops.Dropout states (each None when its rate is 0) for
        HGTCavAttention.drop_out (hmsa.py:155), BaseWindowAttention.to_out's Dropout (mswin.py:47) and FeedForward's two
        Dropouts (base_transformer.py:22,24); masks are regenerated from (seed, site) in the backward.
        pre_warp: [N, 2, 3] normalised affine maps applied to the shrunk features before the transformer (the legacy
        point_pillar_v2xvit resamples every agent's map into the ego frame first, point_pillar_v2xvit.py:140-166) or None."""
        dropc, dropw, dropf = drops if drops is not None else (None, None, None)
        self._begin_step()
        rec = []
        W = self._pack_weights(P)
        y1, y2, cat = self._encode_train(P, W, lidar, layout, rec)
        canvas_nz = self._buf("comm_rate", (1,), torch.int64)
        canvas_nz.copy_(self._canvas_nz)
        self.last_aux = {"comm_rate": canvas_nz}
        enc = self.enc
        ca, pw = enc["cav_att_config"], enc["pwindow_att_config"]
        record_len = layout["record_len"]
        B, (N, h, w, C) = len(record_len), y2.shape
        dev = self.device
        prior = prior.detach().cpu().float()
        starts = [sum(record_len[:b]) for b in range(B)]
        valid = [(b, l) for b in range(B) for l in range(record_len[b])]
        types = [int(prior[b, l, 2]) for b, l in valid]
        assert all(t in (0, 1) for t in types), "prior_encoding[..., 2] (agent type) must be 0 or 1 (hmsa.py:8)"
        types_dev = torch.tensor(types, dtype=torch.int32, device=dev)
        X = self._buf("vit.x", (N, h, w, C))
        if pre_warp is not None:
            ops.warp_affine_fwd(y2, pre_warp, Act(X), align_corners=False)
        else:
            X.copy_(y2)
        rte_idx = None
        if ca["use_RTE"]:
            rte_idx = torch.tensor([int(prior[b, l, 1]) * ca["RTE_ratio"] for b, l in valid], dtype=torch.int32, device=dev)
            rp = "fusion_net.encoder.rte.emb"
            ops.rte_add(X, P[rp + ".emb.weight"], rte_idx, P[rp + ".lin.weight"], P[rp + ".lin.bias"], self._buf("vit.rte", (N, C)))
        dr, ds = enc["sttf"]["voxel_size"][0], enc["sttf"]["downsample_rate"]
        theta_all = warp.sttf_theta(scm, dr, ds, h, w)
        theta = torch.stack([theta_all[b, l] for b, l in valid]).to(dev)
        Xw = self._buf("vit.xw", (N, h, w, C))
        ops.warp_affine_fwd(X, theta, Act(Xw), align_corners=True)
        for s in starts:
            Xw[s].copy_(X[s])
        X = Xw
        kmask = self._buf("vit.kmask", (N, h, w))
        if enc["use_roi_mask"]:
            ops.roi_mask(theta, None, N, h, w, kmask, align_corners=True)
        else:
            kmask.fill_(1.0)
        runs = self._runs(types)
        layers = []
        for d in range(enc["depth"]):
            lp, bp, hg, pwp = self._names(d)
            f = hg + ".fn"
            tag = "sv%d." % d
            sv = {}
            # HGT multi-agent attention
            ln = sv["ln_h"] = self._act(tag + "ln_h", X.shape)
            ops.layernorm_fwd(X, P[hg + ".norm.weight"], P[hg + ".norm.bias"], ln)
            qkv = sv["hqkv"] = self._buf(tag + "hqkv", (N, h, w, 5 * C))
            for s, e, t in runs:
                ops.linear_fwd(ln.narrow_n(s, e - s), W["%s.fold.%d" % (f, t)], Act(qkv[s:e]),
                               bias=W["%s.fold.%d.bias" % (f, t)])
            att = sv["hatt"] = self._act(tag + "hatt", X.shape)
            for b in range(B):
                s, e = starts[b], starts[b] + record_len[b]
                ops.hgt_attention_fwd(qkv[s:e], types_dev[s:e], kmask[s:e], ca["heads"], ca["dim_head"], att.narrow_n(s, e - s))
            # x_new = x + dropout(a_linear(att)) in the GEMM epilogue, into a fresh buffer: the sublayer input is kept (LayerNorm
            # backward) without a copy; the per-type slices index the site's mask by their offset in the whole tensor
            if dropc is not None:
                sv["site_h"] = dropc.site()
            Xn = self._buf(tag + "x_h", X.shape)
            for s, e, t in runs:
                ops.linear_dropout_residual_fwd(att.narrow_n(s, e - s), W["%s.a_linears.%d.weight" % (f, t)], Xn[s:e],
                                                bias=P["%s.a_linears.%d.bias" % (f, t)], residual=X[s:e], drop=dropc,
                                                site=sv.get("site_h", 0), elem_offset=s * h * w * C)
            sv["xin_h"] = X
            X = Xn
            # pyramid window attention + split attention
            sv["xin_p"] = self._buf(tag + "xin_p", X.shape)
            sv["xin_p"].copy_(X)
            ln = sv["ln_p"] = self._act(tag + "ln_p", X.shape)
            ops.layernorm_fwd(X, P[pwp + ".norm.weight"], P[pwp + ".norm.bias"], ln)
            sv["wqkv"], sv["watt"], sv["win"], sv["table"] = [], [], [], []
            for lv, (hh, dh, ws) in enumerate(zip(pw["heads"], pw["dim_head"], pw["window_size"])):
                bp_l = "%s.fn.pwmsa.%d" % (pwp, lv)
                wqkv = self._buf(tag + "wqkv%d" % lv, (N, h, w, 3 * C))
                ops.linear_fwd(ln, W[bp_l + ".to_qkv.weight"], Act(wqkv))
                table = P[bp_l + ".pos_embedding"].flip(0, 1).reshape(-1, 1).expand(-1, hh).contiguous()
                watt = self._act(tag + "watt%d" % lv, X.shape)
                ops.window_attention_fwd(wqkv, table, None, N, 1, hh, dh, ws, False, watt)
                win = self._buf(tag + "win%d" % lv, (N, h, w, C))
                if dropw is not None:
                    sv.setdefault("site_w", []).append(dropw.site())
                ops.linear_dropout_residual_fwd(watt, W[bp_l + ".to_out.0.weight"], win, bias=P[bp_l + ".to_out.0.bias"],
                                                drop=dropw, site=sv["site_w"][-1] if dropw is not None else 0)
                sv["wqkv"].append(wqkv)
                sv["watt"].append(watt)
                sv["win"].append(win)
                sv["table"].append(table)
            sp = pwp + ".fn.split_attn"
            sv["sa_sums"], sv["sa_w"] = self._buf(tag + "sa_sums", (N, C)), self._buf(tag + "sa_w", (N, 3, C))
            ops.split_attn_fuse(sv["win"][0], sv["win"][1], sv["win"][2], P[sp + ".fc1.weight"], P[sp + ".bn1.weight"],
                                P[sp + ".bn1.bias"], P[sp + ".fc2.weight"], sv["sa_sums"], sv["sa_w"], X)
            # feed forward (pre-activation kept in fp32: GELU' needs it)
            ln = sv["ln_f"] = self._act(tag + "ln_f", X.shape)
            ops.layernorm_fwd(X, P[lp + ".1.norm.weight"], P[lp + ".1.norm.bias"], ln)
            pre = sv["hpre"] = self._buf(tag + "hpre", (N, h, w, enc["feed_forward"]["mlp_dim"]))
            ops.linear_fwd(ln, W[lp + ".1.fn.net.0.weight"], Act(pre), bias=P[lp + ".1.fn.net.0.bias"])
            hid = sv["hid"] = self._act(tag + "hid", pre.shape)
            if dropf is None:
                ops.gelu_fwd(pre, hid)
            else:   # x_new = x + dropout(W2 dropout(gelu(pre)) + b2)
                sv["site_f1"] = dropf.site()
                ops.gelu_dropout_fwd(pre, dropf, sv["site_f1"], hid)
                sv["site_f2"] = dropf.site()
            Xn = self._buf(tag + "x_f", X.shape)
            ops.linear_dropout_residual_fwd(hid, W[lp + ".1.fn.net.3.weight"], Xn, bias=P[lp + ".1.fn.net.3.bias"], residual=X,
                                            drop=dropf, site=sv.get("site_f2", 0))
            sv["xin_f"] = X
            X = Xn
            layers.append(sv)
        fused = self._act("vit.fused", (B, h, w, C))
        ones = self._buf("vit.ones", (B,), torch.int32)
        ones.fill_(1)
        ops.regroup(X, layout["scene_start"], ones, B, 1, fused)
        heads = self._buf("heads.out", (B, h, w, HEAD_PAD))
        ops.linear_fwd(fused, W["heads"], Act(heads), bias=W["heads.bias"])
        self.saved = dict(rec=rec, W=W, layers=layers, fused=fused, y1=y1, y2=y2, cat=cat, layout=layout, B=B, runs=runs,
                          starts=starts, types_dev=types_dev, kmask=kmask, theta=theta, rte_idx=rte_idx,
                          drops=(dropc, dropw, dropf), pre_warp=pre_warp)
        return heads

    def backward_train(self, P, dheads, grads):
        """dheads: [B,h,w,HEAD_PAD] gradient w.r.t. the head logits; grads: name -> fp32 tensor (written; the unused
        `prior_feed` stays zero, as autograd leaves it in the reference)."""
        S = self.saved
        W, layout, B, runs, starts = S["W"], S["layout"], S["B"], S["runs"], S["starts"]
        dropc, dropw, dropf = S["drops"]
        enc = self.enc
        ca, pw = enc["cav_att_config"], enc["pwindow_att_config"]
        record_len = layout["record_len"]
        nc, nr = self.A * self.K, 7 * self.A
        C = self.dim
        N, h, w, _ = S["y2"].shape
        unpack = []
        dwps = {}

        def zero_f32(n):
            return self._zeroed("z", n, torch.float32)

        def dwp_for(wname, co, ci):
            if wname not in dwps:
                dwps[wname] = zero_f32(co * ci).view(1, co, ci)
                unpack.append(ops.conv_unpack_job(dwps[wname], grads[wname].view(co, ci, 1, 1)))
            return dwps[wname]

        def lin_wgrad(x_act, dy_act, wname):
            """grad of nn.Linear weight [out, in] = 1x1 conv wgrad; repeated calls for one weight accumulate"""
            ops.conv_wgrad(x_act, dy_act, 1, 1, dwp_for(wname, dy_act.shape[3], x_act.shape[3]))

        def col_sums(t, Cc, outs):
            sums = self._zeroed("bias.sums", 2 * Cc, torch.float64)
            ops.channel_stats(t, sums)
            for out, c0 in outs:
                unpack.append(ops.sums_unpack_job(sums, out, c0))

        def split_of(t, name):
            a = self._act(name, t.shape)
            n_, h_, w_, c_ = t.shape
            if c_ > 1024:  # elementwise: present wide rows to the kernel as C-wide ones
                k = c_ // C
                ops.affine_act(t.view(n_, h_, w_ * k, C), None, None, False,
                               Act(a.hi.view(n_, h_, w_ * k, C), None if a.b16 is None else a.b16.view(2, n_, h_, w_ * k, C)),
                               write_hi=False)
            else:
                ops.affine_act(t, None, None, False, a, write_hi=False)  # GEMM operand only
            return a

        def ln_bwd(xin, d_ln, pre_norm, dX):
            acc = self._zeroed(pre_norm + ".lnacc", 2 * C, torch.float64)
            ops.layernorm_bwd(xin, d_ln, P[pre_norm + ".weight"], dX, acc[:C], acc[C:])
            unpack.append(ops.sums_unpack_job(acc, grads[pre_norm + ".weight"], 0))
            unpack.append(ops.sums_unpack_job(acc, grads[pre_norm + ".bias"], C))

        # ---- heads; only the ego rows of the encoder output are used (V2XTransformer returns output[:, 0])
        dh = split_of(dheads, "bwd.dheads")
        dwp = zero_f32(HEAD_PAD * self.c_shrink).view(1, HEAD_PAD, self.c_shrink)
        ops.conv_wgrad(S["fused"], dh, 1, 1, dwp)
        for name, row0 in self._head_rows():
            unpack.append(ops.conv_unpack_job(dwp, grads[name + ".weight"], row0))
        col_sums(dheads, HEAD_PAD, [(grads[name + ".bias"], row0) for name, row0 in self._head_rows()])
        d_fused = self._buf("bwd.d_fused", S["fused"].shape)
        ops.conv_dgrad(dh, W["heads"], 1, 1, d_fused)
        dX = self._buf("bwd.dX", (N, h, w, C))
        dX.zero_()
        for b, s in enumerate(starts):
            dX[s].copy_(d_fused[b])
        d_ln = self._buf("bwd.d_ln", (N, h, w, C))
        for d in range(enc["depth"] - 1, -1, -1):
            lp, bp, hg, pwp = self._names(d)
            f = hg + ".fn"
            sv = S["layers"][d]
            # ---- feed forward
            pre = lp + ".1"
            if dropf is None:
                dXs, g_out = split_of(dX, "bwd.dXs"), dX
            else:
                dXs = self._act("bwd.dXs", dX.shape)
                ops.dropout_apply(dX, dropf, sv["site_f2"], dXs)
                g_out = dXs.hi
            lin_wgrad(sv["hid"], dXs, pre + ".fn.net.3.weight")
            col_sums(g_out, C, [(grads[pre + ".fn.net.3.bias"], 0)])
            d_hid = self._buf("bwd.d_hid", sv["hid"].shape)
            ops.conv_dgrad(dXs, W[pre + ".fn.net.3.weight"], 1, 1, d_hid)
            d_pre = self._act("bwd.d_pre", sv["hid"].shape)
            if dropf is None:
                ops.gelu_bwd(d_hid, sv["hpre"], d_pre)
            else:
                ops.gelu_dropout_bwd(d_hid, sv["hpre"], dropf, sv["site_f1"], d_pre)
            lin_wgrad(sv["ln_f"], d_pre, pre + ".fn.net.0.weight")
            col_sums(d_pre.hi, d_pre.shape[3], [(grads[pre + ".fn.net.0.bias"], 0)])
            ops.conv_dgrad(d_pre, W[pre + ".fn.net.0.weight"], 1, 1, d_ln)
            ln_bwd(sv["xin_f"], d_ln, pre + ".norm", dX)
            # ---- pyramid window attention + split attention
            sp = pwp + ".fn.split_attn"
            dwin = [self._act("bwd.dwin%d" % lv, (N, h, w, C)) for lv in range(3)]
            for n in (".fc1.weight", ".bn1.weight", ".bn1.bias", ".fc2.weight"):
                grads[sp + n].zero_()
            ops.split_attn_bwd(dX, sv["win"][0], sv["win"][1], sv["win"][2], P[sp + ".fc1.weight"], P[sp + ".bn1.weight"],
                               P[sp + ".bn1.bias"], P[sp + ".fc2.weight"], sv["sa_sums"], sv["sa_w"],
                               self._buf("bwd.sa_dw", (N, 3, C)), self._buf("bwd.sa_dgap", (N, C)), dwin[0], dwin[1], dwin[2],
                               grads[sp + ".fc1.weight"], grads[sp + ".bn1.weight"], grads[sp + ".bn1.bias"],
                               grads[sp + ".fc2.weight"])
            for lv, (hh, dhd, ws) in enumerate(zip(pw["heads"], pw["dim_head"], pw["window_size"])):
                bp_l = "%s.fn.pwmsa.%d" % (pwp, lv)
                if dropw is not None:   # d(to_out output) = d(branch) through the branch's dropout mask
                    ops.dropout_apply(dwin[lv].hi, dropw, sv["site_w"][lv], dwin[lv])
                lin_wgrad(sv["watt"][lv], dwin[lv], bp_l + ".to_out.0.weight")
                col_sums(dwin[lv].hi, C, [(grads[bp_l + ".to_out.0.bias"], 0)])
                d_att = self._buf("bwd.d_att", (N, h, w, C))
                ops.conv_dgrad(dwin[lv], W[bp_l + ".to_out.0.weight"], 1, 1, d_att)
                dqs = self._act("bwd.dwqs", (N, h, w, 3 * C))  # written as a GEMM operand: no fp32 plane, no conversion pass
                dbias = self._buf("bwd.dbias%d" % lv, sv["table"][lv].shape)
                dbias.zero_()
                ops.window_attention_bwd(sv["wqkv"][lv], d_att, sv["table"][lv], None, N, 1, hh, dhd, ws, False, dqs, dbias,
                                         write_hi=False)
                # the table is the flipped pos_embedding broadcast over heads (mswin.py:15-20)
                grads[bp_l + ".pos_embedding"].copy_(dbias.sum(1).view(2 * ws - 1, 2 * ws - 1).flip(0, 1))
                lin_wgrad(sv["ln_p"], dqs, bp_l + ".to_qkv.weight")
                ops.conv_dgrad(dqs, W[bp_l + ".to_qkv.weight"], 1, 1, d_ln, accumulate=lv > 0)
            ln_bwd(sv["xin_p"], d_ln, pwp + ".norm", dX)
            # ---- HGT multi-agent attention
            if dropc is None:
                dXs, g_out = split_of(dX, "bwd.dXs"), dX
            else:
                dXs = self._act("bwd.dXs", dX.shape)
                ops.dropout_apply(dX, dropc, sv["site_h"], dXs)
                g_out = dXs.hi
            d_att = self._buf("bwd.d_att", (N, h, w, C))
            asums = self._zeroed("hgt.asums", 2 * 2 * C, torch.float64).view(2, 2 * C)
            for t in range(2):  # an agent type absent from the batch gets a zero gradient
                dwp_for("%s.a_linears.%d.weight" % (f, t), C, C)
            for s, e, t in runs:
                lin_wgrad(sv["hatt"].narrow_n(s, e - s), dXs.narrow_n(s, e - s), "%s.a_linears.%d.weight" % (f, t))
                ops.channel_stats(g_out[s:e], asums[t])
                ops.conv_dgrad(dXs.narrow_n(s, e - s), W["%s.a_linears.%d.weight" % (f, t)], 1, 1, d_att[s:e])
            for t in range(2):
                unpack.append(ops.sums_unpack_job(asums[t], grads["%s.a_linears.%d.bias" % (f, t)], 0))
            dhq = self._buf("bwd.dhqkv", (N, h, w, 5 * C))
            for b in range(B):
                s, e = starts[b], starts[b] + record_len[b]
                ops.hgt_attention_bwd(sv["hqkv"][s:e], S["types_dev"][s:e], S["kmask"][s:e], d_att[s:e], ca["heads"],
                                      ca["dim_head"], dhq[s:e])
            dhs = split_of(dhq, "bwd.dhqs")
            dwf = zero_f32(2 * 5 * C * C).view(2, 5 * C, C)
            fsums = self._zeroed("hgt.fsums", 2 * 5 * 2 * C, torch.float64).view(2, 5, 2 * C)
            for s, e, t in runs:
                ops.conv_wgrad(sv["ln_h"].narrow_n(s, e - s), dhs.narrow_n(s, e - s), 1, 1, dwf[t].view(1, 5 * C, C))
                for j in range(5):
                    ops.channel_stats(dhq[s:e, :, :, j * C:(j + 1) * C], fsums[t, j])
                ops.conv_dgrad(dhs.narrow_n(s, e - s), W["%s.fold.%d" % (f, t)], 1, 1, d_ln[s:e])
            dbf = self._buf("bwd.dbf", (2, 5 * C))
            for t in range(2):
                for j in range(5):
                    ops.sums_to_float(fsums[t, j], C, dbf[t, j * C:(j + 1) * C])
            pair = lambda n, s_: (P["%s.%s.0.%s" % (f, n, s_)], P["%s.%s.1.%s" % (f, n, s_)])
            gpair = lambda n, s_: (grads["%s.%s.0.%s" % (f, n, s_)], grads["%s.%s.1.%s" % (f, n, s_)])
            ops.hgt_fold_bwd(dwf, dbf, pair("k_linears", "weight"), pair("k_linears", "bias"), pair("v_linears", "weight"),
                             pair("v_linears", "bias"), P[f + ".relation_att"], P[f + ".relation_msg"], ca["heads"],
                             gpair("q_linears", "weight"), gpair("q_linears", "bias"), gpair("k_linears", "weight"),
                             gpair("k_linears", "bias"), gpair("v_linears", "weight"), gpair("v_linears", "bias"),
                             grads[f + ".relation_att"], grads[f + ".relation_msg"])
            ln_bwd(sv["xin_h"], d_ln, hg + ".norm", dX)
        # ---- STTF^T: non-ego gradients go back through the bilinear resampling, the ego map was copied
        d_pre_warp = self._buf("bwd.d_prewarp", (N, h, w, C))
        d_pre_warp.zero_()
        ops.warp_affine_bwd(dX, S["theta"], d_pre_warp, align_corners=True)
        for s in starts:
            d_pre_warp[s].copy_(dX[s])
        # ---- RTE^T: the per-agent vector's gradient is the column sum of its map's gradient
        if S["rte_idx"] is not None:
            rp = "fusion_net.encoder.rte.emb"
            rsums = self._zeroed("rte.sums", N * 2 * C, torch.float64).view(N, 2 * C)
            for a in range(N):
                ops.channel_stats(d_pre_warp[a:a + 1], rsums[a])
            for n in (".lin.weight", ".lin.bias", ".emb.weight"):
                grads[rp + n].zero_()
            ops.rte_bwd(rsums, P[rp + ".emb.weight"], S["rte_idx"], P[rp + ".lin.weight"], grads[rp + ".lin.weight"],
                        grads[rp + ".lin.bias"], grads[rp + ".emb.weight"])
        for n in ("fusion_net.encoder.prior_feed.weight", "fusion_net.encoder.prior_feed.bias"):
            if n in grads:
                grads[n].zero_()
        if S.get("pre_warp") is not None:   # the legacy model's ego-frame resampling, transposed
            d_y2 = self._buf("bwd.d_y2", (N, h, w, C))
            d_y2.zero_()
            ops.warp_affine_bwd(d_pre_warp, S["pre_warp"], d_y2, align_corners=False)
            d_pre_warp = d_y2
        self._encoder_backward(P, S, d_pre_warp, grads, unpack)
        for lo in range(0, len(unpack), 128):
            ops.unpack_wgrads_batched(self._job_table("unpack%d" % lo, unpack[lo:lo + 128]))
        return grads

    def forward(self, P, lidar, layout, training, prior=None, scm=None):
        if training:
            raise NotImplementedError("Airv2xV2XVit: train-mode forward under torch.no_grad() is not implemented (model.eval() "
                                      "for inference; model(batch) with grad enabled or train_step() for training)")
        self._begin_step()
        W = self._pack_weights(P)
        canvas_nz = self._buf("comm_rate", (1,), torch.int64)
        feat = self.encode(P, W, lidar, layout)
        canvas_nz.copy_(self._canvas_nz)
        fused = self.fusion(P, W, feat, layout, prior, scm)
        heads = self._buf("heads.out", (fused.shape[0], feat.shape[1], feat.shape[2], HEAD_PAD))
        ops.linear_fwd(fused, W["heads"], Act(heads), bias=W["heads.bias"])
        return heads, {"comm_rate": canvas_nz}
