"""TEST INFRASTRUCTURE ONLY: CPU restatement (torch fp32) of the reference's V2X-ViT path, BASELINE config 3
(`airv2x_intermediate_v2xvit.yaml`). Only tests/, __graft_entry__.smoke() and scripts/ may import it; the product
path never does. Pinned against the REAL reference by scripts/make_golden_v2xvit.py (eval mode).

Follows, function by function:
  opencood/models/airv2x_v2xvit.py:108-167                               model forward
  opencood/models/common_modules/fuse_utils.py:13-63                     regroup
  opencood/models/v2xvit_modules/v2xvit_basic.py:17-38                   STTF (warp non-ego maps, align_corners=True)
  opencood/models/v2xvit_modules/v2xvit_basic.py:41-80                   RTE (sinusoid table + Linear, added per agent)
  opencood/models/v2xvit_modules/v2xvit_basic.py:83-213                  V2XFusionBlock / V2XTEncoder / V2XTransformer
  opencood/models/v2xvit_modules/hmsa.py:6-158                           HGTCavAttention (typed q/k/v/a, relation tensors)
  opencood/models/v2xvit_modules/mswin.py:15-145                         Base / PyramidWindowAttention
  opencood/models/v2xvit_modules/split_attn.py:6-63                      RadixSoftmax / SplitAttn
  opencood/models/common_modules/torch_transformation_utils.py:15-143, :146-381   ROI mask, pixel homographies, warp
"""
import torch
import torch.nn.functional as F

from . import cobevt_oracle as CO, w2c_oracle as O


# ----------------------------------------------------------------------------------------------- geometry helpers
def discretized_matrix(m, discrete_ratio, downsample_rate):
    """torch_transformation_utils.py:116-143: rows 0,1 / cols 0,1,3 of the 4x4 pose, translation in feature pixels"""
    m = m[:, :, [0, 1], :][:, :, :, [0, 1, 3]].clone()
    m[:, :, :, -1] = m[:, :, :, -1] / (discrete_ratio * downsample_rate)
    return m.float()


def transformation_matrix(M, hw):
    """:265-308: rotation about the image centre + pixel translation (N,2,3)"""
    H, W = hw
    N = M.shape[0]
    eye = torch.eye(3, dtype=M.dtype).repeat(N, 1, 1)
    shift, shift_inv, rot = eye.clone(), eye.clone(), eye.clone()
    center = torch.tensor([W / 2, H / 2], dtype=M.dtype)
    shift[:, :2, 2] = center
    shift_inv[:, :2, 2] = -center
    rot[:, :2, :2] = M[:, :2, :2]
    T = (shift @ rot @ shift_inv)[:, :2, :].clone()
    T[..., 2] += M[..., 2]
    return T


def _norm_pixel(h, w, dtype):
    t = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], dtype=dtype)
    t[0, 0] = t[0, 0] * 2.0 / (1e-14 if w == 1 else w - 1.0)
    t[1, 1] = t[1, 1] * 2.0 / (1e-14 if h == 1 else h - 1.0)
    return t[None]


def warp_theta(M, hw):
    """the (N,2,3) matrices warp_affine hands to F.affine_grid(align_corners=True): inverse of the normalised
    homography (:337-381, :203-262)"""
    H, W = hw
    M3 = F.pad(M, [0, 0, 0, 1], "constant", 0.0).clone()
    M3[..., -1, -1] += 1.0
    n = _norm_pixel(H, W, M.dtype)
    dst_norm_trans_src_norm = n @ (M3 @ torch.inverse(n))
    return torch.inverse(dst_norm_trans_src_norm)[:, :2, :]


def warp_affine(src, M, hw, mode="bilinear"):
    theta = warp_theta(M, hw)
    grid = F.affine_grid(theta, [src.shape[0], src.shape[1], hw[0], hw[1]], align_corners=True)
    return F.grid_sample(src, grid, align_corners=True, mode=mode, padding_mode="zeros")


def roi_and_cav_mask(shape, cav_mask, scm, discrete_ratio, downsample_rate):
    """:15-53 -> (B,H,W,1,L)"""
    B, L, H, W, _ = shape
    T = transformation_matrix(discretized_matrix(scm, discrete_ratio, downsample_rate).reshape(-1, 2, 3), (H, W))
    roi = warp_affine(torch.ones(B * L, 1, H, W, dtype=T.dtype), T, (H, W), mode="nearest").reshape(B, L, 1, H, W)
    com = roi * cav_mask[:, :, None, None, None]
    return com.permute(0, 3, 4, 2, 1)


# ----------------------------------------------------------------------------------------------- encoder pieces
def _ln(sd, pre, x):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + ".weight"], sd[pre + ".bias"], 1e-5)


def rte(sd, pre, x, dts, ratio):
    """v2xvit_basic.py:41-80: x[b, l] += Linear(emb[dt * ratio])"""
    e = F.embedding(dts * ratio, sd[pre + ".emb.emb.weight"])
    return x + F.linear(e, sd[pre + ".emb.lin.weight"], sd[pre + ".emb.lin.bias"])[:, :, None, None, :]


def sttf(x, scm, discrete_ratio, downsample_rate):
    """v2xvit_basic.py:17-38: x (B,L,H,W,C); non-ego maps are warped by the spatial correction"""
    B, L, H, W, C = x.shape
    xc = x.permute(0, 1, 4, 2, 3)
    T = transformation_matrix(discretized_matrix(scm, discrete_ratio, downsample_rate)[:, 1:].reshape(-1, 2, 3), (H, W))
    cav = warp_affine(xc[:, 1:].reshape(-1, C, H, W), T, (H, W)).reshape(B, -1, C, H, W)
    return torch.cat([xc[:, :1], cav], 1).permute(0, 1, 3, 4, 2)


def hgt_attention(sd, pre, x, mask, prior, heads, dim_head, num_types=2, drop=None):
    """hmsa.py:117-158. x (B,L,H,W,C) normalised; mask (B,H,W,1,L); prior (B,L,H,W,3); drop: the nn.Dropout of :155"""
    B, L, H, W, C = x.shape
    types = prior[:, :, 0, 0, 2].to(torch.int)

    def typed(name, t):
        out = torch.empty(B, L, H, W, sd["%s.%s.0.weight" % (pre, name)].shape[0], dtype=t.dtype)
        for b in range(B):
            for i in range(L):
                ty = int(types[b, i])
                out[b, i] = F.linear(t[b, i], sd["%s.%s.%d.weight" % (pre, name, ty)], sd["%s.%s.%d.bias" % (pre, name, ty)])
        return out

    def split(t):  # (B,L,H,W,M*c) -> (B,M,H,W,L,c)
        return t.reshape(B, L, H, W, heads, dim_head).permute(0, 4, 2, 3, 1, 5)

    q, k, v = split(typed("q_linears", x)), split(typed("k_linears", x)), split(typed("v_linears", x))
    e = types[:, :, None] * num_types + types[:, None, :]                       # (B, L_i, L_j)
    w_att = sd[pre + ".relation_att"][e.long()].permute(0, 3, 1, 2, 4, 5)        # (B,M,L,L,c,c)
    w_msg = sd[pre + ".relation_msg"][e.long()].permute(0, 3, 1, 2, 4, 5)
    att = torch.einsum("bmhwip,bmijpq,bmhwjq->bmhwij", q, w_att, k) * dim_head ** -0.5
    att = att.masked_fill(mask.unsqueeze(1) == 0, -float("inf")).softmax(-1)
    v_msg = torch.einsum("bmijpc,bmhwjp->bmhwijc", w_msg, v)
    out = torch.einsum("bmhwij,bmhwijc->bmhwic", att, v_msg)
    out = out.permute(0, 4, 2, 3, 1, 5).reshape(B, L, H, W, heads * dim_head)
    out = typed("a_linears", out)
    return drop(out) if drop is not None else out


def base_window_attention(sd, pre, x, heads, dim_head, ws, drop=None):
    """mswin.py:23-108 (relative_pos_embedding = True); drop: to_out's nn.Dropout (:47)"""
    B, L, H, W, C = x.shape
    nh, nw = H // ws, W // ws
    q, k, v = F.linear(x, sd[pre + ".to_qkv.weight"]).chunk(3, -1)

    def win(t):  # b l (nh wh) (nw ww) (m c) -> b l m (nh nw) (wh ww) c
        return t.reshape(B, L, nh, ws, nw, ws, heads, dim_head).permute(0, 1, 6, 2, 4, 3, 5, 7).reshape(
            B, L, heads, nh * nw, ws * ws, dim_head)

    q, k, v = win(q), win(k), win(v)
    dots = q @ k.transpose(-1, -2) * dim_head ** -0.5
    idx = torch.tensor([[a, b] for a in range(ws) for b in range(ws)])
    rel = idx[None, :, :] - idx[:, None, :] + ws - 1
    dots = dots + sd[pre + ".pos_embedding"][rel[:, :, 0], rel[:, :, 1]]
    out = dots.softmax(-1) @ v
    out = out.reshape(B, L, heads, nh, nw, ws, ws, dim_head).permute(0, 1, 3, 5, 4, 6, 2, 7).reshape(B, L, H, W, heads * dim_head)
    out = F.linear(out, sd[pre + ".to_out.0.weight"], sd[pre + ".to_out.0.bias"])
    return drop(out) if drop is not None else out


def split_attn(sd, pre, wins):
    """split_attn.py:28-63"""
    sw, mw, bw = wins
    B, L, _, _, C = sw.shape
    gap = (sw + mw + bw).mean((2, 3), keepdim=True)
    gap = F.relu(_ln(sd, pre + ".bn1", F.linear(gap, sd[pre + ".fc1.weight"])))
    a = F.linear(gap, sd[pre + ".fc2.weight"])
    a = a.view(B, L, 1, 3, -1).softmax(3).reshape(B, -1).view(B, L, 1, 1, -1)
    return sw * a[..., 0:C] + mw * a[..., C:2 * C] + bw * a[..., 2 * C:]


def pyramid_window_attention(sd, pre, x, cfg, drop=None):
    wins = [base_window_attention(sd, "%s.pwmsa.%d" % (pre, i), x, h, d, ws, drop)
            for i, (h, d, ws) in enumerate(zip(cfg["heads"], cfg["dim_head"], cfg["window_size"]))]
    if cfg["fusion_method"] == "split_attn":
        return split_attn(sd, pre + ".split_attn", wins)
    return sum(wins) / len(wins)


class MaskedDropouts:
    """test hook: the three nn.Dropout groups of the encoder (HGT output, window branches, feed forward) with GIVEN keep
    masks, consumed in call order; each mask is laid out like the tensor it drops, (B, L, H, W, C) flattened."""

    def __init__(self, ps, masks):
        """ps = (p_cav, p_window, p_ffn); masks = {"cav": [...], "win": [...], "ffn": [...]} (uint8 tensors)"""
        self.ps, self.masks, self.i = ps, masks, {"cav": 0, "win": 0, "ffn": 0}

    def fn(self, group):
        p = self.ps[("cav", "win", "ffn").index(group)]
        if p <= 0:
            return None

        def apply(t):
            m = self.masks[group][self.i[group]]
            self.i[group] += 1
            return t * m.reshape(t.shape).to(t.dtype) / (1.0 - p)
        return apply


def v2x_encoder(sd, enc, x, mask, scm, pre="fusion_net.encoder", keep=None, dropouts=None):
    """V2XTEncoder.forward v2xvit_basic.py:174-200 + V2XTransformer (:211-213). x (B,L,H,W,C+3)"""
    d_cav = dropouts.fn("cav") if dropouts is not None else None
    d_win = dropouts.fn("win") if dropouts is not None else None
    d_ffn = dropouts.fn("ffn") if dropouts is not None else None
    prior = x[..., -3:]
    x = x[..., :-3]
    ca, pw = enc["cav_att_config"], enc["pwindow_att_config"]
    if ca["use_RTE"]:
        x = rte(sd, pre + ".rte", x, prior[:, :, 0, 0, 1].to(torch.int), ca["RTE_ratio"])
    dr, ds = enc["sttf"]["voxel_size"][0], enc["sttf"]["downsample_rate"]
    x = sttf(x, scm, dr, ds)
    if keep is not None:
        keep["sttf"] = x
    com = roi_and_cav_mask(x.shape, mask, scm, dr, ds) if enc["use_roi_mask"] else mask[:, None, None, None, :]
    if keep is not None:
        keep["com_mask"] = com
    for d in range(enc["depth"]):
        lp = "%s.layers.%d" % (pre, d)
        for blk in range(enc["num_blocks"]):
            bp = "%s.0.layers.%d" % (lp, blk)
            assert ca["use_hetero"]
            x = hgt_attention(sd, bp + ".0.fn", _ln(sd, bp + ".0.norm", x), com, prior, ca["heads"], ca["dim_head"],
                              drop=d_cav) + x
            x = pyramid_window_attention(sd, bp + ".1.fn", _ln(sd, bp + ".1.norm", x), pw, d_win) + x
        x = CO.feed_forward(sd, lp + ".1.fn", _ln(sd, lp + ".1.norm", x), d_ffn) + x
        if keep is not None:
            keep["layer%d" % d] = x
    return x[:, 0]


def v2xvit_forward(sd, args, data_dict, training=False, keep=None, dropouts=None):
    """models/airv2x_v2xvit.py:108-167 (task == det; eval mode: dropout = identity)."""
    buffers = {}
    mf = args["modality_fusion"]
    sf, record_len = O.extract_features(sd, args, data_dict, training, buffers, keep)
    comm_rate = int(sf.count_nonzero().item())
    feat = O.backbone_forward(sd, mf["base_bev_backbone"], sf, training, buffers)
    if mf["shrink_header"]["use"]:
        feat = O.shrink_conv(sd, mf["shrink_header"], feat)
    L = sum(args["max_cav"].values())
    x, mask = CO.regroup(feat, record_len.tolist(), L)                              # (B,L,C,H,W)
    prior = data_dict["prior_encoding"][:, :, :, None, None].expand(-1, -1, -1, x.shape[3], x.shape[4]).to(x.dtype)
    x = torch.cat([x, prior], 2).permute(0, 1, 3, 4, 2).contiguous()
    fused = v2x_encoder(sd, args["transformer"]["encoder"], x, mask, data_dict["spatial_correction_matrix"], keep=keep,
                        dropouts=dropouts)
    fused = fused.permute(0, 3, 1, 2).contiguous()
    if keep is not None:
        keep["fused_feature"] = fused
    out = {"psm": F.conv2d(fused, sd["cls_head.weight"], sd["cls_head.bias"]),
           "rm": F.conv2d(fused, sd["reg_head.weight"], sd["reg_head.bias"])}
    if args["obj_head"]:
        out["obj"] = F.conv2d(fused, sd["obj_head.weight"], sd["obj_head.bias"])
    out["comm_rate"] = comm_rate
    return out, buffers


def pp_v2xvit_forward(sd, args, data_dict, training=False, keep=None):
    """models/point_pillar_v2xvit.py:82-185 (`point_pillar_v2xvit`; dropout = identity): legacy encoder, regroup, zero
    prior encoding, every agent's map resampled into the ego frame by pairwise_t_matrix[b, 0] (warp_affine_simple:
    F.affine_grid / F.grid_sample, align_corners False, torch_transformation_utils.py:327-334; normalisation with the
    pillar canvas' H, W and downsample_rate 1, :140-156), V2XTransformer with an identity spatial correction."""
    buffers = {}
    feat, comm_rate, (H, W) = CO._legacy_encoder(sd, args, data_dict, training, buffers)
    L = args["max_cav"]
    x, mask = CO.regroup(feat, data_dict["record_len"].tolist(), L)                     # (B,L,C,h,w)
    B, _, C, h, w = x.shape
    x = torch.cat([x, torch.zeros(B, L, 3, h, w)], 2)
    t = data_dict["pairwise_t_matrix"][:, :, :, [0, 1], :][:, :, :, :, [0, 1, 3]].clone().float()
    dr = args["voxel_size"][0]
    t[..., 0, 1] = t[..., 0, 1] * H / W
    t[..., 1, 0] = t[..., 1, 0] * W / H
    t[..., 0, 2] = t[..., 0, 2] / (1 * dr * W) * 2
    t[..., 1, 2] = t[..., 1, 2] / (1 * dr * H) * 2
    out = []
    for b in range(B):
        grid = F.affine_grid(t[b, 0], [L, C + 3, h, w], align_corners=False)
        out.append(F.grid_sample(x[b], grid, align_corners=False))
    x = torch.stack(out).permute(0, 1, 3, 4, 2)
    if keep is not None:
        keep["warped"] = x
    scm = torch.eye(4).expand(B, L, 4, 4)
    fused = v2x_encoder(sd, args["transformer"]["encoder"], x, mask, scm, keep=keep).permute(0, 3, 1, 2)
    return {"psm": F.conv2d(fused, sd["cls_head.weight"], sd["cls_head.bias"]),
            "rm": F.conv2d(fused, sd["reg_head.weight"], sd["reg_head.bias"]),
            "mask": 0, "each_mask": 0, "comm_rate": comm_rate}, buffers
