"""TEST INFRASTRUCTURE ONLY: CPU restatement (torch fp32) of the reference's CoBEVT path, BASELINE config 4
(`airv2x_intermediate_cobevt.yaml`). Only tests/, __graft_entry__.smoke() and scripts/ may import it; the product
path never does. Pinned against the REAL reference by scripts/make_golden_cobevt.py (differences 0.0 in eval mode).

Follows, function by function:
  opencood/models/airv2x_cobevt.py:112-156                      model forward
  opencood/models/cobevt_modules/fuse_utils.py:13-63            regroup (zero-pad every scene to L agents + mask)
  opencood/models/cobevt_modules/swap_fusion_modules.py:14-127  Attention (3-D window attention, rel-pos bias, key mask)
  opencood/models/cobevt_modules/swap_fusion_modules.py:130-195 SwapFusionBlockMask (window attn, FFN, grid attn, FFN)
  opencood/models/cobevt_modules/swap_fusion_modules.py:233-280 SwapFusionEncoder (+ mean over agents, LN, Linear)
  opencood/models/cobevt_modules/base_transformer.py:6-28       PreNormResidual, FeedForward (GELU)
The encoder half (PillarVFE, scatter, backbone, shrink) is shared with oracle/w2c_oracle.py.
"""
import torch
import torch.nn.functional as F

from . import w2c_oracle as O


def regroup(x, record_len, L):
    """fuse_utils.py:13-63: [sum n, C, H, W] -> [B, L, C, H, W] zero padded, mask [B, L]"""
    outs, masks, pos = [], [], 0
    for n in [int(v) for v in record_len]:
        f = x[pos:pos + n]
        pos += n
        pad = torch.zeros(L - n, *f.shape[1:], dtype=f.dtype)
        outs.append(torch.cat([f, pad], 0))
        masks.append([1] * n + [0] * (L - n))
    return torch.stack(outs), torch.tensor(masks)


def relative_position_index(L, w):
    """swap_fusion_modules.py:53-76"""
    cd, ch, cw = torch.meshgrid(torch.arange(L), torch.arange(w), torch.arange(w), indexing="ij")
    c = torch.stack([cd, ch, cw]).flatten(1)              # 3, n
    rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).contiguous()
    rel[..., 0] += L - 1
    rel[..., 1] += w - 1
    rel[..., 2] += w - 1
    rel[..., 0] *= (2 * w - 1) * (2 * w - 1)
    rel[..., 1] *= 2 * w - 1
    return rel.sum(-1)


def window_attention(sd, pre, x, mask, dim_head):
    """x: [b, L, X, Y, w1, w2, d] (already LayerNorm-ed); mask: [b, X, Y, w1, w2, 1, L]; :78-127"""
    b, L, X, Y, w1, w2, d = x.shape
    h = d // dim_head
    t = x.permute(0, 2, 3, 1, 4, 5, 6).reshape(b * X * Y, L * w1 * w2, d)
    q, k, v = F.linear(t, sd[pre + ".to_qkv.weight"]).chunk(3, -1)

    def heads(z):
        return z.reshape(z.shape[0], z.shape[1], h, dim_head).permute(0, 2, 1, 3)

    q, k, v = heads(q) * dim_head ** -0.5, heads(k), heads(v)
    sim = q @ k.transpose(-1, -2)
    idx = sd.get(pre + ".relative_position_index")
    if idx is None:
        idx = relative_position_index(L, w1)
    bias = F.embedding(idx, sd[pre + ".relative_position_bias_table.weight"])    # n, n, h
    sim = sim + bias.permute(2, 0, 1)
    if mask is not None:
        m = mask.permute(0, 1, 2, 5, 6, 3, 4).reshape(b * X * Y, 1, 1, L * w1 * w2)   # (b x y) 1 1 (l w1 w2)
        sim = sim.masked_fill(m == 0, -float("inf"))
    out = sim.softmax(-1) @ v
    out = out.permute(0, 2, 1, 3).reshape(b * X * Y, L, w1, w2, d)
    out = F.linear(out, sd[pre + ".to_out.0.weight"])
    return out.reshape(b, X, Y, L, w1, w2, d).permute(0, 3, 1, 2, 4, 5, 6)


class MaskedDropout:
    """nn.Dropout with GIVEN keep masks (test hook): the reference draws a fresh Bernoulli mask per call; a parity test
    feeds the masks the implementation under test used. masks: flat uint8 tensors in call order, each laid out like the
    implementation's token tensor [b*L, H, W, C]; `part` maps a [b, L, C, H, W] tensor to the layout of the tensor being
    dropped (the window / grid partition of the caller)."""

    def __init__(self, p, masks):
        self.p, self.masks, self.i = float(p), list(masks), 0

    def __call__(self, t, part, bLHW):
        b, L, H, W = bLHW
        m = self.masks[self.i]
        self.i += 1
        m5 = m.reshape(b, L, H, W, -1).permute(0, 1, 4, 2, 3).to(t.dtype)
        return t * part(m5) / (1.0 - self.p)


def feed_forward(sd, pre, x, drop=None):
    """base_transformer.py:16-28: Linear, GELU, Dropout, Linear, Dropout (drop: callable or None = identity)"""
    x = F.gelu(F.linear(x, sd[pre + ".net.0.weight"], sd[pre + ".net.0.bias"]))
    if drop is not None:
        x = drop(x)
    x = F.linear(x, sd[pre + ".net.3.weight"], sd[pre + ".net.3.bias"])
    return drop(x) if drop is not None else x


def _ln(sd, pre, x):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + ".weight"], sd[pre + ".bias"], 1e-5)


def swap_fusion_block(sd, pre, x, mask, w, dim_head, dropout=None):
    """x: [b, L, d, H, W]; mask: [b, H, W, 1, L]; SwapFusionBlockMask.forward :155-195. dropout: MaskedDropout or None;
    call order = the module's: Attention.to_out's Dropout (:43), FeedForward's two (base_transformer.py:32,34), per
    window then grid half."""
    b, L, d, H, W = x.shape
    X, Y = H // w, W // w
    part_w = lambda t: t.reshape(b, L, t.shape[2], X, w, Y, w).permute(0, 1, 3, 5, 4, 6, 2)
    part_g = lambda t: t.reshape(b, L, t.shape[2], w, X, w, Y).permute(0, 1, 4, 6, 3, 5, 2)
    dw = (lambda t: dropout(t, part_w, (b, L, H, W))) if dropout is not None else None
    dg = (lambda t: dropout(t, part_g, (b, L, H, W))) if dropout is not None else None
    ident = lambda t: t
    # window partition: (x w1) (y w2)
    xw = part_w(x)                                                                 # b L X Y w1 w2 d
    mw = mask.reshape(b, X, w, Y, w, 1, L).permute(0, 1, 3, 2, 4, 5, 6) if mask is not None else None
    xw = (dw or ident)(window_attention(sd, pre + ".window_attention.fn", _ln(sd, pre + ".window_attention.norm", xw), mw,
                                        dim_head)) + xw
    xw = feed_forward(sd, pre + ".window_ffd.fn", _ln(sd, pre + ".window_ffd.norm", xw), dw) + xw
    x = xw.permute(0, 1, 6, 2, 4, 3, 5).reshape(b, L, d, H, W)
    # grid partition: (w1 x) (w2 y)
    xg = part_g(x)                                                                 # b L X Y w1 w2 d
    mg = mask.reshape(b, w, X, w, Y, 1, L).permute(0, 2, 4, 1, 3, 5, 6) if mask is not None else None
    xg = (dg or ident)(window_attention(sd, pre + ".grid_attention.fn", _ln(sd, pre + ".grid_attention.norm", xg), mg,
                                        dim_head)) + xg
    xg = feed_forward(sd, pre + ".grid_ffd.fn", _ln(sd, pre + ".grid_ffd.norm", xg), dg) + xg
    return xg.permute(0, 1, 6, 4, 2, 5, 3).reshape(b, L, d, H, W)


def swap_fusion_encoder(sd, fa, x, mask, pre="fusion_net", keep=None, dropout=None):
    """SwapFusionEncoder.forward :277-280"""
    for i in range(fa["depth"]):
        x = swap_fusion_block(sd, "%s.layers.%d" % (pre, i), x, mask if fa.get("mask", False) else None,
                              fa["window_size"], fa["dim_head"], dropout)
        if keep is not None:
            keep["block%d" % i] = x
    x = x.mean(1).permute(0, 2, 3, 1)                                              # b h w d  (padded agents included)
    x = _ln(sd, pre + ".mlp_head.2", x)
    x = F.linear(x, sd[pre + ".mlp_head.3.weight"], sd[pre + ".mlp_head.3.bias"])
    return x.permute(0, 3, 1, 2)


def naive_compressor(sd, x, training, buffers, pre="naive_compressor"):
    """common_modules/naive_compress.py:38-42: encoder (conv3x3+BN+ReLU) then decoder (2 x conv3x3+BN+ReLU)"""
    for conv, bn in (("encoder.0", "encoder.1"), ("decoder.0", "decoder.1"), ("decoder.3", "decoder.4")):
        x = F.conv2d(x, sd["%s.%s.weight" % (pre, conv)], sd["%s.%s.bias" % (pre, conv)], padding=1)
        x = F.relu(O._bn(x, sd, "%s.%s" % (pre, bn), training, buffers))
    return x


def cobevt_forward(sd, args, data_dict, training=False, keep=None, dropout=None):
    """models/airv2x_cobevt.py:112-156 (task == det). dropout: None = nn.Dropout is the identity (eval mode / drop_out 0),
    or a MaskedDropout carrying the keep masks of every nn.Dropout call of the fusion network."""
    buffers = {}
    sf, record_len = O.extract_features(sd, args, data_dict, training, buffers, keep)
    feat = O.backbone_forward(sd, args["base_bev_backbone"], sf, training, buffers)
    if args["shrink_header"]["use"]:
        feat = O.shrink_conv(sd, args["shrink_header"], feat)
    if args.get("compression", 0):
        feat = naive_compressor(sd, feat, training, buffers)
    L = sum(args["max_cav"].values())
    x, mask = regroup(feat, record_len.tolist(), L)
    if keep is not None:
        keep["regroup"] = x
    H, W = x.shape[3], x.shape[4]
    com_mask = mask[:, None, None, None, :].expand(-1, H, W, 1, -1)                 # b h w 1 l
    fused = swap_fusion_encoder(sd, args["fax_fusion"], x, com_mask, keep=keep, dropout=dropout)
    if keep is not None:
        keep["fused_feature"] = fused
    out = {"psm": F.conv2d(fused, sd["cls_head.weight"], sd["cls_head.bias"]),
           "rm": F.conv2d(fused, sd["reg_head.weight"], sd["reg_head.bias"])}
    if args["obj_head"]:
        out["obj"] = F.conv2d(fused, sd["obj_head.weight"], sd["obj_head.bias"])
    return out, buffers


def _legacy_encoder(sd, args, data_dict, training, buffers):
    """the shared front of the legacy `point_pillar_*` models (point_pillar_cobevt.py:76-106): one PillarVFE, scatter,
    backbone, 3x3 stride-2 shrink header, optional NaiveCompressor. Returns (features [N,256,h,w], comm_rate)."""
    lid = data_dict[args.get("use_modality", "processed_lidar")]
    record_len = data_dict["record_len"]
    pf, _ = O.pillar_vfe(sd, "pillar_vfe", lid["voxel_features"], lid["voxel_num_points"], lid["voxel_coords"],
                         args["voxel_size"], args["lidar_range"], training, buffers)
    nx, ny, _ = [int(v) for v in args["point_pillar_scatter"]["grid_size"]]
    sf = O.scatter(pf, lid["voxel_coords"], nx, ny, int(record_len.sum()))
    comm_rate = int(sf.count_nonzero().item())
    feat = O.backbone_forward(sd, args["base_bev_backbone"], sf, training, buffers)
    if "shrink_header" in args:
        feat = O.shrink_conv(sd, args["shrink_header"], feat)
    if args["compression"] > 0:
        feat = naive_compressor(sd, feat, training, buffers)
    return feat, comm_rate, (ny, nx)


def pp_cobevt_forward(sd, args, data_dict, training=False, keep=None):
    """models/point_pillar_cobevt.py:76-128 (`point_pillar_cobevt`; dropout = identity)"""
    buffers = {}
    feat, comm_rate, _ = _legacy_encoder(sd, args, data_dict, training, buffers)
    x, mask = regroup(feat, data_dict["record_len"].tolist(), args["max_cav"])
    H, W = x.shape[3], x.shape[4]
    com_mask = mask[:, None, None, None, :].expand(-1, H, W, 1, -1)
    fused = swap_fusion_encoder(sd, args["fax_fusion"], x, com_mask, keep=keep)
    return {"psm": F.conv2d(fused, sd["cls_head.weight"], sd["cls_head.bias"]),
            "rm": F.conv2d(fused, sd["reg_head.weight"], sd["reg_head.bias"]),
            "mask": 0, "each_mask": 0, "comm_rate": comm_rate}, buffers
