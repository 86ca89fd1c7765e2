"""TEST INFRASTRUCTURE (oracle) — CPU fp32 restatement of the reference Where2comm detection path.

Every function cites the reference lines it follows (paths relative to the reference repo root). The restatement
is functional: it consumes a reference-format ``state_dict`` (same key names) and the reference-format
``data_dict`` and returns the same output dict. It is pinned against the real reference modules run in the build
container (scripts/make_golden.py -> tests/golden/*.npz); voxelisation (third-party spconv) is the one stage whose
parity is unpinned (see oracle/voxelize.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F

from . import voxelize as _V

AGENT_TYPES = ("vehicle", "rsu", "drone")
TYPE_PREFIX = {"vehicle": "veh_models", "rsu": "rsu_models", "drone": "drone_models"}
BN_EPS = 1e-3
BN_MOM = 0.01


# --------------------------------------------------------------------------------------------------- helpers
def _bn(x, sd, prefix, training, buffers):
    """BatchNorm (eps 1e-3, momentum 0.01; base_bev_backbone.py:52, airv2x_pillar_vfe.py:21).
    `buffers` collects updated running stats (dict) so the triple update per step can be checked."""
    rm = buffers.get(prefix + ".running_mean", sd[prefix + ".running_mean"]).clone()
    rv = buffers.get(prefix + ".running_var", sd[prefix + ".running_var"]).clone()
    y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training, BN_MOM, BN_EPS)
    if training:
        buffers[prefix + ".running_mean"] = rm
        buffers[prefix + ".running_var"] = rv
    return y


# --------------------------------------------------------------------------------------------------- a4/a5 PillarVFE
def pillar_vfe(sd, prefix, voxel_features, voxel_num_points, coords, voxel_size, lidar_range, training, buffers):
    """models/common_modules/airv2x_pillar_vfe.py:105-160 (feature build) and :27-49 (PFNLayer, last layer)."""
    vx, vy, vz = voxel_size
    x_off = vx / 2 + lidar_range[0]
    y_off = vy / 2 + lidar_range[1]
    z_off = vz / 2 + lidar_range[2]
    vf = voxel_features
    points_mean = vf[:, :, :3].sum(dim=1, keepdim=True) / voxel_num_points.type_as(vf).view(-1, 1, 1)
    f_cluster = vf[:, :, :3] - points_mean
    f_center = torch.zeros_like(vf[:, :, :3])
    f_center[:, :, 0] = vf[:, :, 0] - (coords[:, 3].to(vf.dtype).unsqueeze(1) * vx + x_off)
    f_center[:, :, 1] = vf[:, :, 1] - (coords[:, 2].to(vf.dtype).unsqueeze(1) * vy + y_off)
    f_center[:, :, 2] = vf[:, :, 2] - (coords[:, 1].to(vf.dtype).unsqueeze(1) * vz + z_off)
    feats = torch.cat([vf, f_cluster, f_center], dim=-1)  # use_absolute_xyz, no distance
    t = vf.shape[1]
    mask = (voxel_num_points.int().unsqueeze(1) > torch.arange(t, dtype=torch.int32, device=vf.device).view(1, -1)).unsqueeze(-1)
    feats = feats * mask.type_as(vf)
    x = F.linear(feats, sd[prefix + ".pfn_layers.0.linear.weight"])          # [M,32,64]
    x = _bn(x.permute(0, 2, 1), sd, prefix + ".pfn_layers.0.norm", training, buffers).permute(0, 2, 1)
    x = F.relu(x)
    return torch.max(x, dim=1)[0], feats                                      # [M,64]


# --------------------------------------------------------------------------------------------------- a6 scatter
def scatter(pillar_features, coords, nx, ny, n_agents):
    """models/common_modules/point_pillar_scatter.py:15-82 -> [n_agents, C, ny, nx]."""
    c = pillar_features.shape[1]
    out = []
    for b in range(n_agents):
        canvas = torch.zeros(c, nx * ny, dtype=pillar_features.dtype, device=pillar_features.device)
        m = coords[:, 0] == b
        tc = coords[m]
        idx = (tc[:, 1] + tc[:, 2] * nx + tc[:, 3]).long()
        canvas[:, idx] = pillar_features[m].t()
        out.append(canvas.view(c, ny, nx))
    return torch.stack(out, 0)


# --------------------------------------------------------------------------------------------------- a7 extract
def extract_features(sd, args, data_dict, training, buffers, keep=None, camera_bev=None):
    """models/common_modules/airv2x_base_model.py:101-248 (per-type encoders, repack to scene-major).
    camera_bev: {type: [n_type, 64, ny, nx]} outputs of that type's camera encoder (LiftSplatShootEncoder): `fuse_bev`
    (:167-177) averages the modalities' `spatial_features` per agent type."""
    per_type = {}
    for t in AGENT_TYPES:
        if t not in args["collaborators"] or len(data_dict[t]["batch_idxs"]) == 0:
            continue
        lid = data_dict[t]["batch_merged_lidar_features_torch"]
        la = args[t]["lidar"]
        prefix = TYPE_PREFIX[t] + ".0.0"
        pf, feats = pillar_vfe(sd, prefix, lid["voxel_features"], lid["voxel_num_points"], lid["voxel_coords"],
                               la["voxel_size"], la["lidar_range"], training, buffers)
        nx, ny, _ = [int(v) for v in la["point_pillar_scatter"]["grid_size"]]
        n_agents = int(lid["voxel_coords"][:, 0].max().item()) + 1
        per_type[t] = scatter(pf, lid["voxel_coords"], nx, ny, n_agents)
        if camera_bev is not None and t in camera_bev:
            per_type[t] = torch.stack([per_type[t], camera_bev[t]], 0).mean(0)
        if keep is not None:
            keep["pillar_features_" + t] = pf
    bsz = max(len(data_dict[t]["batch_idxs"]) for t in AGENT_TYPES)
    maps, record_len = [], []
    for b in range(bsz):
        n_b = 0
        for t, feat in per_type.items():
            idxs = data_dict[t]["batch_idxs"]
            if b not in idxs:
                continue
            ti = idxs.index(b)
            rl = data_dict[t]["record_len"]
            rl = rl[rl > 0]
            cs = torch.cumsum(rl, 0)
            start = 0 if ti == 0 else int(cs[ti - 1])
            maps.append(feat[start:int(cs[ti])])
            n_b += int(data_dict[t]["record_len"][b])
        record_len.append(n_b)
    return torch.cat(maps, 0), torch.tensor(record_len, dtype=torch.int32)


# --------------------------------------------------------------------------------------------------- a8 backbone
def backbone_block(sd, i, x, layer_num, training, buffers):
    """base_bev_backbone.py:41-72: ZeroPad2d(1)+Conv3x3 s2 +BN+ReLU, then layer_num x (Conv3x3 p1 +BN+ReLU)."""
    p = "backbone.blocks.%d" % i
    x = F.conv2d(F.pad(x, (1, 1, 1, 1)), sd[p + ".1.weight"], stride=2)
    x = F.relu(_bn(x, sd, p + ".2", training, buffers))
    for k in range(layer_num):
        x = F.conv2d(x, sd["%s.%d.weight" % (p, 4 + 3 * k)], padding=1)
        x = F.relu(_bn(x, sd, "%s.%d" % (p, 5 + 3 * k), training, buffers))
    return x


def backbone_deblock(sd, i, x, stride, training, buffers):
    """base_bev_backbone.py:73-90: ConvTranspose2d(k = s) + BN + ReLU."""
    p = "backbone.deblocks.%d" % i
    x = F.conv_transpose2d(x, sd[p + ".0.weight"], stride=stride)
    return F.relu(_bn(x, sd, p + ".1", training, buffers))


def backbone_forward(sd, bb_args, x, training, buffers):
    """base_bev_backbone.py:125-154."""
    ups = []
    for i, ln in enumerate(bb_args["layer_nums"]):
        x = backbone_block(sd, i, x, ln, training, buffers)
        ups.append(backbone_deblock(sd, i, x, bb_args["upsample_strides"][i], training, buffers))
    return torch.cat(ups, 1)


# --------------------------------------------------------------------------------------------------- a9 shrink
def shrink_conv(sd, sh_args, x):
    """downsample_conv.py:8-54 (conv+bias+ReLU, conv3x3+bias+ReLU per layer)."""
    for li, (k, s, pd) in enumerate(zip(sh_args["kernal_size"], sh_args["stride"], sh_args["padding"])):
        p = "shrink_conv.layers.%d.double_conv" % li
        x = F.relu(F.conv2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], stride=s, padding=pd))
        x = F.relu(F.conv2d(x, sd[p + ".2.weight"], sd[p + ".2.bias"], padding=1))
    return x


# --------------------------------------------------------------------------------------------------- a12 mask
def communication(sd, comm_args, psm_single, record_len, training, keep=None):
    """where2comm_modules/where2comm_fuse.py:83-149. Train mode draws K from Python `random` (:106)."""
    thr = comm_args["threshold"]
    masks, rates, smooths = [], [], []
    start = 0
    for n in record_len.tolist():
        conf = psm_single[start:start + n]
        start += n
        ori, _ = conf.sigmoid().max(dim=1, keepdim=True)
        if "gaussian_smooth" in comm_args:
            k = comm_args["gaussian_smooth"]["k_size"]
            maps = F.conv2d(ori, sd["fusion_net.naive_communication.gaussian_filter.weight"],
                            sd["fusion_net.naive_communication.gaussian_filter.bias"], padding=(k - 1) // 2)
        else:
            maps = ori
        L, _, H, W = maps.shape
        if training:
            K = int(H * W * random.uniform(0, 1))
            flat = maps.reshape(L, H * W)
            _, idx = torch.topk(flat, k=K, sorted=False)
            mask = torch.zeros_like(flat).scatter(-1, idx, torch.ones(L, K, device=flat.device)).reshape(L, 1, H, W)
        elif thr:
            mask = (maps > thr).to(maps.dtype)
        else:
            mask = torch.ones_like(maps)
        rates.append(mask.sum() / (L * H * W))
        mask = mask.clone()
        mask[0] = 1
        masks.append(mask)
        smooths.append(maps)
    if keep is not None:
        keep["smooth"] = torch.cat(smooths, 0)
    return torch.cat(masks, 0), sum(rates) / len(rates)


# --------------------------------------------------------------------------------------------------- a14 fusion
def attention_fusion(x):
    """where2comm_fuse.py:152-164 + :14-45: per-pixel scaled dot-product attention over agents, ego row only."""
    n, c, h, w = x.shape
    q = x.view(n, c, -1).permute(2, 0, 1)
    score = torch.bmm(q, q.transpose(1, 2)) / np.sqrt(c)
    ctx = torch.bmm(F.softmax(score, -1), q)
    return ctx.permute(1, 2, 0).view(n, c, h, w)[0]


def where2comm_fusion(sd, args, spatial_features, psm_single, record_len, training, buffers, keep=None):
    """where2comm_fuse.py:198-263, multi-scale branch."""
    fa = args["where2com_fusion"]
    bb = args["modality_fusion"]["base_bev_backbone"]
    x = spatial_features
    ups = []
    rate = None
    for i, ln in enumerate(fa["layer_nums"]):
        x = backbone_block(sd, i, x, ln, training, buffers)
        if i == 0:
            if fa["fully"]:
                rate = torch.tensor(1)
            else:
                mask, rate = communication(sd, fa["communication"], psm_single, record_len, training, keep)
                if keep is not None and keep.get("mask_override") is not None:
                    # test hook: teacher-force the (discontinuous) top-K / threshold decisions of the implementation
                    # under test, so the comparison downstream of the mask is well-posed (tests check separately that
                    # the two masks differ only at pixels whose smoothed confidence ties with the cut)
                    keep["mask_own"] = mask
                    mask = keep["mask_override"].to(mask.dtype).reshape(mask.shape)
                if x.shape[-1] != mask.shape[-1]:
                    mask = F.interpolate(mask, size=(x.shape[-2], x.shape[-1]), mode="bilinear", align_corners=False)
                if keep is not None:
                    keep["mask"] = mask
                x = x * mask
        fused, start = [], 0
        for n in record_len.tolist():
            fused.append(attention_fusion(x[start:start + n]))
            start += n
        xf = torch.stack(fused)
        if keep is not None:
            keep["fused_l%d" % i] = xf
        ups.append(backbone_deblock(sd, i, xf, bb["upsample_strides"][i], training, buffers))
    return torch.cat(ups, 1), rate


# --------------------------------------------------------------------------------------------------- model forward
def where2com_forward(sd, args, data_dict, training=False, keep=None, camera_bev=None):
    """models/airv2x_where2com.py:117-179 (task == det). Returns (output_dict, updated BN buffers)."""
    buffers = {}
    mf = args["modality_fusion"]
    sf, record_len = extract_features(sd, args, data_dict, training, buffers, keep, camera_bev)
    if keep is not None:
        keep["spatial_features"] = sf
        keep["record_len"] = record_len
    comm_rates = int(sf.count_nonzero().item())
    feat2d = backbone_forward(sd, mf["base_bev_backbone"], sf, training, buffers)       # :119
    feat2d = backbone_forward(sd, mf["base_bev_backbone"], sf, training, buffers)       # :124 (recomputed)
    if keep is not None:
        keep["spatial_features_2d"] = feat2d
    if mf["shrink_header"]["use"]:
        feat2d = shrink_conv(sd, mf["shrink_header"], feat2d)
    psm_single = F.conv2d(feat2d, sd["cls_head.weight"], sd["cls_head.bias"])
    if keep is not None:
        keep["psm_single"] = psm_single
    fused, rate = where2comm_fusion(sd, args, sf, psm_single, record_len, training, buffers, keep)
    if mf["shrink_header"]["use"]:
        fused = shrink_conv(sd, mf["shrink_header"], fused)
    if keep is not None:
        keep["fused_feature"] = fused
    out = {"psm": F.conv2d(fused, sd["cls_head.weight"], sd["cls_head.bias"]),
           "rm": F.conv2d(fused, sd["reg_head.weight"], sd["reg_head.bias"])}
    if args["obj_head"]:
        out["obj"] = F.conv2d(fused, sd["obj_head.weight"], sd["obj_head.bias"])
    out.update({"mask": 0, "com": rate, "comm_rate": comm_rates})
    return out, buffers


# --------------------------------------------------------------------------------------------------- a19 loss
def point_pillar_loss_multiclass(output, target, num_class, cls_weight=1.0, reg_coe=2.0):
    """loss/point_pillar_loss_multiclass.py:96-215, :273-289. Returns (total, reg, cls, obj)."""
    rm, psm, obj = output["rm"], output["psm"], output["obj"]
    B = psm.shape[0]
    targets = target["targets"]
    cls_preds = psm.permute(0, 2, 3, 1).contiguous()
    obj_preds = obj.permute(0, 2, 3, 1).contiguous()
    labels = target["pos_equal_one"].view(B, -1)
    positives = labels > 0
    negatives = labels == 0
    cls_weights = (negatives * 1.0 + 1.0 * positives).float()
    reg_weights = positives.float()
    pos_norm = positives.sum(1, keepdim=True).float()
    reg_weights = reg_weights / torch.clamp(pos_norm, min=1.0)
    cls_weights = cls_weights / torch.clamp(pos_norm, min=1.0)
    cls_targets = target["class_ids"]
    one_hot = torch.zeros(*cls_targets.shape, num_class, dtype=cls_preds.dtype, device=cls_preds.device)
    one_hot.scatter_(-1, cls_targets.unsqueeze(-1).long(), 1.0)
    _, H, W, AC = cls_preds.shape
    A = AC // num_class
    inp = cls_preds.view(B, H, W, A, num_class)
    tgt = one_hot.view(B, H, W, A, num_class)
    wts = cls_weights.view(B, H, W, A, 1)
    ps = torch.sigmoid(inp)
    alpha_w = tgt * 0.25 + (1 - tgt) * 0.75
    pt = tgt * (1.0 - ps) + (1.0 - tgt) * ps
    focal = alpha_w * torch.pow(pt, 2.0)
    bce = torch.clamp(inp, min=0) - inp * tgt + torch.log1p(torch.exp(-torch.abs(inp)))
    cls_loss_src = (focal * bce * wts).sum() / B          # cls_loss_func already divides by B (:215)
    conf_loss = cls_loss_src.sum() / B * cls_weight        # and forward divides again (:148)
    rmv = rm.permute(0, 2, 3, 1).contiguous().view(B, -1, 7)
    tg = targets.view(B, -1, 7)
    p_sin = torch.sin(rmv[..., 6:7]) * torch.cos(tg[..., 6:7])
    t_sin = torch.cos(rmv[..., 6:7]) * torch.sin(tg[..., 6:7])
    b1 = torch.cat([rmv[..., :6], p_sin], -1)
    b2 = torch.cat([tg[..., :6], t_sin], -1)
    b2 = torch.where(torch.isnan(b2), b1, b2)
    n = torch.abs(b1 - b2)
    beta = 1.0 / 9.0
    sl1 = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta) * reg_weights.unsqueeze(-1)
    reg_loss = sl1.sum() / B * reg_coe
    pos_mask = target["pos_equal_one"]
    osig = torch.sigmoid(obj_preds)
    obj_loss = (-(pos_mask * torch.log(osig + 1e-6) + (1 - pos_mask) * torch.log(1 - osig + 1e-6))).mean()
    return reg_loss + conf_loss + obj_loss, reg_loss, conf_loss, obj_loss


def point_pillar_loss(output, target, cls_weight=1.0, reg_coe=2.0):
    """loss/point_pillar_loss.py:77-166, :168-215 — the 1-class loss of the legacy `point_pillar_*` models: sigmoid focal
    loss on `psm` (alpha .25, gamma 2; one logit per anchor, normalised by the positives of each sample, summed and
    divided by B once — the multi-class variant divides twice), smooth-L1 (beta 1/9, sin-difference on yaw) on `rm`,
    no objectness term. Returns (total, reg, conf)."""
    rm, psm = output["rm"], output["psm"]
    B = psm.shape[0]
    cls_preds = psm.permute(0, 2, 3, 1).contiguous()
    labels = target["pos_equal_one"].view(B, -1).contiguous()
    positives = labels > 0
    negatives = labels == 0
    cls_weights = (negatives * 1.0 + 1.0 * positives).float()
    reg_weights = positives.float()
    pos_norm = positives.sum(1, keepdim=True).float()
    reg_weights = reg_weights / torch.clamp(pos_norm, min=1.0)
    cls_weights = cls_weights / torch.clamp(pos_norm, min=1.0)
    one_hot = torch.zeros(*labels.shape, 2, dtype=cls_preds.dtype, device=labels.device)
    one_hot.scatter_(-1, labels.unsqueeze(-1).long(), 1.0)
    inp = cls_preds.view(B, -1, 1)
    tgt = one_hot[..., 1:]
    ps = torch.sigmoid(inp)
    alpha_w = tgt * 0.25 + (1 - tgt) * 0.75
    pt = tgt * (1.0 - ps) + (1.0 - tgt) * ps
    focal = alpha_w * torch.pow(pt, 2.0)
    bce = torch.clamp(inp, min=0) - inp * tgt + torch.log1p(torch.exp(-torch.abs(inp)))
    conf_loss = (focal * bce * cls_weights.unsqueeze(-1)).sum() / B * cls_weight
    rmv = rm.permute(0, 2, 3, 1).contiguous().view(B, -1, 7)
    tg = target["targets"].view(B, -1, 7)
    p_sin = torch.sin(rmv[..., 6:7]) * torch.cos(tg[..., 6:7])
    t_sin = torch.cos(rmv[..., 6:7]) * torch.sin(tg[..., 6:7])
    b1 = torch.cat([rmv[..., :6], p_sin], -1)
    b2 = torch.cat([tg[..., :6], t_sin], -1)
    b2 = torch.where(torch.isnan(b2), b1, b2)
    n = torch.abs(b1 - b2)
    beta = 1.0 / 9.0
    sl1 = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta) * reg_weights.unsqueeze(-1)
    reg_loss = sl1.sum() / B * reg_coe
    return reg_loss + conf_loss, reg_loss, conf_loss


# --------------------------------------------------------------------------------------------------- legacy model
def pp_where2comm_forward(sd, args, data_dict, training=False, keep=None):
    """models/point_pillar_where2comm.py:99-151 (`point_pillar_where2comm`, BASELINE config 1; multi_scale branch):
    one PillarVFE for all agents, the backbone evaluated once, a stride-2 shrink header (so the communication mask
    lives at half the resolution of the level-0 features and is bilinearly resized, where2comm_fuse.py:230-236),
    1-class heads without objectness."""
    buffers = {}
    lid = data_dict[args.get("use_modality", "processed_lidar")]
    record_len = data_dict["record_len"]
    pf, _ = pillar_vfe(sd, "pillar_vfe", lid["voxel_features"], lid["voxel_num_points"], lid["voxel_coords"],
                       args["voxel_size"], args["lidar_range"], training, buffers)
    nx, ny, _ = [int(v) for v in args["point_pillar_scatter"]["grid_size"]]
    sf = scatter(pf, lid["voxel_coords"], nx, ny, int(record_len.sum()))
    comm_rates = int(sf.count_nonzero().item())
    feat2d = backbone_forward(sd, args["base_bev_backbone"], sf, training, buffers)
    if "shrink_header" in args:
        feat2d = shrink_conv(sd, args["shrink_header"], feat2d)
    psm_single = F.conv2d(feat2d, sd["cls_head.weight"], sd["cls_head.bias"])
    assert not args["compression"] and args["where2comm_fusion"]["multi_scale"]
    pseudo = {"where2com_fusion": args["where2comm_fusion"], "modality_fusion": {"base_bev_backbone": args["base_bev_backbone"]}}
    fused, rate = where2comm_fusion(sd, pseudo, sf, psm_single, record_len, training, buffers, keep)
    if "shrink_header" in args:
        fused = shrink_conv(sd, args["shrink_header"], fused)
    out = {"psm": F.conv2d(fused, sd["cls_head.weight"], sd["cls_head.bias"]),
           "rm": F.conv2d(fused, sd["reg_head.weight"], sd["reg_head.bias"])}
    out.update({"com": rate, "mask": 0, "each_mask": 0, "comm_rate": comm_rates})
    return out, buffers


def make_scene_legacy(preprocess, n_agents, n_points, seed, max_voxels, sigma_xy=(8.0, 8.0)):
    """`processed_lidar` dict of the legacy datasets: all agents of the scene collated into one voxel batch"""
    rng = preprocess["cav_lidar_range"]
    per = []
    for k in range(n_agents):
        pts = synth_points(seed * 100 + k, n_points, rng, sigma_xy=sigma_xy)
        pts = _V.mask_points(pts, rng, ego_box=(k == 0))
        per.append(_V.voxelize(pts, rng, preprocess["args"]["voxel_size"], preprocess["args"]["max_points_per_voxel"], max_voxels))
    col = _V.collate(per)
    L = 5
    return {"processed_lidar": {k: torch.from_numpy(v) for k, v in col.items()},
            "record_len": torch.tensor([n_agents], dtype=torch.int32),
            "pairwise_t_matrix": torch.eye(4).view(1, 1, 1, 4, 4).repeat(1, L, L, 1, 1)}


# --------------------------------------------------------------------------------------------------- synthetic data
def synth_points(seed, n_points, lidar_range, sigma_xy=(35.0, 15.0)):
    """SURVEY §8d synthetic cloud: x~N(0,sx), y~N(0,sy) clipped to the range, z~U(zlo,zhi), intensity~U(0,1)."""
    rng = np.random.default_rng(seed)
    r = lidar_range
    x = np.clip(rng.normal(0.0, sigma_xy[0], n_points), r[0] + 1e-3, r[3] - 1e-3)
    y = np.clip(rng.normal(0.0, sigma_xy[1], n_points), r[1] + 1e-3, r[4] - 1e-3)
    z = rng.uniform(r[2] + 1e-3, r[5] - 1e-3, n_points)
    i = rng.uniform(0.0, 1.0, n_points)
    return np.stack([x, y, z, i], 1).astype(np.float32)


def det_init_state_dict(shapes, seed=0):
    """Deterministic, size-aware parameter init shared by golden generation and tests (no 29 MB fixture):
    conv/linear weights ~ U(-b, b) with b = sqrt(3 / fan_in); BN gamma ~ U(0.8, 1.2), beta ~ U(-0.1, 0.1),
    running_mean ~ U(-0.1, 0.1), running_var ~ U(0.8, 1.2); biases ~ U(-0.1, 0.1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        shp = tuple(shp)
        u = torch.rand(shp, generator=g) if len(shp) else torch.rand((), generator=g)
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_var"):
            sd[k] = 0.8 + 0.4 * u
        elif k.endswith("running_mean"):
            sd[k] = -0.1 + 0.2 * u
        elif "gaussian_filter" in k:
            continue  # keep the module's analytic init
        elif len(shp) == 1 and k.endswith("weight"):
            sd[k] = 0.8 + 0.4 * u
        elif len(shp) == 1:
            sd[k] = -0.1 + 0.2 * u
        else:
            fan_in = int(np.prod(shp[1:]))
            if "deblocks" in k and k.endswith("0.weight"):
                fan_in = shp[0]
            b = math.sqrt(3.0 / fan_in)
            sd[k] = (2 * u - 1) * b
    return sd


def gaussian_filter_params(k_size=5, sigma=1.0):
    """where2comm_fuse.py:66-81 (note the 1/(2*pi*sigma) normalisation)."""
    c = k_size // 2
    x, y = np.mgrid[0 - c:k_size - c, 0 - c:k_size - c]
    g = 1 / (2 * np.pi * sigma) * np.exp(-(np.square(x) + np.square(y)) / (2 * np.square(sigma)))
    return torch.Tensor(g).unsqueeze(0).unsqueeze(0), torch.zeros(1)


def make_scene(preprocess, agents, n_points, seed, max_voxels, sigma_xy=(10.0, 5.0)):
    """agents: list of types in reference order (vehicles, rsus, drones). One scene (B = 1)."""
    pre = preprocess
    rng = pre["cav_lidar_range"]
    dd = {}
    k = 0
    for t in AGENT_TYPES:
        n_t = sum(1 for a in agents if a == t)
        per = []
        for _ in range(n_t):
            pts = synth_points(seed * 100 + k, n_points, rng, sigma_xy=sigma_xy)
            pts = _V.mask_points(pts, rng, ego_box=(k == 0))
            per.append(_V.voxelize(pts, rng, pre["args"]["voxel_size"], pre["args"]["max_points_per_voxel"], max_voxels))
            k += 1
        if n_t == 0:
            dd[t] = {"batch_merged_lidar_features_torch": None, "record_len": torch.tensor([0], dtype=torch.int32),
                     "batch_idxs": []}
            continue
        col = _V.collate(per)
        dd[t] = {"batch_merged_lidar_features_torch": {k2: torch.from_numpy(v) for k2, v in col.items()},
                 "record_len": torch.tensor([n_t], dtype=torch.int32), "batch_idxs": [0]}
    L = 15
    dd["img_pairwise_t_matrix_collab"] = torch.eye(4).view(1, 1, 1, 4, 4).repeat(1, L, L, 1, 1)
    dd["record_len"] = torch.tensor([len(agents)], dtype=torch.int32)
    return dd


def make_labels(seed, B, H, W, A, n_pos=30):
    g = torch.Generator().manual_seed(seed)
    pos = torch.zeros(B, H, W, A, dtype=torch.float64)
    idx = torch.randperm(B * H * W * A, generator=g)[:n_pos]
    pos.view(-1)[idx] = 1.0
    neg = 1.0 - pos
    ign = torch.randperm(B * H * W * A, generator=g)[:n_pos * 3]
    neg.view(-1)[ign] = 0.0
    targets = (torch.randn(B, H, W, A * 7, generator=g) * 0.3).double() * pos.repeat_interleave(7, dim=-1)
    class_ids = torch.zeros(B, H, W, A, dtype=torch.int64)
    class_ids.view(-1)[idx] = torch.randint(1, 7, (n_pos,), generator=g)
    return {"targets": targets, "pos_equal_one": pos, "neg_equal_one": neg, "class_ids": class_ids}


