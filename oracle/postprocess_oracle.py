"""TEST INFRASTRUCTURE ONLY: CPU restatement (numpy / torch fp32, polygon clipping in float64) of the reference's
detection post-processing and AP evaluation. Only tests/, __graft_entry__.smoke() and scripts/ may import it.

Follows:
  opencood/data_utils/post_processor/voxel_postprocessor.py:33-86    generate_anchor_box
  opencood/data_utils/post_processor/voxel_postprocessor.py:585-635  delta_to_boxes3d
  opencood/data_utils/post_processor/voxel_postprocessor.py:666-840  post_process_airv2x (ego only, identity transform)
  opencood/utils/box_utils.py:195-258, :399-430, :823-868, :981-1035 corners, range mask, nms_rotated, size / z filters
  opencood/utils/common_utils.py:150-191                             polygon IoU (shapely there; shapely is absent here, so
                                                                     the IoU is Sutherland-Hodgman clipping of convex quads:
                                                                     parity of that ONE function is unpinned against shapely
                                                                     and pinned against closed-form cases instead)
  opencood/utils/eval_utils_opv2v.py:15-152                          voc_ap, caluclate_tp_fp, calculate_ap
Pinned by scripts/make_golden_postprocess.py against the real reference functions (decode, corners, filters, the greedy
NMS loop, TP/FP matching and AP with the IoU routine injected).
"""
import math

import numpy as np
import torch


def generate_anchor_box(aa, order="hwl"):
    r = [math.radians(e) for e in aa["r"]]
    A = len(r)
    rng = aa["cav_lidar_range"]
    stride = aa.get("feature_stride", 2)
    x = np.linspace(rng[0] + aa["vw"], rng[3] - aa["vw"], aa["W"] // stride)
    y = np.linspace(rng[1] + aa["vh"], rng[4] - aa["vh"], aa["H"] // stride)
    cx, cy = np.meshgrid(x, y)
    cx, cy = np.tile(cx[..., None], A), np.tile(cy[..., None], A)
    cz = np.ones_like(cx) * -1.0
    w, l, h = np.ones_like(cx) * aa["w"], np.ones_like(cx) * aa["l"], np.ones_like(cx) * aa["h"]
    r_ = np.ones_like(cx)
    for i in range(A):
        r_[..., i] = r[i]
    assert order == "hwl"
    return np.stack([cx, cy, cz, h, w, l, r_], -1)


def delta_to_boxes3d(rm, anchors):
    N = rm.shape[0]
    d = rm.permute(0, 2, 3, 1).contiguous().view(N, -1, 7)
    a = anchors.view(-1, 7).float()
    ad = torch.sqrt(a[:, 4] ** 2 + a[:, 5] ** 2)
    b = torch.zeros_like(d)
    b[..., 0] = d[..., 0] * ad + a[:, 0]
    b[..., 1] = d[..., 1] * ad + a[:, 1]
    b[..., 2] = d[..., 2] * a[:, 3] + a[:, 2]
    b[..., 3:6] = torch.exp(d[..., 3:6]) * a[:, 3:6]
    b[..., 6] = d[..., 6] + a[:, 6]
    return b


def boxes_to_corners_3d(b):
    """order 'hwl' -> (N, 8, 3)"""
    b_ = b[:, [0, 1, 2, 5, 4, 3, 6]]
    t = b_.new_tensor([[1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, -1], [1, -1, 1], [1, 1, 1], [-1, 1, 1], [-1, -1, 1]]) / 2
    c = b_[:, None, 3:6].repeat(1, 8, 1) * t[None]
    cosa, sina = torch.cos(b_[:, 6]), torch.sin(b_[:, 6])
    z, o = torch.zeros_like(cosa), torch.ones_like(cosa)
    rot = torch.stack((cosa, sina, z, -sina, cosa, z, z, z, o), 1).view(-1, 3, 3).float()
    return torch.matmul(c, rot) + b_[:, None, 0:3]


def poly_area(p):
    x, y = p[:, 0], p[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def quad_iou(a, b):
    """IoU of two convex quads [(4,2) float64]: Sutherland-Hodgman clip of a by b"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    sa, sb = abs(poly_area(a)), poly_area(b)
    if sb < 0:
        b = b[::-1]
        sb = -sb
    poly = [tuple(p) for p in a]
    for e in range(4):
        x1, y1 = b[e]
        x2, y2 = b[(e + 1) % 4]
        ex, ey = x2 - x1, y2 - y1
        out = []
        for i in range(len(poly)):
            px, py = poly[i]
            qx, qy = poly[(i + 1) % len(poly)]
            dp = ex * (py - y1) - ey * (px - x1)
            dq = ex * (qy - y1) - ey * (qx - x1)
            if dp >= 0:
                out.append((px, py))
            if (dp >= 0) != (dq >= 0):
                t = dp / (dp - dq)
                out.append((px + t * (qx - px), py + t * (qy - py)))
        poly = out
        if not poly:
            break
    inter = abs(poly_area(np.array(poly))) if len(poly) >= 3 else 0.0
    uni = sa + sb - inter
    return inter / uni if uni > 0 else 0.0


def compute_iou(box, boxes):
    return np.array([quad_iou(box, b) for b in boxes], dtype=np.float32)


def nms_rotated(corners, scores, threshold, top=1000):
    """box_utils.py:823-868 with polygons = first four corners' xy"""
    if corners.shape[0] == 0:
        return np.array([], dtype=np.int32)
    polys = corners[:, :4, :2].detach().cpu().numpy().astype(np.float64)
    s = scores.detach().cpu().numpy()
    ixs = s.argsort()[::-1][:top]
    pick = []
    while len(ixs) > 0:
        i = ixs[0]
        pick.append(i)
        iou = compute_iou(polys[i], polys[ixs[1:]])
        remove = np.where(iou > threshold)[0] + 1
        ixs = np.delete(ixs, remove)
        ixs = np.delete(ixs, 0)
    return np.array(pick, dtype=np.int32)


def post_process(out, params):
    """post_process_airv2x for the ego (identity transformation_matrix). Returns corners, scores, labels, boxes3d,
    anchor indices (all torch, score order of the NMS picks) or five Nones."""
    aa = params["anchor_args"]
    anchors = torch.from_numpy(generate_anchor_box(aa, params["order"]))
    C = aa.get("num_class", 7)
    rng = aa["cav_lidar_range"]
    obj = out["obj"].permute(0, 2, 3, 1).contiguous()
    objectness = torch.sigmoid(obj).view(1, -1)
    psm = out["psm"]
    B, AC, H, W = psm.shape
    A = AC // C
    prob = torch.sigmoid(psm.view(B, C, A, H, W).permute(0, 3, 4, 2, 1).contiguous()).view(1, -1, C)[:, :, 1:]
    _, labels = torch.max(prob, -1)
    labels = labels + 1
    mask = objectness > params["target_args"]["obj_threshold"]
    if mask.sum() == 0:
        return (None,) * 5
    boxes = delta_to_boxes3d(out["rm"], anchors)[0]
    idx = torch.nonzero(mask[0]).squeeze(1)
    boxes3d, scores, lab = boxes[idx], objectness[0][idx], labels[0][idx]
    corners = boxes_to_corners_3d(boxes3d)
    x_len = corners[:, :, 0].max(1)[0] - corners[:, :, 0].min(1)[0]
    y_len = corners[:, :, 1].max(1)[0] - corners[:, :, 1].min(1)[0]
    z_len = corners[:, :, 2].max(1)[0] - corners[:, :, 2].min(1)[0]
    keep = (x_len <= 6) & (y_len <= 6) & z_len.bool()
    keep &= (corners[:, :, 2].min(1)[0] >= rng[2]) & (corners[:, :, 2].max(1)[0] <= rng[5])
    corners, scores, lab, boxes3d, idx = corners[keep], scores[keep], lab[keep], boxes3d[keep], idx[keep]
    pick = torch.from_numpy(nms_rotated(corners, scores, params["nms_thresh"]).astype(np.int64))
    corners, scores, lab, boxes3d, idx = corners[pick], scores[pick], lab[pick], boxes3d[pick], idx[pick]
    lo = torch.tensor(rng[:2]).view(1, 1, -1)
    hi = torch.tensor(rng[3:5]).view(1, 1, -1)
    m = torch.all(torch.all(corners[:, :, :2] >= lo, -1) & torch.all(corners[:, :, :2] <= hi, -1), -1)
    return corners[m], scores[m], lab[m], boxes3d[m], idx[m]


def voc_ap(rec, prec):
    mrec = [0.0] + list(rec) + [1.0]
    mpre = [0.0] + list(prec) + [0.0]
    for i in range(len(mpre) - 2, -1, -1):
        mpre[i] = max(mpre[i], mpre[i + 1])
    ap = 0.0
    for i in range(1, len(mrec)):
        if mrec[i] != mrec[i - 1]:
            ap += (mrec[i] - mrec[i - 1]) * mpre[i]
    return ap


def tp_fp(det_corners, det_score, gt_corners, stat, thr):
    fp, tp = [], []
    gt = gt_corners.shape[0]
    if det_corners is not None:
        score = det_score.detach().cpu().numpy()
        order = np.argsort(-score)
        dets = det_corners[:, :4, :2].detach().cpu().numpy().astype(np.float64)
        gts = [g for g in gt_corners[:, :4, :2].detach().cpu().numpy().astype(np.float64)]
        for i in order:
            ious = compute_iou(dets[i], gts)
            if len(gts) == 0 or np.max(ious) < thr:
                fp.append(1)
                tp.append(0)
                continue
            fp.append(0)
            tp.append(1)
            gts.pop(int(np.argmax(ious)))
        stat[thr]["score"] += score[order].tolist()
    stat[thr]["fp"] += fp
    stat[thr]["tp"] += tp
    stat[thr]["gt"] += gt


def calculate_ap(stat, thr):
    fp, tp = np.cumsum(stat[thr]["fp"]).tolist(), np.cumsum(stat[thr]["tp"]).tolist()
    rec = [float(t) / stat[thr]["gt"] for t in tp]
    prec = [float(t) / (f + t) for f, t in zip(fp, tp)]
    return voc_ap(rec, prec)


# ------------------------------------------------------------------------------------------------ synthetic frames
def synth_frame(params, seed, n_gt=12):
    """planted ground-truth boxes + head tensors that detect most of them (several overlapping candidates per object,
    some false positives, a few misses) — deterministic in `seed`. Returns (output_dict, gt_corners [n,8,3])."""
    g = torch.Generator().manual_seed(seed)
    aa = params["anchor_args"]
    anchors = torch.from_numpy(generate_anchor_box(aa, params["order"])).float()
    H, W, A, _ = anchors.shape
    C = aa.get("num_class", 7)
    rng = aa["cav_lidar_range"]
    obj = -4.0 + 0.5 * torch.randn(1, A, H, W, generator=g)
    rm = 0.05 * torch.randn(1, 7 * A, H, W, generator=g)
    psm = torch.randn(1, A * C, H, W, generator=g)
    gts = []
    flat = anchors.view(-1, 7)
    ad = torch.sqrt(flat[:, 4] ** 2 + flat[:, 5] ** 2)
    for k in range(n_gt):
        u = torch.rand(8, generator=g)
        x = rng[0] + 6 + float(u[0]) * (rng[3] - rng[0] - 12)
        y = rng[1] + 5 + float(u[1]) * (rng[4] - rng[1] - 10)
        yaw = (float(u[2]) - 0.5) * 0.6 + (math.pi / 2 if u[3] > 0.5 else 0.0)
        s = 0.8 + 0.4 * float(u[4])
        box = torch.tensor([x, y, -1.0 + 0.2 * (float(u[5]) - 0.5), aa["h"] * s, aa["w"] * s, aa["l"] * s, yaw])
        gts.append(box)
        if k % 5 == 4:
            continue  # a miss
        a_sel = 1 if u[3] > 0.5 else 0
        dist = (flat[:, 0] - x) ** 2 + (flat[:, 1] - y) ** 2 + (flat.view(H, W, A, 7)[..., 6].reshape(-1) != flat[a_sel, 6]).float() * 1e6
        for rank, fi in enumerate(torch.topk(-dist, 4).indices.tolist()):
            pix, a = fi // A, fi % A
            h_, w_ = pix // W, pix % W
            noise = 0.5 + 0.5 * rank + 1.5 * float(u[6])
            tgt = box.clone()
            tgt[0] += noise * (float(torch.rand((), generator=g)) - 0.5)
            tgt[1] += noise * (float(torch.rand((), generator=g)) - 0.5)
            tgt[3:6] *= 1.0 + 0.3 * (float(torch.rand((), generator=g)) - 0.5)
            tgt[6] += 0.2 * (float(torch.rand((), generator=g)) - 0.5)
            d = torch.empty(7)
            d[0] = (tgt[0] - flat[fi, 0]) / ad[fi]
            d[1] = (tgt[1] - flat[fi, 1]) / ad[fi]
            d[2] = (tgt[2] - flat[fi, 2]) / flat[fi, 3]
            d[3:6] = torch.log(tgt[3:6] / flat[fi, 3:6])
            d[6] = tgt[6] - flat[fi, 6]
            rm[0, a * 7:(a + 1) * 7, h_, w_] = d
            obj[0, a, h_, w_] = 3.0 - 0.7 * rank + 0.2 * float(torch.rand((), generator=g))
    for _ in range(10):  # false positives
        fi = int(torch.randint(0, flat.shape[0], (1,), generator=g))
        pix, a = fi // A, fi % A
        obj[0, a, pix // W, pix % W] = float(torch.rand((), generator=g)) * 1.5 - 0.5
    gt_corners = boxes_to_corners_3d(torch.stack(gts))
    return {"psm": psm, "rm": rm, "obj": obj}, gt_corners
