"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's anchor-target assignment. Only tests/,
__graft_entry__.smoke() and scripts/ may import it; the product path never does. Pinned against the REAL reference
(VoxelPostprocessor.generate_label_airv2x + the Cython bbox_overlaps built from the reference's .pyx in a temp dir) by
scripts/make_golden_labels.py: pos / neg / class maps identical, targets identical (float64).

Follows:
  opencood/data_utils/post_processor/voxel_postprocessor.py:33-86    generate_anchor_box (oracle/postprocess_oracle.py)
  opencood/data_utils/post_processor/voxel_postprocessor.py:217-354  generate_label_airv2x
  opencood/data_utils/post_processor/voxel_postprocessor.py:392-430  collate_batch_airv2x
  opencood/utils/box_overlaps.pyx:17-56                              bbox_overlaps (float32, the "+1" pixel convention)
  opencood/utils/box_utils.py:195-258, :279-302                      boxes_to_corners_3d (fp32 torch), standup boxes
"""
import numpy as np
import torch

from . import postprocess_oracle as PO


def standup_boxes(boxes7):
    """[n,7] (x,y,z,h,w,l,yaw) any float dtype -> float32 [n,4] (xmin, ymin, xmax, ymax) of the rotated footprint.
    The corners go through fp32 torch ops (check_numpy_to_torch(...).float(), box_utils.py:230, common_utils.py:60-82)."""
    b = np.asarray(boxes7)
    if b.shape[0] == 0:
        return np.zeros((0, 4), np.float32)
    corners = PO.boxes_to_corners_3d(torch.from_numpy(b).float()).numpy()          # [n, 8, 3] fp32
    out = np.zeros((b.shape[0], 4))
    out[:, 0], out[:, 1] = corners[:, :, 0].min(1), corners[:, :, 1].min(1)
    out[:, 2], out[:, 3] = corners[:, :, 0].max(1), corners[:, :, 1].max(1)
    return np.ascontiguousarray(out).astype(np.float32)


def bbox_overlaps(boxes, query):
    """box_overlaps.pyx:17-56 with the arithmetic the Cython build performs (Cython 3 emits the `+ 1` as the C double
    literal 1.0, so those sums and the products around them run in double before being stored to the float variables):
      box_area = float( ((q2-q0)_f32 + 1.0) * ((q3-q1)_f32 + 1.0) )          iw, ih = float( (min-max)_f32 + 1.0 )
      ua = float( ((b2-b0)_f32 + 1.0) * ((b3-b1)_f32 + 1.0) + box_area - (iw*ih)_f32 )      iou = (iw*ih)_f32 / ua  (f32)"""
    b, q = np.asarray(boxes, np.float32), np.asarray(query, np.float32)
    f64 = lambda x: x.astype(np.float64)
    area_q = ((f64(q[:, 2] - q[:, 0]) + 1.0) * (f64(q[:, 3] - q[:, 1]) + 1.0)).astype(np.float32)
    area_b = (f64(b[:, 2] - b[:, 0]) + 1.0) * (f64(b[:, 3] - b[:, 1]) + 1.0)                       # stays double
    iw = (f64(np.minimum(b[:, None, 2], q[None, :, 2]) - np.maximum(b[:, None, 0], q[None, :, 0])) + 1.0).astype(np.float32)
    ih = (f64(np.minimum(b[:, None, 3], q[None, :, 3]) - np.maximum(b[:, None, 1], q[None, :, 1])) + 1.0).astype(np.float32)
    inter = iw * ih                                                                                  # float32 product
    ua = (area_b[:, None] + f64(area_q)[None, :] - f64(inter)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        ov = inter / ua
    return np.where((iw > 0) & (ih > 0), ov, np.float32(0)).astype(np.float32)


def generate_label(gt_box_center, mask, class_ids_padded, anchors, pos_threshold, neg_threshold):
    """voxel_postprocessor.py:217-354. gt_box_center [max_num,7], mask [max_num], class_ids_padded [max_num],
    anchors [H,W,A,7] float64. Returns pos_equal_one / neg_equal_one [H,W,A] f64, targets [H,W,7A] f64, cls_labels int."""
    gt_box_center = np.asarray(gt_box_center)
    mask = np.asarray(mask)
    H, W, A = anchors.shape[:3]
    flat = anchors.reshape(-1, 7)
    anchors_d = np.sqrt(flat[:, 4] ** 2 + flat[:, 5] ** 2)
    pos = np.zeros((H, W, A))
    neg = np.zeros((H, W, A))
    targets = np.zeros((H, W, A * 7))
    cls = np.zeros((H, W, A), dtype=int)
    valid = gt_box_center[mask == 1]
    cls_valid = np.asarray(class_ids_padded)[mask == 1]
    iou = bbox_overlaps(standup_boxes(flat), standup_boxes(valid))               # [N, n]
    n = iou.shape[1]
    # per ground truth: the anchor with the highest IoU (first one on ties), kept if that IoU is positive
    best = np.argmax(iou.T, axis=1) if n else np.zeros((0,), np.int64)
    best_gt = np.arange(n)
    keep = iou.T[best_gt, best] > 0 if n else np.zeros((0,), bool)
    best, best_gt = best[keep], best_gt[keep]
    a_pos, g_pos = np.where(iou > pos_threshold)
    a_neg = np.where(np.sum(iou < neg_threshold, axis=1) == n)[0]
    a_all = np.concatenate([a_pos, best])
    g_all = np.concatenate([g_pos, best_gt])
    a_uni, first = np.unique(a_all, return_index=True)                           # first occurrence decides the match
    g_uni = g_all[first]
    ih, iw, ia = np.unravel_index(a_uni, (H, W, A))
    pos[ih, iw, ia] = 1
    cls[ih, iw, ia] = cls_valid[g_uni]
    # NB the reference indexes the PADDED gt array with indices into the valid subset (:317-338); identical whenever the
    # valid boxes come first, which is how the dataset pads them
    g, a = gt_box_center[g_uni], flat[a_uni]
    ia7 = np.array(ia) * 7
    targets[ih, iw, ia7] = (g[:, 0] - a[:, 0]) / anchors_d[a_uni]
    targets[ih, iw, ia7 + 1] = (g[:, 1] - a[:, 1]) / anchors_d[a_uni]
    targets[ih, iw, ia7 + 2] = (g[:, 2] - a[:, 2]) / a[:, 3]
    targets[ih, iw, ia7 + 3] = np.log(g[:, 3] / a[:, 3])
    targets[ih, iw, ia7 + 4] = np.log(g[:, 4] / a[:, 4])
    targets[ih, iw, ia7 + 5] = np.log(g[:, 5] / a[:, 5])
    targets[ih, iw, ia7 + 6] = g[:, 6] - a[:, 6]
    ih, iw, ia = np.unravel_index(a_neg, (H, W, A))
    neg[ih, iw, ia] = 1
    ih, iw, ia = np.unravel_index(best, (H, W, A))
    neg[ih, iw, ia] = 0
    return {"pos_equal_one": pos, "neg_equal_one": neg, "targets": targets, "cls_labels": cls}


def collate(label_list):
    """collate_batch_airv2x :392-430"""
    return {"targets": torch.from_numpy(np.array([d["targets"] for d in label_list])),
            "pos_equal_one": torch.from_numpy(np.array([d["pos_equal_one"] for d in label_list])),
            "neg_equal_one": torch.from_numpy(np.array([d["neg_equal_one"] for d in label_list])),
            "class_ids": torch.from_numpy(np.array([d["cls_labels"] for d in label_list]))}


def synth_gt(params, seed, n_gt=20, max_num=None):
    """SURVEY §8d planted boxes: (h,w,l) = anchor size x U(0.8,1.2), yaw near 0 or pi/2, class in 1..6, padded to
    max_num with a prefix mask. A few boxes are placed to overlap each other / sit exactly on anchor centres so ties
    (equal IoU for two anchors, an anchor above the positive threshold for two boxes) occur."""
    g = np.random.default_rng(seed)
    aa = params["anchor_args"]
    rng = aa["cav_lidar_range"]
    max_num = max_num or params["max_num"]
    box = np.zeros((max_num, 7), np.float32)
    mask = np.zeros((max_num,), np.int64)
    cls = np.zeros((max_num,), np.int64)
    for k in range(n_gt):
        s = g.uniform(0.8, 1.2)
        yaw = g.uniform(-0.3, 0.3) + (np.pi / 2 if g.uniform() > 0.5 else 0.0)
        x, y = g.uniform(rng[0] + 6, rng[3] - 6), g.uniform(rng[1] + 5, rng[4] - 5)
        if k % 4 == 1:                       # next to the previous box: one anchor can exceed the threshold for both
            x, y = box[k - 1, 0] + g.uniform(-0.6, 0.6), box[k - 1, 1] + g.uniform(-0.4, 0.4)
            yaw = box[k - 1, 6]
        if k % 4 == 2:                       # snapped between anchor centres (0.8 m grid): equal IoU for neighbours
            x, y, yaw, s = round(x / 0.8) * 0.8, round(y / 0.8) * 0.8, 0.0, 1.0
        if k % 7 == 6:                       # tiny far-off box: only the "highest IoU anchor" rule can match it
            s = 0.25
        box[k] = [x, y, -1.0 + g.uniform(-0.1, 0.1), aa["h"] * s, aa["w"] * s, aa["l"] * s, yaw]
        mask[k] = 1
        cls[k] = g.integers(1, 7)
    return box, mask, cls
