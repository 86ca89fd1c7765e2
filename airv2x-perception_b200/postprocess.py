"""Host side of detection post-processing and AP evaluation on the B200 kernels (csrc/postprocess.cu).

Mirrors VoxelPostprocessor.{generate_anchor_box, post_process_airv2x}
(opencood/data_utils/post_processor/voxel_postprocessor.py:33-86, :666-840) and caluclate_tp_fp / calculate_ap / voc_ap
(opencood/utils/eval_utils_opv2v.py:15-152). The decode, filters, rotated IoU and the greedy NMS run on the GPU; only the
per-frame greedy TP / FP matching over the small detections x ground-truth IoU matrix and the AP integration stay on the
host (they are O(100) scalar steps).
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import _ptr, call, stream_ptr


def generate_anchor_box(anchor_args, order="hwl"):
    """[H/stride, W/stride, A, 7] anchors (x, y, z, h, w, l, yaw) — voxel_postprocessor.py:33-86"""
    W, H = anchor_args["W"], anchor_args["H"]
    r = [math.radians(e) for e in anchor_args["r"]]
    A = len(r)
    rng = anchor_args["cav_lidar_range"]
    stride = anchor_args.get("feature_stride", 2)
    x = np.linspace(rng[0] + anchor_args["vw"], rng[3] - anchor_args["vw"], W // stride)
    y = np.linspace(rng[1] + anchor_args["vh"], rng[4] - anchor_args["vh"], H // stride)
    cx, cy = np.meshgrid(x, y)
    cx = np.tile(cx[..., np.newaxis], A)
    cy = np.tile(cy[..., np.newaxis], A)
    cz = np.ones_like(cx) * -1.0
    w = np.ones_like(cx) * anchor_args["w"]
    l = np.ones_like(cx) * anchor_args["l"]
    h = np.ones_like(cx) * anchor_args["h"]
    r_ = np.ones_like(cx)
    for i in range(A):
        r_[..., i] = r[i]
    if order != "hwl":
        raise NotImplementedError("only the PointPillar 'hwl' box order is implemented")
    return np.stack([cx, cy, cz, h, w, l, r_], axis=-1)


class DetPostprocessor:
    """post_process_airv2x for one scene on the GPU. `params` = hypes["postprocess"]."""

    def __init__(self, params, device, max_out=1000):
        self.params = params
        self.device = torch.device(device)
        aa = params["anchor_args"]
        self.anchors = torch.from_numpy(generate_anchor_box(aa, params["order"])).float().to(self.device).contiguous()
        self.H, self.W, self.A = self.anchors.shape[:3]
        self.K = aa.get("num_class", 7)
        self.range = (ctypes.c_float * 6)(*[float(v) for v in aa["cav_lidar_range"]])
        self.obj_thr = float(params["target_args"]["obj_threshold"])
        self.nms_thr = float(params["nms_thresh"])
        lib = _lib.load()
        lib.a2x_postprocess_workspace_bytes.restype = ctypes.c_size_t
        self.ws_bytes = lib.a2x_postprocess_workspace_bytes(ctypes.c_int(self.H * self.W * self.A))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.max_out = max_out
        d = self.device
        self.corners = torch.empty(max_out, 8, 3, device=d)
        self.scores = torch.empty(max_out, device=d)
        self.labels = torch.empty(max_out, dtype=torch.int32, device=d)
        self.boxes = torch.empty(max_out, 7, device=d)
        self.anchor_idx = torch.empty(max_out, dtype=torch.int32, device=d)
        self.counts = torch.zeros(2, dtype=torch.int32, device=d)  # n_out, status

    def __call__(self, output_dict):
        """output_dict: the model's {"psm","rm","obj"} (logical NCHW views of the NHWC head tensor, batch 1).
        Returns (pred_box3d [n,8,3], scores [n], labels [n], boxes3d [n,7]) or (None,)*4 like the reference."""
        psm, rm, obj = output_dict["psm"], output_dict["rm"], output_dict["obj"]
        assert psm.shape[0] == 1, "inference only has 1 batch"
        heads = torch.cat([psm, rm, obj], 1).permute(0, 2, 3, 1).contiguous()  # NHWC (a view chain if already NHWC)
        assert heads.shape[1] == self.H and heads.shape[2] == self.W, "anchor grid and head grid differ"
        call("a2x_postprocess_det", _ptr(heads), ctypes.c_int(heads.shape[3]), ctypes.c_int(self.H), ctypes.c_int(self.W),
             ctypes.c_int(self.A), ctypes.c_int(self.K), _ptr(self.anchors), ctypes.c_float(self.obj_thr),
             ctypes.c_float(self.nms_thr), self.range, _ptr(self.ws), ctypes.c_size_t(self.ws_bytes), _ptr(self.corners),
             _ptr(self.scores), _ptr(self.labels), _ptr(self.boxes), _ptr(self.anchor_idx), ctypes.c_int(self.max_out),
             _ptr(self.counts[0:1]), _ptr(self.counts[1:2]), stream_ptr())
        n, status = [int(v) for v in self.counts.tolist()]  # the one D2H read of the post-processing
        if status:
            raise RuntimeError("postprocess capacity overflow (status %d)" % status)
        if n == 0:
            return None, None, None, None
        return self.corners[:n], self.scores[:n], self.labels[:n].long(), self.boxes[:n]


def rotated_iou_matrix(boxes_a, boxes_b):
    """[na,8,3] x [nb,8,3] device tensors -> [na,nb] IoU of the xy polygons of the first four corners"""
    a, b = boxes_a.contiguous().float(), boxes_b.contiguous().float()
    out = torch.empty(a.shape[0], b.shape[0], device=a.device)
    call("a2x_rotated_iou_matrix", _ptr(a), ctypes.c_int(a.shape[0]), _ptr(b), ctypes.c_int(b.shape[0]), _ptr(out),
         stream_ptr())
    return out


def calculate_tp_fp(det_boxes, det_score, gt_boxes, result_stat, iou_thresh):
    """eval_utils_opv2v.py:41-95 (`caluclate_tp_fp`): greedy matching in score order, each GT matched at most once"""
    fp, tp = [], []
    gt = int(gt_boxes.shape[0])
    if det_boxes is not None:
        score = det_score.detach().cpu().numpy()
        order = np.argsort(-score)
        iou = rotated_iou_matrix(det_boxes, gt_boxes).cpu().numpy() if gt > 0 else np.zeros((len(score), 0), np.float32)
        alive = list(range(gt))
        for i in order:
            if not alive or np.max(iou[i, alive]) < iou_thresh:
                fp.append(1)
                tp.append(0)
                continue
            fp.append(0)
            tp.append(1)
            alive.pop(int(np.argmax(iou[i, alive])))
        result_stat[iou_thresh]["score"] += score[order].tolist()
    result_stat[iou_thresh]["fp"] += fp
    result_stat[iou_thresh]["tp"] += tp
    result_stat[iou_thresh]["gt"] += gt


def voc_ap(rec, prec):
    """eval_utils_opv2v.py:15-38"""
    mrec = [0.0] + list(rec) + [1.0]
    mpre = [0.0] + list(prec) + [0.0]
    for i in range(len(mpre) - 2, -1, -1):
        mpre[i] = max(mpre[i], mpre[i + 1])
    ap = 0.0
    for i in range(1, len(mrec)):
        if mrec[i] != mrec[i - 1]:
            ap += (mrec[i] - mrec[i - 1]) * mpre[i]
    return ap, mrec, mpre


def calculate_ap(result_stat, iou, global_sort_detections=False):
    """eval_utils_opv2v.py:98-152"""
    st = result_stat[iou]
    fp, tp = list(st["fp"]), list(st["tp"])
    if global_sort_detections:
        idx = np.argsort(-np.array(st["score"]))
        fp, tp = [fp[i] for i in idx], [tp[i] for i in idx]
    gt_total = st["gt"]
    fp, tp = np.cumsum(fp).tolist(), np.cumsum(tp).tolist()
    rec = [float(t) / gt_total for t in tp]
    prec = [float(t) / (f + t) for f, t in zip(fp, tp)]
    return voc_ap(rec, prec)
