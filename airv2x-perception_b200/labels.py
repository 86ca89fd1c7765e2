"""Anchor-target assignment on the GPU: `TargetAssigner(hypes["postprocess"], device)` replaces
`VoxelPostprocessor.generate_label_airv2x` + `collate_batch_airv2x`
(opencood/data_utils/post_processor/voxel_postprocessor.py:217-354, :392-430) for a batch and returns the label dict in
the device layout `model.train_step` / `a2x_det_loss` read (fp32 targets / pos / neg, int32 class ids) — the 70 400 x n
IoU matrix, the thresholds, the per-box best anchor and the target encoding never touch the host.

The footprints of the ground-truth boxes (a few dozen per sample) are prepared on the host with the reference's own
fp32 corner arithmetic (box_utils.py:195-258, common_utils.py:60-82: dims/2 * corner template, rotation by a batched
fp32 matmul, + centre; min / max over the corners); the anchors' footprints once per assigner.
"""
import numpy as np
import torch

from . import ops
from .postprocess import generate_anchor_box


def _standup(boxes7):
    """[n,7] (x,y,z,h,w,l,yaw) -> float32 [n,4] (xmin, ymin, xmax, ymax) of the yaw-rotated footprint"""
    b = torch.as_tensor(np.asarray(boxes7)).float()
    if b.shape[0] == 0:
        return torch.zeros(0, 4)
    dims = b[:, [5, 4, 3]]                                     # l, w, h along x, y, z
    sign = torch.tensor([[1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, -1], [1, -1, 1], [1, 1, 1], [-1, 1, 1], [-1, -1, 1]],
                        dtype=torch.float32) / 2
    corners = dims[:, None, :].repeat(1, 8, 1) * sign[None]
    c, s = torch.cos(b[:, 6]), torch.sin(b[:, 6])
    z, o = torch.zeros_like(c), torch.ones_like(c)
    rot = torch.stack((c, s, z, -s, c, z, z, z, o), dim=1).view(-1, 3, 3).float()
    corners = torch.matmul(corners, rot) + b[:, None, 0:3]
    return torch.stack([corners[:, :, 0].min(1).values, corners[:, :, 1].min(1).values,
                        corners[:, :, 0].max(1).values, corners[:, :, 1].max(1).values], 1).contiguous()


class TargetAssigner:
    def __init__(self, params, device):
        if params["order"] != "hwl":
            raise NotImplementedError("only the PointPillar 'hwl' box order is implemented")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("TargetAssigner (B200) needs a CUDA device; there is no CPU path")
        self.pos_threshold = float(params["target_args"]["pos_threshold"])
        self.neg_threshold = float(params["target_args"]["neg_threshold"])
        anchors = generate_anchor_box(params["anchor_args"], params["order"])
        self.H, self.W, self.A = anchors.shape[:3]
        flat = anchors.reshape(-1, 7)
        self.N = flat.shape[0]
        self.anchors = torch.from_numpy(np.ascontiguousarray(flat)).to(self.device)                 # float64
        self.anchor_standup = _standup(flat).to(self.device)
        self._ws = {}

    def __call__(self, gt_box_center, mask, class_ids_padded):
        """gt_box_center [B,max_num,7], mask [B,max_num], class_ids_padded [B,max_num] (numpy / torch, host or device).
        Returns {"targets" [B,H,W,7A] f32, "pos_equal_one" / "neg_equal_one" [B,H,W,A] f32, "class_ids" [B,H,W,A] i32}."""
        gt = torch.as_tensor(np.asarray(gt_box_center) if not torch.is_tensor(gt_box_center) else gt_box_center).cpu()
        mk = torch.as_tensor(np.asarray(mask) if not torch.is_tensor(mask) else mask).cpu()
        cl = torch.as_tensor(np.asarray(class_ids_padded) if not torch.is_tensor(class_ids_padded) else class_ids_padded).cpu()
        B = gt.shape[0]
        offs, stand, boxes, cls = [0], [], [], []
        for b in range(B):
            sel = mk[b] == 1
            n = int(sel.sum())
            offs.append(offs[-1] + n)
            stand.append(_standup(gt[b][sel]))
            # the reference encodes targets from the PADDED array indexed by the valid-subset index (:317-338)
            boxes.append(gt[b][:n].double())
            cls.append(cl[b][sel].to(torch.int32))
        total = offs[-1]
        dev = self.device
        gt_standup = torch.cat(stand).to(dev) if total else None
        gt_boxes = torch.cat(boxes).contiguous().to(dev) if total else None
        gt_cls = torch.cat(cls).contiguous().to(dev) if total else None
        gt_off = torch.tensor(offs, dtype=torch.int32, device=dev)
        key = (B, max(total, 1))
        ws = self._ws.get(key)
        if ws is None:
            ws = (torch.empty(B * self.N, dtype=torch.int32, device=dev), torch.empty(max(total, 1), dtype=torch.int64, device=dev))
            self._ws = {key: ws}
        out = {"targets": torch.empty(B, self.H, self.W, 7 * self.A, device=dev),
               "pos_equal_one": torch.empty(B, self.H, self.W, self.A, device=dev),
               "neg_equal_one": torch.empty(B, self.H, self.W, self.A, device=dev),
               "class_ids": torch.empty(B, self.H, self.W, self.A, dtype=torch.int32, device=dev)}
        ops.assign_targets(self.anchor_standup, self.anchors, gt_standup, gt_boxes, gt_cls, gt_off, B, self.pos_threshold,
                           self.neg_threshold, ws[0], ws[1], out["targets"], out["pos_equal_one"], out["neg_equal_one"],
                           out["class_ids"])
        return out
