"""Pin oracle/postprocess_oracle.py against the REAL reference functions (imported from /root/reference, CPU) and write
tests/golden/postprocess.npz. shapely is absent in this container, so the reference's polygon IoU
(common_utils.compute_iou / convert_format) is the ONE function replaced (by the oracle's convex-quad clipping) while
the reference's own nms_rotated / caluclate_tp_fp / calculate_ap loops run unmodified around it.

    python scripts/make_golden_postprocess.py
"""
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import postprocess_oracle as PO, ref_import  # noqa: E402


def main():
    ref_import.install()
    sys.modules.setdefault("opencood.utils.box_overlaps", MagicMock())   # Cython label-assignment helper: not on this path
    from opencood.data_utils.post_processor.voxel_postprocessor import VoxelPostprocessor
    from opencood.utils import box_utils, common_utils, eval_utils_opv2v as EU

    common_utils.convert_format = lambda boxes: np.array([b[:4, :2].astype(np.float64) for b in boxes])
    common_utils.compute_iou = lambda box, boxes: PO.compute_iou(box, boxes)
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "w2c_small_config.json")))
    params = cfg["postprocess"]
    out_npz = {}
    stat_ref = {t: {"tp": [], "fp": [], "gt": 0, "score": []} for t in (0.3, 0.5, 0.7)}
    stat_ora = {t: {"tp": [], "fp": [], "gt": 0, "score": []} for t in (0.3, 0.5, 0.7)}
    anchors = torch.from_numpy(PO.generate_anchor_box(params["anchor_args"], params["order"]))
    for frame, seed in enumerate((101, 102, 103)):
        out, gt = PO.synth_frame(params, seed)
        # --- the reference's own pieces
        boxes = VoxelPostprocessor.delta_to_boxes3d(out["rm"], anchors)
        assert float((boxes - PO.delta_to_boxes3d(out["rm"], anchors)).abs().max()) == 0.0
        objectness = torch.sigmoid(out["obj"].permute(0, 2, 3, 1).contiguous()).view(1, -1)
        mask = objectness > params["target_args"]["obj_threshold"]
        idx = torch.nonzero(mask[0]).squeeze(1)
        b3, sc = boxes[0][idx], objectness[0][idx]
        corners = box_utils.boxes_to_corners_3d(b3, order=params["order"])
        assert float((corners - PO.boxes_to_corners_3d(b3)).abs().max()) < 1e-6
        rng = params["anchor_args"]["cav_lidar_range"]
        keep = torch.logical_and(box_utils.remove_large_pred_bbx(corners, "airv2x"),
                                 box_utils.remove_bbx_abnormal_z(corners, z_min=rng[2], z_max=rng[5]))
        corners, sc, idx = corners[keep], sc[keep], idx[keep]
        pick = torch.from_numpy(box_utils.nms_rotated(corners, sc, params["nms_thresh"]).astype(np.int64))
        corners, sc, idx = corners[pick], sc[pick], idx[pick]
        m = box_utils.get_mask_for_boxes_within_range_torch(corners, rng)
        corners, sc, idx = corners[m], sc[m], idx[m]
        # --- the oracle end to end
        oc, os_, ol, ob, oi = PO.post_process(out, params)
        assert torch.equal(oi, idx), (oi, idx)
        assert float((oc - corners).abs().max()) < 1e-6 and float((os_ - sc).abs().max()) == 0.0
        print("frame %d: %d candidates -> %d detections, %d gt" % (frame, int(mask.sum()), len(idx), gt.shape[0]))
        for t in (0.3, 0.5, 0.7):
            EU.caluclate_tp_fp(corners, sc, gt, stat_ref, t)
            PO.tp_fp(oc, os_, gt, stat_ora, t)
        out_npz["frame%d_anchor_idx" % frame] = idx.numpy()
        out_npz["frame%d_scores" % frame] = sc.numpy()
        out_npz["frame%d_corners" % frame] = corners.numpy()
        out_npz["frame%d_labels" % frame] = ol.numpy()
    for t in (0.3, 0.5, 0.7):
        assert stat_ref[t]["tp"] == stat_ora[t]["tp"] and stat_ref[t]["fp"] == stat_ora[t]["fp"]
        ap_ref, _, _ = EU.calculate_ap({k: {kk: list(vv) if isinstance(vv, list) else vv for kk, vv in v.items()}
                                        for k, v in stat_ref.items()}, t, False)
        ap_ora = PO.calculate_ap(stat_ora, t)
        print("AP@%.1f reference loops %.6f oracle %.6f  (tp %d fp %d gt %d)" % (t, ap_ref, ap_ora, sum(stat_ora[t]["tp"]),
                                                                                sum(stat_ora[t]["fp"]), stat_ora[t]["gt"]))
        assert abs(ap_ref - ap_ora) < 1e-12
        out_npz["ap_%d" % int(t * 10)] = ap_ref
        out_npz["tp_%d" % int(t * 10)] = np.array(stat_ora[t]["tp"])
    out_npz["seeds"] = np.array([101, 102, 103])
    dst = os.path.join(ROOT, "tests", "golden", "postprocess.npz")
    np.savez_compressed(dst, **out_npz)
    print("wrote", dst)


if __name__ == "__main__":
    main()
