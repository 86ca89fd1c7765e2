"""CPU (-m "not gpu"): the post-processing oracle against closed-form polygon IoUs and the golden vectors produced with
the REAL reference's decode / filters / nms_rotated / caluclate_tp_fp / calculate_ap loops (scripts/make_golden_postprocess.py)."""
import json
import math
import os

import numpy as np
import torch

from oracle import postprocess_oracle as PO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_quad_iou_closed_forms():
    sq = np.array([[0, 0], [2, 0], [2, 2], [0, 2]], float)
    assert abs(PO.quad_iou(sq, sq) - 1.0) < 1e-12
    assert abs(PO.quad_iou(sq, sq + [1, 0]) - 2.0 / 6.0) < 1e-12              # half overlap: 2 / (4 + 4 - 2)
    assert abs(PO.quad_iou(sq, sq[::-1] + [1, 1]) - 1.0 / 7.0) < 1e-12         # clockwise clip polygon
    assert PO.quad_iou(sq, sq + [5, 0]) == 0.0
    c, s = math.cos(math.pi / 4), math.sin(math.pi / 4)
    dia = (np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], float) @ np.array([[c, s], [-s, c]])) * math.sqrt(2) / 2 * 2 / math.sqrt(2)
    big = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], float)
    # unit-radius diamond (area 2) inside the 2x2 square (area 4): IoU = 2 / 4
    diamond = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]], float)
    assert abs(PO.quad_iou(big, diamond) - 0.5) < 1e-12 and dia.shape == (4, 2)
    # octagon intersection: square vs the same square rotated by 45 degrees -> area 8 (sqrt 2 - 1) r^2 with r = 1
    rot = big @ np.array([[c, s], [-s, c]])
    inter = 8 * (math.sqrt(2) - 1)
    assert abs(PO.quad_iou(big, rot) - inter / (8 - inter)) < 1e-12


def test_oracle_matches_reference_golden():
    cfg = json.load(open(os.path.join(GOLD, "w2c_small_config.json")))
    gold = np.load(os.path.join(GOLD, "postprocess.npz"))
    params = cfg["postprocess"]
    stat = {t: {"tp": [], "fp": [], "gt": 0, "score": []} for t in (0.3, 0.5, 0.7)}
    for frame, seed in enumerate(gold["seeds"].tolist()):
        out, gt = PO.synth_frame(params, seed)
        c, s, l, b, idx = PO.post_process(out, params)
        assert np.array_equal(idx.numpy(), gold["frame%d_anchor_idx" % frame])
        assert np.abs(c.numpy() - gold["frame%d_corners" % frame]).max() < 1e-6
        assert np.array_equal(l.numpy(), gold["frame%d_labels" % frame])
        for t in stat:
            PO.tp_fp(c, s, gt, stat, t)
    for t in stat:
        assert stat[t]["tp"] == gold["tp_%d" % int(t * 10)].tolist()
        assert abs(PO.calculate_ap(stat, t) - float(gold["ap_%d" % int(t * 10)])) < 1e-12
    assert float(gold["ap_3"]) > float(gold["ap_7"])       # the frames discriminate between the IoU thresholds


def test_host_anchor_and_ap_helpers_match_oracle():
    import a2x_import

    P = a2x_import.pkg("postprocess")
    cfg = json.load(open(os.path.join(GOLD, "w2c_small_config.json")))
    aa = cfg["postprocess"]["anchor_args"]
    assert np.array_equal(P.generate_anchor_box(aa), PO.generate_anchor_box(aa))
    rec, prec = [0.1, 0.2, 0.2, 0.5], [1.0, 1.0, 0.66, 0.7]
    assert abs(P.voc_ap(rec, prec)[0] - PO.voc_ap(rec, prec)) < 1e-15
    stat = {0.5: {"tp": [1, 0, 1, 1, 0], "fp": [0, 1, 0, 0, 1], "gt": 4, "score": [0.9, 0.8, 0.7, 0.6, 0.5]}}
    assert abs(P.calculate_ap(stat, 0.5)[0] - PO.calculate_ap(stat, 0.5)) < 1e-15
