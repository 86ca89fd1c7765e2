"""CPU (-m "not gpu"): the anchor-target-assignment oracle against the golden vectors recorded from the REAL reference
(scripts/make_golden_labels.py: VoxelPostprocessor.generate_label_airv2x + the Cython bbox_overlaps)."""
import json
import os

import numpy as np

from oracle import labels_oracle as LO, postprocess_oracle as PO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    cfg = json.load(open(os.path.join(ROOT, "configs", "airv2x_intermediate_where2com.json")))
    return cfg["postprocess"], np.load(os.path.join(ROOT, "tests", "golden", "labels.npz"))


def test_oracle_reproduces_reference_labels():
    params, gold = load()
    anchors = PO.generate_anchor_box(params["anchor_args"], params["order"])
    for s in gold["seeds"]:
        box, mask, cls = LO.synth_gt(params, int(s))
        lab = LO.generate_label(box, mask, cls, anchors, params["target_args"]["pos_threshold"], params["target_args"]["neg_threshold"])
        pos = np.flatnonzero(lab["pos_equal_one"].reshape(-1))
        assert np.array_equal(pos, gold["pos_idx_%d" % s])
        assert np.array_equal(lab["targets"].reshape(-1, 7)[pos], gold["pos_targets_%d" % s])
        assert np.array_equal(lab["cls_labels"].reshape(-1)[pos], gold["pos_cls_%d" % s])
        assert np.array_equal(np.flatnonzero(lab["neg_equal_one"].reshape(-1) == 0), gold["not_neg_idx_%d" % s])
        assert np.count_nonzero(lab["targets"]) == np.count_nonzero(gold["pos_targets_%d" % s])         # zero elsewhere


def test_assignment_rules_on_a_hand_case():
    """one anchor above the positive threshold for two boxes takes the lower box index; a box no anchor reaches 0.6 for
    still claims its best anchor, which then is neither negative; no ground truth -> everything negative"""
    params, _ = load()
    anchors = PO.generate_anchor_box(params["anchor_args"], params["order"])
    H, W, A = anchors.shape[:3]
    a0 = anchors[50, 100, 0]
    box = np.zeros((8, 7), np.float32)
    box[0] = a0
    box[1] = a0
    box[1, 0] += 0.05
    box[2] = [a0[0] + 40.3, a0[1] + 10.1, -1, 0.4, 0.4, 1.0, 0.3]          # tiny: best IoU << 0.45
    mask = np.array([1, 1, 1, 0, 0, 0, 0, 0])
    cls = np.array([3, 5, 2, 0, 0, 0, 0, 0])
    lab = LO.generate_label(box, mask, cls, anchors, 0.6, 0.45)
    assert lab["pos_equal_one"][50, 100, 0] == 1 and lab["cls_labels"][50, 100, 0] == 3
    assert np.all(np.abs(lab["targets"][50, 100, :7]) < 1e-6)            # the box is the anchor (in fp32)
    assert (lab["cls_labels"] == 2).sum() == 1
    iy, ix, ia = [int(v[0]) for v in np.nonzero(lab["cls_labels"] == 2)]
    assert lab["pos_equal_one"][iy, ix, ia] == 1 and lab["neg_equal_one"][iy, ix, ia] == 0
    empty = LO.generate_label(box, np.zeros(8, int), cls, anchors, 0.6, 0.45)
    assert empty["pos_equal_one"].sum() == 0 and empty["neg_equal_one"].sum() == H * W * A and np.all(empty["targets"] == 0)


def test_host_footprints_match_the_oracle():
    """labels._standup (the host half of TargetAssigner: fp32 corner arithmetic of boxes_to_corners_3d) == the oracle's
    standup boxes bit for bit, for the anchors and for rotated ground-truth boxes; a CPU device is refused loudly"""
    import pytest
    import torch

    import a2x_import

    L = a2x_import.pkg("labels")
    params, _ = load()
    anchors = PO.generate_anchor_box(params["anchor_args"], params["order"]).reshape(-1, 7)
    assert np.array_equal(L._standup(anchors).numpy().view(np.uint32), LO.standup_boxes(anchors).view(np.uint32))
    box, mask, _ = LO.synth_gt(params, 201)
    valid = box[mask == 1]
    assert np.array_equal(L._standup(valid).numpy().view(np.uint32), LO.standup_boxes(valid).view(np.uint32))
    assert L._standup(valid[:0]).shape == (0, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        L.TargetAssigner(params, "cpu")
