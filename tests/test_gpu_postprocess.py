"""GPU (-m gpu): detection decode + rotated NMS + AP on the B200 kernels against the oracle / the golden vectors made
with the reference's own loops: identical kept anchors, identical TP / FP lists, identical AP@{0.3, 0.5, 0.7}."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import postprocess_oracle as PO

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def pp():
    import a2x_import

    return a2x_import.pkg("postprocess")


def test_rotated_iou_matrix(pp):
    g = torch.Generator().manual_seed(0)
    b = torch.rand(40, 7, generator=g)
    b[:, 0:2] = b[:, 0:2] * 10
    b[:, 3:6] = 1.0 + 3 * b[:, 3:6]
    b[:, 6] = (b[:, 6] - 0.5) * 6
    c = PO.boxes_to_corners_3d(b)
    got = pp.rotated_iou_matrix(c[:25].cuda(), c[25:].cuda()).cpu().numpy()
    want = np.array([[PO.quad_iou(x[:4, :2].numpy(), y[:4, :2].numpy()) for y in c[25:]] for x in c[:25]])
    assert np.abs(got - want).max() < 1e-6


def test_postprocess_and_ap_match_reference_golden(pp):
    cfg = json.load(open(os.path.join(GOLD, "w2c_small_config.json")))
    gold = np.load(os.path.join(GOLD, "postprocess.npz"))
    params = cfg["postprocess"]
    post = pp.DetPostprocessor(params, "cuda")
    stat = {t: {"tp": [], "fp": [], "gt": 0, "score": []} for t in (0.3, 0.5, 0.7)}
    for frame, seed in enumerate(gold["seeds"].tolist()):
        out, gt = PO.synth_frame(params, seed)
        c, s, l, b = post({k: v.cuda() for k, v in out.items()})
        assert np.array_equal(post.anchor_idx[:c.shape[0]].cpu().numpy(), gold["frame%d_anchor_idx" % frame])   # same boxes, same order
        assert np.abs(c.cpu().numpy() - gold["frame%d_corners" % frame]).max() < 1e-4
        assert np.abs(s.cpu().numpy() - gold["frame%d_scores" % frame]).max() < 1e-6
        assert np.array_equal(l.cpu().numpy(), gold["frame%d_labels" % frame])
        for t in stat:
            pp.calculate_tp_fp(c, s, gt.cuda(), stat, t)
    for t in stat:
        assert stat[t]["tp"] == gold["tp_%d" % int(t * 10)].tolist()
        assert abs(pp.calculate_ap(stat, t)[0] - float(gold["ap_%d" % int(t * 10)])) < 1e-12       # identical AP


def test_postprocess_many_candidates_top1000(pp):
    """more than 1000 candidates above the objectness gate: the top-1000 selection + sort must equal the oracle"""
    cfg = json.load(open(os.path.join(GOLD, "w2c_small_config.json")))
    params = cfg["postprocess"]
    out, _ = PO.synth_frame(params, 7)
    g = torch.Generator().manual_seed(5)
    out["obj"] = out["obj"] + 4.0 + torch.rand(out["obj"].shape, generator=g)        # ~ everything passes the gate
    post = pp.DetPostprocessor(params, "cuda")
    c, s, l, b = post({k: v.cuda() for k, v in out.items()})
    oc, os_, ol, ob, oi = PO.post_process(out, params)
    assert np.array_equal(post.anchor_idx[:c.shape[0]].cpu().numpy(), oi.numpy())
    assert np.abs(c.cpu().numpy() - oc.numpy()).max() < 1e-4
    # nothing passes: the reference returns Nones
    out["obj"] = out["obj"] - 100.0
    assert post({k: v.cuda() for k, v in out.items()}) == (None, None, None, None)
