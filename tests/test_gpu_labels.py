"""GPU (-m gpu): anchor-target assignment on the CUDA kernel (labels.TargetAssigner -> a2x_assign_targets) == the oracle
(pinned to the real reference): positive / negative / class maps identical, regression targets = the reference's float64
targets rounded to fp32 (double log on the device: within 1 ulp of fp32)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import labels_oracle as LO, postprocess_oracle as PO

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def setup():
    import a2x_import

    cfg = json.load(open(os.path.join(ROOT, "configs", "airv2x_intermediate_where2com.json")))
    params = cfg["postprocess"]
    return params, PO.generate_anchor_box(params["anchor_args"], params["order"]), a2x_import.pkg("labels").TargetAssigner(params, "cuda")


def compare(out, b, lab):
    pos, neg = out["pos_equal_one"][b].cpu().numpy(), out["neg_equal_one"][b].cpu().numpy()
    assert np.array_equal(pos, lab["pos_equal_one"].astype(np.float32))
    assert np.array_equal(neg, lab["neg_equal_one"].astype(np.float32))
    assert np.array_equal(out["class_ids"][b].cpu().numpy(), lab["cls_labels"].astype(np.int32))
    ref = lab["targets"].astype(np.float32)
    got = out["targets"][b].cpu().numpy()
    assert np.array_equal(got == 0, ref == 0)
    assert np.all(np.abs(got - ref) <= 1.2e-7 * np.abs(ref))


def test_batch_matches_reference_golden(setup):
    params, anchors, assigner = setup
    gold = np.load(os.path.join(ROOT, "tests", "golden", "labels.npz"))
    seeds = [int(s) for s in gold["seeds"]]
    gts = [LO.synth_gt(params, s) for s in seeds]
    out = assigner(np.stack([g[0] for g in gts]), np.stack([g[1] for g in gts]), np.stack([g[2] for g in gts]))
    assert out["targets"].dtype == torch.float32 and out["class_ids"].dtype == torch.int32
    for b, s in enumerate(seeds):
        lab = LO.generate_label(*gts[b], anchors, params["target_args"]["pos_threshold"], params["target_args"]["neg_threshold"])
        compare(out, b, lab)
        pos = np.flatnonzero(out["pos_equal_one"][b].cpu().numpy().reshape(-1))
        assert np.array_equal(pos, gold["pos_idx_%d" % s])                                   # the recorded reference run
        assert np.array_equal(out["class_ids"][b].cpu().numpy().reshape(-1)[pos], gold["pos_cls_%d" % s])


def test_empty_crowded_and_ragged_samples(setup):
    """no ground truth, 200 boxes (two shared-memory chunks), and a single box in one batch"""
    params, anchors, assigner = setup
    g = np.random.default_rng(9)
    mx = params["max_num"]
    crowded = LO.synth_gt(params, 77, n_gt=200)
    empty = (np.zeros((mx, 7), np.float32), np.zeros(mx, np.int64), np.zeros(mx, np.int64))
    single = LO.synth_gt(params, 78, n_gt=1)
    gts = [empty, crowded, single]
    out = assigner(np.stack([x[0] for x in gts]), np.stack([x[1] for x in gts]), np.stack([x[2] for x in gts]))
    for b, x in enumerate(gts):
        compare(out, b, LO.generate_label(*x, anchors, params["target_args"]["pos_threshold"], params["target_args"]["neg_threshold"]))
    assert float(out["pos_equal_one"][0].sum()) == 0 and float(out["neg_equal_one"][0].sum()) == anchors.shape[0] * anchors.shape[1] * anchors.shape[2]


def test_labels_feed_the_fused_loss(setup):
    """the assigner's output is the label dict the loss kernel reads: loss == the reference's loss on the reference's labels"""
    import a2x_import
    from oracle import w2c_oracle as O

    params, anchors, assigner = setup
    ops = a2x_import.pkg("ops")
    box, mask, cls = LO.synth_gt(params, 201)
    out = assigner(box[None], mask[None], cls[None])
    H, W, A = anchors.shape[:3]
    K = 7
    g = torch.Generator().manual_seed(1)
    cs = 32
    heads = torch.randn(1, H, W, cs, generator=g).cuda()
    lab = LO.collate([LO.generate_label(box, mask, cls, anchors, params["target_args"]["pos_threshold"], params["target_args"]["neg_threshold"])])
    nchw = heads.cpu().permute(0, 3, 1, 2)
    ref = O.point_pillar_loss_multiclass({"psm": nchw[:, :A * K], "rm": nchw[:, A * K:A * K + 7 * A], "obj": nchw[:, A * K + 7 * A:A * K + 8 * A]},
                                         lab, K, 1.0, 2.0)[0]
    loss3 = torch.zeros(3, dtype=torch.float64, device="cuda")
    dheads = torch.zeros_like(heads)
    ops.det_loss(heads, A, K, out["targets"], out["pos_equal_one"], out["class_ids"], 1.0, 2.0,
                 torch.zeros(1, device="cuda"), dheads, loss3)
    assert abs(float(loss3.sum()) - float(ref)) < 1e-4 * abs(float(ref))
