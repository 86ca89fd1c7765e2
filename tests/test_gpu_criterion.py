"""GPU (-m gpu): the criterion objects `train_utils.create_loss` instantiates (`opencood.loss.point_pillar_loss_multiclass`,
`opencood.loss.point_pillar_loss` of this repo: one fused value + gradient kernel behind `criterion(output_dict, label_dict)`
/ `loss.backward()` / `criterion.logging`) against the oracle restatements of the reference's loss modules with torch
autograd on the CPU. Tolerances: value 1e-4 relative, gradients 1e-6 + 1e-4 relative (fp32 kernels vs fp32 autograd)."""
import json
import os

import pytest
import torch

from oracle import labels_oracle as LO, postprocess_oracle as PO, w2c_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def setup():
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "w2c_small_config.json")))
    pp = cfg["postprocess"]
    anchors = PO.generate_anchor_box(pp["anchor_args"], pp["order"])
    labs = [LO.generate_label(*LO.synth_gt(pp, s, n_gt=n), anchors, pp["target_args"]["pos_threshold"],
                              pp["target_args"]["neg_threshold"]) for s, n in ((301, 9), (302, 0), (303, 25))]
    lab = LO.collate(labs)
    assert float(lab["pos_equal_one"].sum()) > 0
    return anchors.shape[:3], lab


def _oracle(fn, outs, lab, *args):
    leaves = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in outs.items()}
    res = fn(leaves, lab, *args)
    res[0].backward()
    return [float(r) for r in res], {k: v.grad for k, v in leaves.items()}


def _check(total, grads, ref_vals, ref_grads):
    assert abs(float(total) - ref_vals[0]) <= 1e-4 * abs(ref_vals[0])
    for k, g in grads.items():
        r = ref_grads[k]
        assert g.shape == r.shape
        err = (g.detach().cpu().double() - r.double()).abs()
        assert bool((err <= 1e-6 + 1e-4 * r.double().abs()).all()), (k, float(err.max()))


@pytest.mark.parametrize("fused_layout", [True, False])
def test_multiclass_criterion_matches_the_oracle_with_autograd(setup, fused_layout):
    import a2x_import

    (H, W, A), lab = setup
    K, B = 7, 3
    crit = a2x_import.pkg("opencood.loss.point_pillar_loss_multiclass").PointPillarLossMultiClass(
        {"cls_weight": 1.0, "reg": 2.0, "num_class": K})
    g = torch.Generator().manual_seed(5)
    nc, nr = A * K, 7 * A
    if fused_layout:        # what the models return: channel slices of one padded NHWC logit tensor
        heads = torch.randn(B, H, W, 64, generator=g).cuda().requires_grad_(True)
        nchw = heads.permute(0, 3, 1, 2)
        outs = {"psm": nchw[:, :nc], "rm": nchw[:, nc:nc + nr], "obj": nchw[:, nc + nr:nc + nr + A]}
    else:                   # any other producer: three separate NCHW tensors
        outs = {"psm": torch.randn(B, nc, H, W, generator=g).cuda().requires_grad_(True),
                "rm": torch.randn(B, nr, H, W, generator=g).cuda().requires_grad_(True),
                "obj": torch.randn(B, A, H, W, generator=g).cuda().requires_grad_(True)}
    label_dict = {k: v.cuda() for k, v in lab.items()}
    total = crit(outs, label_dict)
    assert total.dtype == torch.float64 and total.dim() == 0
    (2.0 * total).backward()
    ref_vals, ref_grads = _oracle(O.point_pillar_loss_multiclass, outs, lab, K, 1.0, 2.0)
    if fused_layout:
        gh = heads.grad.permute(0, 3, 1, 2)
        assert float(heads.grad[..., nc + nr + A:].abs().max()) == 0.0
        grads = {"psm": gh[:, :nc] / 2, "rm": gh[:, nc:nc + nr] / 2, "obj": gh[:, nc + nr:nc + nr + A] / 2}
    else:
        grads = {k: v.grad / 2 for k, v in outs.items()}
    _check(total, grads, ref_vals, ref_grads)
    msg = crit.logging(3, 0, 10)
    assert msg.startswith("[epoch 3][1/10], || Loss: ") and "reg: " in msg and "conf: " in msg and "total: " in msg
    for k, r in (("total_loss", ref_vals[0]), ("reg_loss", ref_vals[1]), ("conf_loss", ref_vals[2])):
        assert abs(crit.loss_dict[k] - r) <= 1e-4 * abs(r), k


def test_legacy_criterion_matches_the_oracle_with_autograd(setup):
    import a2x_import

    (H, W, A), lab = setup
    B = 3
    crit = a2x_import.pkg("opencood.loss.point_pillar_loss").PointPillarLoss({"cls_weight": 1.0, "reg": 2.0})
    g = torch.Generator().manual_seed(6)
    outs = {"psm": torch.randn(B, A, H, W, generator=g).cuda().requires_grad_(True),
            "rm": torch.randn(B, 7 * A, H, W, generator=g).cuda().requires_grad_(True)}
    total = crit(outs, {k: v.cuda() for k, v in lab.items()})
    total.backward()
    ref_vals, ref_grads = _oracle(O.point_pillar_loss, outs, lab, 1.0, 2.0)
    _check(total, {k: v.grad for k, v in outs.items()}, ref_vals, ref_grads)
    assert "Loss:" in crit.logging(0, 0, 1)


def test_reference_loop_with_the_criterion_equals_the_fused_step():
    """tools/train.py:216-226 as written — `model(batch)` -> `criterion(output, label_dict)` -> `loss.backward()` — with both
    objects from this repo (torch.library forward/backward ops + the fused criterion) lands on the same loss and the same
    parameter gradients as the one-call `model.train_step` (same kernels, same order of the top-K draws)."""
    import random

    import a2x_import
    import w2c_common as C

    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    cfg, gold = C.load_small()
    model = M.Airv2xWhere2com(cfg["model_args"])
    sd = C.golden_state_dict(model, gold)
    model.load_state_dict(sd)
    model.cuda().train()
    dd = C.to_device(C.golden_scene(cfg, gold), "cuda")
    H, W = gold["train_psm"].shape[2:]
    labels = O.make_labels(int(gold["label_seed"]), 1, H, W, cfg["model_args"]["anchor_number"])
    la = cfg["loss_args"]
    crit = a2x_import.pkg("opencood.loss.point_pillar_loss_multiclass").PointPillarLossMultiClass(
        {"cls_weight": la["cls_weight"], "reg": la["reg"], "num_class": cfg["model_args"]["num_class"]})
    random.seed(7)
    out = model(dd)
    loss = crit(out, labels)
    model.zero_grad()
    loss.backward()
    g_auto = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    assert len(g_auto) > 50
    model.load_state_dict(sd)
    random.seed(7)
    loss3 = model.train_step(dd, labels, la["cls_weight"], la["reg"])
    assert abs(float(loss) - float(loss3.sum())) <= 1e-6 * abs(float(loss3.sum()))
    for n, p in model.named_parameters():
        if n in g_auto:
            scale = float(p.grad.abs().max())
            assert float((g_auto[n] - p.grad).abs().max()) <= 1e-4 * scale + 1e-7, n
