"""Pin oracle/w2c_oracle.py:point_pillar_loss against the REAL reference `PointPillarLoss` (opencood/loss/point_pillar_loss.py)
— value and gradients w.r.t. the head tensors — and write tests/golden/pploss.npz (groundwork for the legacy models'
training step).

    python scripts/make_golden_legacy_loss.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, w2c_oracle as O  # noqa: E402

B, A, H, W, SEED = 2, 2, 32, 32, 91


def inputs():
    g = torch.Generator().manual_seed(SEED)
    psm = torch.randn(B, A, H, W, generator=g)
    rm = 0.3 * torch.randn(B, 7 * A, H, W, generator=g)
    pos = torch.zeros(B, H, W, A, dtype=torch.float64)
    idx = torch.randperm(B * H * W * A, generator=g)[:25]
    pos.view(-1)[idx] = 1.0
    pos[1] = 0.0                                                      # a sample without positives (clamp(min=1) path)
    targets = 0.3 * torch.randn(B, H, W, 7 * A, generator=g, dtype=torch.float64) * pos.repeat_interleave(7, -1)
    return psm, rm, {"pos_equal_one": pos, "targets": targets}


def main():
    ref_import.install()
    from opencood.loss.point_pillar_loss import PointPillarLoss

    psm, rm, lab = inputs()
    crit = PointPillarLoss({"cls_weight": 1.0, "reg": 2.0})
    a, b = psm.clone().requires_grad_(True), rm.clone().requires_grad_(True)
    ref = crit({"psm": a, "rm": b}, lab)
    ref.backward()
    c, d = psm.clone().requires_grad_(True), rm.clone().requires_grad_(True)
    tot, reg, conf = O.point_pillar_loss({"psm": c, "rm": d}, lab, 1.0, 2.0)
    tot.backward()
    print("total %.9f (ref) %.9f (oracle); reg %.6f conf %.6f" % (float(ref), float(tot), float(reg), float(conf)))
    assert float(ref) == float(tot)
    assert float(crit.loss_dict["reg_loss"]) == float(reg) and float(crit.loss_dict["conf_loss"]) == float(conf)
    assert torch.equal(a.grad, c.grad) and torch.equal(b.grad, d.grad)
    dst = os.path.join(ROOT, "tests", "golden", "pploss.npz")
    np.savez_compressed(dst, seed=SEED, shape=np.array([B, A, H, W]), total=float(ref), reg=float(reg), conf=float(conf),
                        dpsm=a.grad.numpy(), drm_sample=b.grad[:, :, ::4, ::4].numpy())
    print("oracle == reference (value and gradients); wrote", dst)


if __name__ == "__main__":
    main()
