"""GPU (-m gpu): BASELINE config 1 — the drop-in PointPillarWhere2comm (stride-2 shrink header, bilinearly resized
communication mask) against the golden vectors recorded from the REAL reference. Tolerance: logits max-abs <= 1e-3."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import test_ppw2c_cpu as T
import w2c_common as C

pytestmark = pytest.mark.gpu


def test_resize_bilinear():
    import a2x_import

    ops = a2x_import.pkg("ops")
    g = torch.Generator().manual_seed(0)
    for (h, w, H, W) in ((32, 32, 64, 64), (25, 88, 50, 176), (7, 5, 21, 10)):
        src = torch.rand(3, h, w, generator=g).cuda()
        dst = torch.empty(3, H, W, device="cuda")
        ops.resize_bilinear(src, dst)
        want = F.interpolate(src[:, None], size=(H, W), mode="bilinear", align_corners=False)[:, 0]
        assert float((dst - want).abs().max()) < 1e-6


def test_eval_matches_reference_golden():
    import a2x_import

    M = a2x_import.pkg("opencood.models.point_pillar_where2comm")
    cfg, gold = T.load()
    model = M.PointPillarWhere2comm(cfg["model_args"])
    model.load_state_dict(T.golden_state_dict(model, gold))
    model.cuda().eval()
    with torch.no_grad():
        out = model(C.to_device(T.golden_scene(cfg, gold), "cuda"))
    for k in ("psm", "rm"):
        assert out[k].shape == gold["eval_" + k].shape
        assert np.abs(out[k].cpu().numpy() - gold["eval_" + k]).max() < 1e-3, k
    assert abs(float(out["com"]) - float(gold["eval_com"])) < 1e-6
    assert out["comm_rate"] == int(gold["eval_comm_rate"]) and out["mask"] == 0
