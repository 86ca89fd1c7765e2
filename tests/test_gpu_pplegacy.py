"""GPU (-m gpu): the drop-in legacy transformer-fusion models PointPillarCoBEVT / PointPillarV2XVit (3x3 stride-2 shrink
header, one PillarVFE, ego-warp by the pairwise pose for V2X-ViT) against the golden vectors recorded from the REAL
reference. Tolerance: logits max-abs <= 1e-3."""
import numpy as np
import pytest
import torch

import test_pplegacy_cpu as T
import w2c_common as C

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(T.CASES))
def test_eval_matches_reference_golden(name):
    model, cfg, gold = T.build(name)
    model.load_state_dict(T.golden_state_dict(model, gold))
    model.cuda().eval()
    with torch.no_grad():
        out = model(C.to_device(T.golden_scene(cfg, gold), "cuda"))
        again = model(C.to_device(T.golden_scene(cfg, gold), "cuda"))
    for k in ("psm", "rm"):
        assert out[k].shape == gold["eval_" + k].shape
        err = np.abs(out[k].cpu().numpy() - gold["eval_" + k]).max()
        print(name, k, "max-abs error %.2e" % err)
        assert err < 1e-3, (k, err)
        assert torch.equal(out[k], again[k])
    assert out["comm_rate"] == int(gold["eval_comm_rate"]) and out["mask"] == 0 and out["each_mask"] == 0
    model.train()
    with pytest.raises(NotImplementedError):
        model(C.to_device(T.golden_scene(cfg, gold), "cuda"))
