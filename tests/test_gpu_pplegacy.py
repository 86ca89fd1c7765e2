"""GPU (-m gpu): the drop-in legacy transformer-fusion models PointPillarCoBEVT / PointPillarV2XVit (3x3 stride-2 shrink
header, one PillarVFE, ego-warp by the pairwise pose for V2X-ViT) against the golden vectors recorded from the REAL
reference (eval logits max-abs <= 1e-3) and their training step against the reference's recorded loss and the oracle's autograd."""
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
    with pytest.raises(NotImplementedError), torch.no_grad():     # train-mode forward exists only with grad enabled
        model(C.to_device(T.golden_scene(cfg, gold), "cuda"))


@pytest.mark.parametrize("name", sorted(T.CASES))
def test_train_step_matches_reference_loss_and_oracle_autograd(name):
    """Training step of the legacy transformer-fusion models (train-mode BatchNorm, dropout off): the loss against the
    value recorded from the REAL reference (model + its PointPillarLoss, scripts/make_golden_legacy_fusion.py), every
    parameter gradient against torch autograd through the oracle (pinned to the reference's gradients to 2e-6 by the same
    script). Fusion / head gradients are tight; encoder gradients pass ReLU / max gates (see tests/test_gpu_model.py) ->
    norm-wise bounds. Then the reference-style loop model(batch) -> loss (torch ops) -> backward gives the same gradients."""
    from oracle import w2c_oracle as O

    model, cfg, gold = T.build(name)
    sd = T.golden_state_dict(model, gold)
    model.load_state_dict(sd)
    model.cuda().train()
    dd = T.golden_scene(cfg, gold)
    loss, out, ref_grads, bufs, lab = T.oracle_train(name, sd, cfg, gold)
    loss3 = model.train_step(C.to_device(dd, "cuda"), lab, 1.0, 2.0, dropout="off").clone()
    assert abs(float(loss3.sum()) - float(gold["train_loss"])) < 1e-3 * abs(float(gold["train_loss"]))
    assert float(loss3[2]) == 0.0
    errs = {}
    for n, q in model.named_parameters():
        ref = ref_grads.get(n)
        if ref is None or float(ref.norm()) < 1e-9:
            continue
        if n.endswith("k_linears.0.bias"):
            # every agent has type 0 here (zero prior encoding): the key bias shifts all scores of a row alike, the softmax
            # is invariant -> a mathematically zero gradient, rounding noise on both sides
            wn = float(ref_grads[n[:-len("bias")] + "weight"].norm())
            assert float(ref.norm()) < 1e-4 * wn and float(q.grad.norm()) < 1e-4 * wn, n
            continue
        errs[n] = float((q.grad.cpu() - ref).norm() / ref.norm())
    fusion = {n: e for n, e in errs.items() if n.startswith("fusion_net") or "head" in n}
    assert len(fusion) > 60 and max(fusion.values()) < 2e-2, sorted(fusion.items(), key=lambda kv: -kv[1])[:5]
    assert float(np.median(list(fusion.values()))) < 2e-3
    assert max(errs.values()) < 0.15 and float(np.median(list(errs.values()))) < 0.05, sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    msd = model.state_dict()
    for k, v in bufs.items():
        if "num_batches" not in k:
            assert float((msd[k].cpu() - v).abs().max()) < 1e-5 + 1e-4 * float(v.abs().max()), k
    # reference-style use (tools/train.py:216-221)
    g_step = {n: q.grad.clone() for n, q in model.named_parameters()}
    model.load_state_dict(sd)
    model.zero_grad()
    model.dropout = "off"
    out2 = model(C.to_device(dd, "cuda"))
    loss2 = O.point_pillar_loss({k: out2[k].cpu() for k in ("psm", "rm")}, lab, 1.0, 2.0)[0]
    assert abs(float(loss2.detach()) - float(gold["train_loss"])) < 1e-3 * abs(float(gold["train_loss"]))
    loss2.backward()
    for n, q in model.named_parameters():
        assert q.grad is not None, n
        assert float((q.grad - g_step[n]).abs().max()) <= 2e-3 * float(g_step[n].abs().max()) + 1e-7, n
