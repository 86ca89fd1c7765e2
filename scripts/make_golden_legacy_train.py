"""Train-mode golden of the REAL reference `point_pillar_where2comm` (forward with batch-statistic BatchNorm and the
top-K communication mask, PointPillarLoss, autograd backward) against the oracle — groundwork for the legacy models'
training step on the kernels. Writes tests/golden/ppw2c_train_small.npz (sampled logits, loss, sampled gradients).

    python scripts/make_golden_legacy_train.py
"""
import json
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_import, w2c_oracle as O  # noqa: E402

K_SEED, LABEL_SEED = 17, 5


def labels(H, W, A):
    g = torch.Generator().manual_seed(LABEL_SEED)
    pos = torch.zeros(1, H, W, A, dtype=torch.float64)
    pos.view(-1)[torch.randperm(H * W * A, generator=g)[:20]] = 1.0
    tg = 0.3 * torch.randn(1, H, W, 7 * A, generator=g, dtype=torch.float64) * pos.repeat_interleave(7, -1)
    return {"pos_equal_one": pos, "targets": tg}


def main():
    import test_ppw2c_cpu as T

    torch.set_num_threads(8)
    os.makedirs("/tmp/a2x_golden/debug", exist_ok=True)
    os.chdir("/tmp/a2x_golden")
    ref_import.install()
    from opencood.loss.point_pillar_loss import PointPillarLoss
    from opencood.models.point_pillar_where2comm import PointPillarWhere2comm

    cfg, gold = T.load()
    args = json.loads(json.dumps(cfg["model_args"]))
    model = PointPillarWhere2comm(args)
    sd = T.golden_state_dict(model, gold)
    model.load_state_dict(sd)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    dd = T.golden_scene(cfg, gold)
    model.train()
    random.seed(K_SEED)
    out = model({k: (v.clone() if torch.is_tensor(v) else v) for k, v in dd.items()})
    A = args["anchor_number"]
    lab = labels(out["psm"].shape[2], out["psm"].shape[3], A)
    loss = PointPillarLoss({"cls_weight": 1.0, "reg": 2.0})(out, lab)
    loss.backward()
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "gaussian" not in k else v.clone())
         for k, v in sd.items()}
    random.seed(K_SEED)
    ora, bufs = O.pp_where2comm_forward(p, args, dd, training=True)
    oloss = O.point_pillar_loss(ora, lab, 1.0, 2.0)[0]
    oloss.backward()
    res = {"k_seed": K_SEED, "label_seed": LABEL_SEED, "loss": float(loss)}
    for k in ("psm", "rm"):
        err = float((out[k] - ora[k]).abs().max())
        print("train %s: ref-vs-oracle %.3e" % (k, err))
        assert err < 1e-5
        res["train_" + k] = out[k].detach().numpy()
    print("loss %.8f (ref) %.8f (oracle); com %.6f / %.6f" % (float(loss), float(oloss), float(out["com"]), float(ora["com"])))
    assert abs(float(loss) - float(oloss)) < 1e-6 * abs(float(loss)) and abs(float(out["com"]) - float(ora["com"])) < 1e-7
    res["train_com"] = float(out["com"])
    worst, n = 0.0, 0
    for name, q in model.named_parameters():
        if q.grad is None:
            assert p[name].grad is None or float(p[name].grad.abs().max()) == 0.0, name
            continue
        e = float((q.grad - p[name].grad).abs().max() / (q.grad.abs().max() + 1e-30))
        worst, n = max(worst, e), n + 1
        res["grad_" + name] = q.grad.flatten()[:: max(1, q.grad.numel() // 256)][:256].numpy()
    print("gradients: %d tensors, worst relative max-abs difference %.3e" % (n, worst))
    assert worst < 1e-4
    for k, v in bufs.items():
        assert float((model.state_dict()[k] - v).abs().max()) < 1e-5, k
    dst = os.path.join(ROOT, "tests", "golden", "ppw2c_train_small.npz")
    np.savez_compressed(dst, **res)
    print("wrote", dst, "%.1f KB" % (os.path.getsize(dst) / 1024))


if __name__ == "__main__":
    main()
