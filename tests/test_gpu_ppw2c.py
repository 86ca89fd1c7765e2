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


def test_train_step_matches_reference_golden():
    """The legacy model's training step on the kernels (train-mode BatchNorm with the 2 / 1 / 1 running-stat updates of
    point_pillar_where2comm.py:118-146, top-K mask at half resolution resized bilinearly, fused PointPillarLoss, full
    backward) against the run of the REAL reference recorded by scripts/make_golden_legacy_train.py and against the oracle
    with the kernel's top-K tie-breaks teacher-forced (see tests/test_gpu_model.py for why)."""
    import os
    import random
    import sys

    import a2x_import
    from oracle import w2c_oracle as O

    sys.path.insert(0, os.path.join(T.ROOT, "scripts"))
    import make_golden_legacy_train as G

    M = a2x_import.pkg("opencood.models.point_pillar_where2comm")
    cfg, gold = T.load()
    tg = np.load(os.path.join(T.GOLD, "ppw2c_train_small.npz"))
    args = cfg["model_args"]
    model = M.PointPillarWhere2comm(args)
    sd = T.golden_state_dict(model, gold)
    model.load_state_dict(sd)
    model.cuda().train()
    dd = T.golden_scene(cfg, gold)
    H, W = tg["train_psm"].shape[2:]
    lab = G.labels(H, W, args["anchor_number"])
    k_seed = int(tg["k_seed"])
    random.seed(k_seed)
    loss3 = model.train_step(C.to_device(dd, "cuda"), lab, 1.0, 2.0).clone()
    assert float(loss3[2]) == 0.0                                     # no objectness term in PointPillarLoss
    mask_lo = C.engine_buf(model, "mask.lo").cpu().clone()
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "gaussian" not in k else v.clone())
         for k, v in sd.items()}
    keep = {"mask_override": mask_lo}
    torch.set_num_threads(8)
    random.seed(k_seed)
    out, bufs = O.pp_where2comm_forward(p, args, dd, training=True, keep=keep)
    loss = O.point_pillar_loss(out, lab, 1.0, 2.0)[0]
    loss.backward()
    flips = C.check_mask_ties(mask_lo, keep)
    heads = C.engine_buf(model, "B.heads").permute(0, 3, 1, 2).cpu()
    A = args["anchor_number"]
    assert float((heads[:, :A] - out["psm"].detach()).abs().max()) < 1e-3
    assert float((heads[:, A:8 * A] - out["rm"].detach()).abs().max()) < 1e-3
    assert abs(float(loss3.sum()) - float(loss.detach())) < 1e-3 * abs(float(loss.detach()))
    if flips == 0:
        assert np.abs(heads[:, :A].numpy() - tg["train_psm"]).max() < 1e-3
        assert abs(float(loss3.sum()) - float(tg["loss"])) < 1e-3 * float(tg["loss"])
    errs = {}
    for n, q in model.named_parameters():
        if p[n].grad is None:
            continue
        errs[n] = float((q.grad.cpu() - p[n].grad).norm() / (p[n].grad.norm() + 1e-30))
    assert len(errs) == 77
    for n in ("cls_head.weight", "cls_head.bias", "reg_head.weight", "reg_head.bias"):
        assert errs[n] < 1e-3, (n, errs[n])
    assert max(errs.values()) < 0.15 and float(np.median(list(errs.values()))) < 0.05, sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    # running statistics: block 0 twice, everything else once per pass (the oracle applies the reference's update counts)
    for n, b in model.named_buffers():
        if n in bufs and (flips == 0 or ".blocks.0." in n or "pfn_layers" in n):
            assert float((b.cpu() - bufs[n]).abs().max()) < 1e-4, n
    # the reference's own loop: model(batch) -> PointPillarLoss (torch ops) -> loss.backward()
    g_step = {n: q.grad.clone() for n, q in model.named_parameters() if q.grad is not None}
    model.load_state_dict(sd)
    model.zero_grad()
    random.seed(k_seed)
    o2 = model(C.to_device(dd, "cuda"))
    l2 = O.point_pillar_loss({"psm": o2["psm"].cpu(), "rm": o2["rm"].cpu()}, lab, 1.0, 2.0)[0]
    assert abs(float(l2.detach()) - float(loss.detach())) < 1e-3 * abs(float(loss.detach()))
    l2.backward()
    for n, q in model.named_parameters():
        if n in g_step:
            assert float((q.grad - g_step[n]).abs().max()) <= 2e-3 * float(g_step[n].abs().max()) + 1e-7, n


def test_raw_point_boundary_equals_voxel_dict_boundary():
    """raw ego-frame clouds in (GPU voxeliser + the dataset's point filters) == the CPU-voxelised `processed_lidar` dict,
    eval logits and the training step's loss, for the legacy model"""
    import random

    import a2x_import
    from oracle import w2c_oracle as O

    M = a2x_import.pkg("opencood.models.point_pillar_where2comm")
    cfg, gold = T.load()
    model = M.PointPillarWhere2comm(cfg["model_args"])
    sd = T.golden_state_dict(model, gold)
    model.load_state_dict(sd)
    model.cuda().eval()
    pre = cfg["preprocess"]
    n_agents, n_points, seed = int(gold["n_agents"]), int(gold["n_points"]), int(gold["scene_seed"])
    clouds = [O.synth_points(seed * 100 + k, n_points, pre["cav_lidar_range"], sigma_xy=(8.0, 8.0)) for k in range(n_agents)]
    offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    raw = {"raw_points": {"points": torch.from_numpy(np.concatenate(clouds, 0)), "offsets": torch.from_numpy(offs),
                          "preprocess": pre, "filter": True},
           "record_len": torch.tensor([n_agents], dtype=torch.int32)}
    dd = T.golden_scene(cfg, gold)
    with torch.no_grad():
        a = model(raw)
        b = model(C.to_device(dd, "cuda"))
    for k in ("psm", "rm"):
        assert torch.equal(a[k], b[k]), k
    assert a["comm_rate"] == b["comm_rate"]
    # training step: same loss through both boundaries (the train cap of the voxeliser is not hit at this size). The two
    # boundaries list the pillars in different orders (cell order vs first-seen order), so the train-mode BatchNorm batch
    # sums round differently: fp32 summation-order tolerance, not bit equality
    import os
    import sys
    sys.path.insert(0, os.path.join(T.ROOT, "scripts"))
    import make_golden_legacy_train as G
    model.train()
    lab = G.labels(a["psm"].shape[2], a["psm"].shape[3], cfg["model_args"]["anchor_number"])
    random.seed(3)
    l_raw = model.train_step(raw, lab, 1.0, 2.0).clone()
    model.load_state_dict(sd)
    random.seed(3)
    l_dd = model.train_step(C.to_device(dd, "cuda"), lab, 1.0, 2.0).clone()
    assert torch.allclose(l_raw, l_dd, rtol=1e-4)
