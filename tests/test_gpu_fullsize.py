"""GPU (-m gpu): parity AT THE SIZES BASELINE.json QUOTES. The CUDA path is compared with
  (a) fixtures recorded from the REAL reference at full size (scripts/make_golden_full.py: strided samples + the last
      rows / columns of every logit map — the cells partial GEMM tiles write), and
  (b) the oracle restatement evaluated on this box's CPU on the same inputs (every cell of every map),
for config 2 (Where2comm, 5 agents x 60 000 points, 200 x 704: eval forward AND one training step with the 32 000-pillar
cap hit), config 3 (V2X-ViT; the reference pads to L = 15, the CUDA path runs the valid agents), config 4 (CoBEVT with 5
and 7 agents at the shipped L = 7 and 8 agents at max_cav 3/3/2) and config 5's 504 x 504 grid (lidar branch), plus the
chain  CUDA logits -> CUDA decode / rotated NMS -> AP  ==  oracle logits -> reference decode / NMS loops -> AP.
Tolerance (north_star): logits max-abs <= 1e-3 against the fp32 reference; kept boxes identical."""
import copy
import json
import os
import random

import numpy as np
import pytest
import torch

import fullsize_common as FC
import w2c_common as C
from oracle import cobevt_oracle as CO, postprocess_oracle as PO, v2xvit_oracle as VO, w2c_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _cfg(name):
    return json.load(open(os.path.join(ROOT, "configs", name)))


def _gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def _full_diff(out, ora):
    worst = 0.0
    for k in ("psm", "rm", "obj"):
        e = float((out[k].detach().float().cpu() - ora[k].detach()).abs().max())
        assert e < TOL, (k, e)
        worst = max(worst, e)
    return worst


def _raw_dict(pre, agents, training):
    """the same clouds FC.scene() voxelises on the CPU, as the raw-point boundary (GPU voxeliser)"""
    rng = pre["cav_lidar_range"]
    clouds = [O.synth_points(FC.SCENE_SEED * 100 + k, FC.N_POINTS, rng, FC.SIGMA_XY) for k in range(len(agents))]
    offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    raw = {"raw_points": {"points": torch.from_numpy(np.concatenate(clouds, 0)), "offsets": torch.from_numpy(offs),
                          "preprocess": pre, "filter": True}}
    for t in O.AGENT_TYPES:
        n = sum(1 for a in agents if a == t)
        raw[t] = {"record_len": [n], "batch_idxs": [0] if n else []}
    return raw


# ---------------------------------------------------------------------------------------------------- config 2 / 5
@pytest.fixture(scope="module")
def w2c_full():
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    cfg = _cfg("airv2x_intermediate_where2com.json")
    model = M.Airv2xWhere2com(cfg["model_args"])
    sd = FC.seeded_state_dict(model, FC.W2C_PARAM_SEED, cls_shift=FC.W2C_CLS_SHIFT)
    model.load_state_dict(sd)
    model.cuda()
    return cfg, model, sd, {}


def _w2c_eval(cfg, model, sd, gold, cache):
    """Eval forward at full size against the recorded reference numbers and, cell by cell, the oracle. The eval-mode
    communication mask is a hard threshold on the smoothed confidence (where2comm_fuse.py:122-123): a pixel whose
    confidence sits within fp32 noise of the threshold may fall on the other side, which switches that cell's features
    on / off for the fusion (an O(0.1) local change of the logits, in the reference's own CPU-vs-GPU runs as well). The
    test therefore (i) requires the CUDA mask to differ from the oracle's at no more than 5 pixels, all within 1e-4 of
    the threshold, and (ii) when a pixel did flip, compares against the oracle with that decision teacher-forced."""
    dd = FC.scene(cfg["preprocess"], training=False)
    model.load_state_dict(sd)
    model.eval()
    keep = {}
    with torch.no_grad():
        out = model(C.to_device(dd, "cuda"))
        mask_gpu = C.engine_buf(model, "mask").cpu().clone()
        raw = model(_raw_dict(cfg["preprocess"], FC.AGENTS, False))
        ora, _ = O.where2com_forward(sd, cfg["model_args"], dd, training=False, keep=keep)
    own = keep["mask"].reshape(mask_gpu.shape)
    diff = own != mask_gpu
    flips = int(diff.sum())
    thr = float(cfg["model_args"]["where2com_fusion"]["communication"]["threshold"])
    if flips:
        assert flips <= 5, "communication mask differs at %d pixels" % flips
        assert float((keep["smooth"].reshape(mask_gpu.shape)[diff] - thr).abs().max()) < 1e-4
        with torch.no_grad():
            ora, _ = O.where2com_forward(sd, cfg["model_args"], dd, training=False, keep={"mask_override": mask_gpu})
    else:
        FC.compare_with_golden(out, gold, "eval_", TOL)             # the real reference's numbers
    _full_diff(out, ora)                                            # every cell, against the oracle
    for k in ("psm", "rm", "obj"):
        assert torch.equal(out[k], raw[k]), k                       # GPU voxeliser == CPU voxeliser at 60k points
    # comm_rate = count_nonzero(BEV canvas) (airv2x_where2com.py:122): of the 8.2 M positive PillarVFE outputs a handful sit
    # within an fp32 ulp of the ReLU's zero, so the count may differ from the reference's by a few units (1e-6 relative)
    assert ora["comm_rate"] == int(gold["eval_comm_rate"]) and out["comm_rate"] == raw["comm_rate"]
    assert abs(out["comm_rate"] - ora["comm_rate"]) <= max(2, int(1e-6 * ora["comm_rate"])), (out["comm_rate"], ora["comm_rate"])
    # the rate is taken BEFORE the ego rows are forced to one (where2comm_fuse.py:137-143): near-threshold pixels of the
    # ego agent can flip without showing in the mask, so allow a few pixels beyond the visible flips
    assert abs(float(out["com"]) - float(gold["eval_com"])) <= (flips + 4.5) / mask_gpu.numel() + 1e-7
    cache["eval"] = (out, ora, flips)


def _w2c_train(cfg, model, sd, gold):
    dd = FC.scene(cfg["preprocess"], training=True)
    model.load_state_dict(sd)
    model.train()
    H, W = (int(v) for v in gold["train_psm_shape"][2:])
    labels = O.make_labels(FC.LABEL_SEED, 1, H, W, cfg["model_args"]["anchor_number"])
    random.seed(FC.K_SEED)
    loss3 = model.train_step(C.to_device(dd, "cuda"), labels, cfg["loss_args"]["cls_weight"], cfg["loss_args"]["reg"])
    total = float(loss3.sum())
    heads = C.engine_buf(model, "B.heads").permute(0, 3, 1, 2)
    A, K = cfg["model_args"]["anchor_number"], cfg["model_args"]["num_class"]
    out = {"psm": heads[:, :A * K], "rm": heads[:, A * K:A * K + 7 * A], "obj": heads[:, A * K + 7 * A:A * K + 8 * A]}
    mask_gpu = C.engine_buf(model, "mask").cpu().clone()
    ref_out, ref_loss, ref_grads, keep = C.oracle_train_step(sd, cfg, dd, labels, FC.K_SEED, mask_override=mask_gpu)
    flips = C.check_mask_ties(mask_gpu, keep)
    _full_diff(out, ref_out)
    assert abs(total - float(ref_loss)) < 1e-3 * abs(float(ref_loss)), (total, float(ref_loss))
    if flips == 0:   # no top-K tie broke differently: the recorded reference numbers apply directly
        FC.compare_with_golden(out, gold, "train_", TOL)
        assert abs(total - float(gold["train_loss"])) < 1e-3 * abs(float(gold["train_loss"]))
    for n in ("cls_head.weight", "cls_head.bias", "reg_head.weight", "reg_head.bias", "obj_head.weight", "obj_head.bias"):
        g = dict(model.named_parameters())[n].grad
        ref = ref_grads[n]
        err = float((g.cpu() - ref).norm() / (ref.norm() + 1e-30))
        assert err < 1e-3, (n, err)
        if flips == 0:
            assert np.abs(C.sample(g, 512) - gold["grad_" + n]).max() <= 1e-3 * np.abs(gold["grad_" + n]).max(), n
    errs = []
    for n, p in model.named_parameters():
        if n in ref_grads and p.grad is not None:
            errs.append(float((p.grad.cpu() - ref_grads[n]).norm() / (ref_grads[n].norm() + 1e-30)))
    assert max(errs) < 0.15 and float(np.median(errs)) < 0.05, (max(errs), float(np.median(errs)))
    return flips


def test_config2_eval_logits_full_size(w2c_full):
    cfg, model, sd, cache = w2c_full
    _w2c_eval(cfg, model, sd, _gold("full_w2c.npz"), cache)


def test_config2_train_step_full_size(w2c_full):
    cfg, model, sd, _ = w2c_full
    gold = _gold("full_w2c.npz")
    assert gold["train_pillars"].tolist() == [64000, 64000, 32000]      # the 32 000-pillar cap is hit by every agent
    _w2c_train(cfg, model, sd, gold)


def test_config2_chain_logits_to_nms_to_ap(w2c_full):
    """CUDA logits -> CUDA decode + rotated NMS -> TP / FP -> AP@{0.3, 0.5, 0.7}  ==  oracle logits -> the reference's
    decode / NMS / matching loops (oracle restatement) -> AP. The objectness bias is shifted so that ~400 anchors pass
    the 0.20 gate (random-init weights pass nearly all 70 400 otherwise); ground truth = jittered copies of every
    second box the reference keeps, so that TPs and FPs both occur."""
    import a2x_import

    cfg, model, sd, cache = w2c_full
    pp = a2x_import.pkg("postprocess")
    if "eval" not in cache:
        _w2c_eval(cfg, model, sd, _gold("full_w2c.npz"), cache)
    _, ora, _ = cache["eval"]
    params = cfg["postprocess"]
    thr = float(params["target_args"]["obj_threshold"])
    obj_sorted = torch.sort(ora["obj"].reshape(-1), descending=True)[0]
    shift = float(np.log(thr / (1 - thr)) - 0.5 * (obj_sorted[399] + obj_sorted[400]))   # gate between ranks 400 / 401
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2["obj_head.bias"] = sd2["obj_head.bias"] + shift
    model.load_state_dict(sd2)
    model.eval()
    dd = FC.scene(cfg["preprocess"], training=False)
    with torch.no_grad():
        out = model(C.to_device(dd, "cuda"))
    ora2 = {"psm": ora["psm"], "rm": ora["rm"], "obj": ora["obj"] + sd2["obj_head.bias"].view(1, -1, 1, 1)
            - sd["obj_head.bias"].view(1, -1, 1, 1)}
    oc, osc, ol, ob, oi = PO.post_process(ora2, params)
    assert oc is not None and 20 < oc.shape[0] <= 400
    post = pp.DetPostprocessor(params, "cuda")
    c, s, l, b = post({k: out[k].contiguous() for k in ("psm", "rm", "obj")})
    gi = post.anchor_idx[:c.shape[0]].cpu().numpy()
    # ---- identical kept boxes. Scores agree to ~1e-5, so the ORDER of two detections can differ only where the
    # reference's own scores are closer than that; the comparison is exact unless such a near-tie exists.
    gaps = np.diff(np.sort(osc.numpy()))
    tie_free = gaps.size == 0 or float(gaps.min()) > 2e-4
    assert sorted(gi.tolist()) == sorted(oi.numpy().tolist()), "different boxes survive the NMS"
    if tie_free:
        assert np.array_equal(gi, oi.numpy())
    order_g, order_o = np.argsort(gi), np.argsort(oi.numpy())
    assert np.abs(c.cpu().numpy()[order_g] - oc.numpy()[order_o]).max() < 1e-3
    assert np.abs(s.cpu().numpy()[order_g] - osc.numpy()[order_o]).max() < 1e-4
    assert np.array_equal(l.cpu().numpy()[order_g], ol.numpy()[order_o])
    # ---- AP chain
    gt_boxes = ob[::2][:40].clone()
    gt_boxes[:, 0] += 0.3
    gt_boxes[:, 1] -= 0.2
    gt = PO.boxes_to_corners_3d(gt_boxes)
    stat_g = {t: {"tp": [], "fp": [], "gt": 0, "score": []} for t in (0.3, 0.5, 0.7)}
    stat_o = copy.deepcopy(stat_g)
    for t in stat_g:
        pp.calculate_tp_fp(c, s, gt.cuda(), stat_g, t)
        PO.tp_fp(oc, osc, gt, stat_o, t)
        ap_g, ap_o = pp.calculate_ap(stat_g, t)[0], PO.calculate_ap(stat_o, t)
        assert sum(stat_g[t]["tp"]) == sum(stat_o[t]["tp"]) and sum(stat_g[t]["fp"]) == sum(stat_o[t]["fp"]), t
        if tie_free:
            assert stat_g[t]["tp"] == stat_o[t]["tp"]
            assert abs(ap_g - ap_o) < 1e-9, (t, ap_g, ap_o)
        else:
            assert abs(ap_g - ap_o) < 2e-2, (t, ap_g, ap_o)
    assert 0.0 < PO.calculate_ap(stat_o, 0.5) < 1.0                   # a non-degenerate AP (both TPs and FPs)
    model.load_state_dict(sd)


def test_config5_grid_504_lidar_branch():
    """the 504 x 504 BEV of config 5 (lidar range +-100.8 m): 252-cell maps against 16 x 8 GEMM tiles, 63-cell deep level"""
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    cfg = json.load(open(os.path.join(GOLD, "full_w2c504_config.json")))
    gold = _gold("full_w2c504.npz")
    model = M.Airv2xWhere2com(cfg["model_args"])
    sd = FC.seeded_state_dict(model, FC.W2C_PARAM_SEED, cls_shift=FC.W2C_CLS_SHIFT)
    model.load_state_dict(sd)
    model.cuda()
    assert tuple(int(v) for v in gold["eval_psm_shape"]) == (1, 14, 252, 252)
    _w2c_eval(cfg, model, sd, gold, {})
    _w2c_train(cfg, model, sd, gold)


# ---------------------------------------------------------------------------------------------------- config 4
@pytest.mark.parametrize("case", [c[0] for c in FC.COBEVT_CASES])
def test_config4_cobevt_eval_full_size(case):
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_cobevt")
    cfg = _cfg("airv2x_intermediate_cobevt.json")
    name, agents, max_cav = [c for c in FC.COBEVT_CASES if c[0] == case][0]
    args = copy.deepcopy(cfg["model_args"])
    if max_cav is not None:
        args["max_cav"] = dict(max_cav)
    gold = _gold("full_cobevt.npz")
    model = M.Airv2xCoBEVT(args)
    sd = FC.seeded_state_dict(model, FC.COBEVT_PARAM_SEED, skip=("relative_position_index",))
    model.load_state_dict(sd)
    model.cuda().eval()
    dd = FC.scene(cfg["preprocess"], training=False, agents=agents)
    with torch.no_grad():
        out = model(C.to_device(dd, "cuda"))
    FC.compare_with_golden(out, gold, name + "_eval_", TOL)
    if case == "a8":     # every cell against the oracle for the largest case (the other two are pinned by the fixtures)
        with torch.no_grad():
            ora, _ = CO.cobevt_forward(sd, args, dd, training=False)
        _full_diff(out, ora)


# ---------------------------------------------------------------------------------------------------- config 3
def test_config3_v2xvit_eval_full_size():
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_v2xvit")
    cfg = _cfg("airv2x_intermediate_v2xvit.json")
    gold = _gold("full_v2xvit.npz")
    args = cfg["model_args"]
    model = M.Airv2xV2XVit(args)
    sd = FC.seeded_state_dict(model, FC.V2XVIT_PARAM_SEED, skip=("rte.emb.emb.weight",))
    model.load_state_dict(sd)
    model.cuda().eval()
    L = int(gold["max_cav_num"])
    dd = FC.scene(cfg["preprocess"], training=False)
    dd["prior_encoding"], dd["spatial_correction_matrix"] = FC.v2xvit_extras(FC.AGENTS, L)
    with torch.no_grad():
        out = model(C.to_device(dd, "cuda"))
    FC.compare_with_golden(out, gold, "eval_", TOL)                    # the reference ran L = 15 padded
    assert abs(out["comm_rate"] - int(gold["eval_comm_rate"])) <= max(2, int(1e-6 * int(gold["eval_comm_rate"])))
    # every cell, against the oracle on the valid agents only (exact: padded agents are masked keys)
    n = len(FC.AGENTS)
    a5 = copy.deepcopy(args)
    a5["max_cav"] = {t: sum(1 for a in FC.AGENTS if a == t) for t in O.AGENT_TYPES}
    dd5 = dict(dd)
    dd5["prior_encoding"], dd5["spatial_correction_matrix"] = dd["prior_encoding"][:, :n], dd["spatial_correction_matrix"][:, :n]
    with torch.no_grad():
        ora, _ = VO.v2xvit_forward({k: v.cpu() for k, v in sd.items()}, a5, dd5, training=False)
    _full_diff(out, ora)
