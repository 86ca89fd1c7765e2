"""GPU (-m gpu): the dataset side feeds the kernels. A batch collated by
`intermediate_fusion_dataset.IntermediateFusionDatasetAirv2x` (raw sensor-frame clouds + poses + padded boxes) must give the
model what the reference's CPU pipeline gives it: the same logits as the reference-layout voxel dict built from the same
clouds by the restated filters + sequential voxeliser (oracle = checker), and a training step through `Trainer` whose GPU
anchor targets equal the label oracle's on the dataset's boxes. Integer work bit-exact, logits `torch.equal`."""
import copy
import json
import os
import random
import sys

import numpy as np
import pytest
import torch

import w2c_common as C
import dataset_common as DC
from oracle import labels_oracle as LO, postprocess_oracle as PO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import make_golden_dataset as MGD  # noqa: E402

pytestmark = pytest.mark.gpu


def _setup():
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    DS = a2x_import.pkg("intermediate_fusion_dataset")
    cfg, gold = C.load_small()
    hypes = json.load(open(os.path.join(ROOT, "tests", "golden", "dataset_config.json")))
    hypes = dict(hypes, preprocess=cfg["preprocess"], postprocess=cfg["postprocess"])      # the small grid of the fixtures
    model = M.Airv2xWhere2com(cfg["model_args"])
    model.load_state_dict(C.golden_state_dict(model, gold))
    model.cuda()
    kw = dict(cameras=False, n_pts=6000, obj_span=(20.0, 9.0), agent_spread=0.15, pts_sigma=(12.0, 6.0))
    scenes = [DC.synth_scene(DS, seed=51, n_veh=2, n_rsu=1, n_drone=1, **kw),
              DC.synth_scene(DS, seed=52, n_veh=1, n_rsu=0, n_drone=2, far=False, **kw)]
    return DS, cfg, hypes, model, scenes


def _voxel_dict(ours, hypes, train):
    vox = MGD.voxelise_like_the_reference(ours, hypes, train)
    dd = {"record_len": ours["record_len"]}
    for t in ("vehicle", "rsu", "drone"):
        dd[t] = {"batch_merged_lidar_features_torch": (None if vox[t] is None else
                                                       {k: torch.from_numpy(v) for k, v in vox[t].items()}),
                 "record_len": ours[t]["record_len"], "batch_idxs": ours[t]["batch_idxs"]}
    return dd


def test_dataset_batch_gives_the_logits_of_the_reference_voxel_dict():
    DS, cfg, hypes, model, scenes = _setup()
    _, _, batch = MGD.run_ours(DS, hypes, False, scenes, seed=2)
    ours = batch["ego"]
    assert ours["record_len"].tolist() == [4, 3]
    model.eval()
    with torch.no_grad():
        a = model(ours)
        b = model(C.to_device(_voxel_dict(ours, hypes, False), "cuda"))
    for k in ("psm", "rm", "obj"):
        assert a[k].shape[0] == 2 and torch.equal(a[k], b[k]), k
    assert a["comm_rate"] == b["comm_rate"] and a["comm_rate"] > 0


def test_trainer_step_on_a_dataset_batch():
    import a2x_import

    DS, cfg, hypes, model, scenes = _setup()
    TL = a2x_import.pkg("train_loop")
    th = dict(hypes, loss={"args": cfg["loss_args"]},
              optimizer={"core_method": "Adam", "lr": 0.002, "args": {"eps": 1e-10, "weight_decay": 1e-4}},
              lr_scheduler={"core_method": "multistep", "gamma": 0.1, "step_size": [10, 25, 40]})
    _, _, batch = MGD.run_ours(DS, hypes, True, scenes, seed=2)
    ours = batch["ego"]
    assert int(ours["object_bbx_mask"].sum()) >= 6
    tr = TL.Trainer(model, th)
    lab = tr.labels(ours)
    pp = hypes["postprocess"]
    anchors = PO.generate_anchor_box(pp["anchor_args"], pp["order"])
    for b in range(2):
        ref = LO.generate_label(ours["object_bbx_center"][b].numpy(), ours["object_bbx_mask"][b].numpy(),
                                ours["object_class_ids"][b].numpy(), anchors, pp["target_args"]["pos_threshold"],
                                pp["target_args"]["neg_threshold"])
        assert np.array_equal(lab["pos_equal_one"][b].cpu().numpy(), ref["pos_equal_one"])
        assert np.array_equal(lab["neg_equal_one"][b].cpu().numpy(), ref["neg_equal_one"])
        assert np.abs(lab["targets"][b].cpu().numpy() - ref["targets"]).max() < 1e-5
    assert float(lab["pos_equal_one"].sum()) > 0
    random.seed(4)
    losses = [float(tr.step(copy.copy(ours)).sum()) for _ in range(8)]
    assert all(np.isfinite(losses)) and min(losses[4:]) < losses[0]
    for p in model.parameters():
        assert p.grad is None or bool(torch.isfinite(p.grad).all())


def test_evaluator_over_dataset_batches_matches_the_oracle_chain():
    """eval loop (tools/inference_multi_scenario.py:330-432): dataset test collate -> eval forward -> GPU decode + NMS ->
    TP / FP -> AP, against the reference's python loops (oracle) fed with the same logits and the same ground truth.
    The objectness bias is shifted so that ~150 of the 4096 anchors pass the gate (random weights pass nearly all)."""
    import a2x_import

    DS, cfg, hypes, model, scenes = _setup()
    TL = a2x_import.pkg("train_loop")
    pp = hypes["postprocess"]
    np.random.seed(2)
    ds = DS.IntermediateFusionDatasetAirv2x(hypes, False, False, source=scenes)
    batches = [ds.collate_batch_test([ds[i]]) for i in range(len(scenes))]
    model.eval()
    with torch.no_grad():
        obj = model(batches[0]["ego"])["obj"].reshape(-1).sort(descending=True)[0]
        thr = float(pp["target_args"]["obj_threshold"])
        model.obj_head.bias.add_(float(np.log(thr / (1 - thr)) - 0.5 * (obj[149] + obj[150])))
    ev = TL.Evaluator(model, ds)
    stat_o = {t: {"tp": [], "fp": [], "gt": 0, "score": []} for t in (0.3, 0.5, 0.7)}
    n_gt, tie_free = 0, True
    for batch in batches:
        pred_box, score, labels, boxes3d, gt_box = ev.step(batch)
        assert gt_box.shape[0] == int(batch["ego"]["object_bbx_mask"].sum()) > 0 and gt_box.shape[1:] == (8, 3)
        n_gt += gt_box.shape[0]
        with torch.no_grad():
            out = model(batch["ego"])
        oc, osc, ol, ob, oi = PO.post_process({k: out[k].cpu() for k in ("psm", "rm", "obj")}, pp)
        assert oc is not None and pred_box.shape[0] == oc.shape[0] > 0
        gaps = np.diff(np.sort(osc.numpy()))
        tie_free = tie_free and (gaps.size == 0 or float(gaps.min()) > 2e-4)
        for t in stat_o:
            PO.tp_fp(oc, osc, gt_box.cpu(), stat_o, t)
    res = ev.summary()
    for t in (0.3, 0.5, 0.7):
        st = ev.stats["all"][t]
        assert st["gt"] == n_gt == stat_o[t]["gt"]
        assert sum(st["tp"]) == sum(stat_o[t]["tp"]) and sum(st["fp"]) == sum(stat_o[t]["fp"]), t
        if tie_free:
            assert st["tp"] == stat_o[t]["tp"] and abs(res["all"][t] - PO.calculate_ap(stat_o, t)) < 1e-9
        else:
            assert abs(res["all"][t] - PO.calculate_ap(stat_o, t)) < 2e-2
        assert 0.0 <= res["all"][t] <= 1.0
    assert res["comm_rate"] > 0


def test_graphed_step_on_dataset_batches_matches_eager():
    """the CUDA-graph replay takes the dataset's sensor-frame clouds + poses (static pose buffer refreshed per step):
    same loss and gradients as the eager fused step over changing scenes of one layout, different point counts included"""
    import a2x_import

    DS, cfg, hypes, model, _ = _setup()
    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    TL = a2x_import.pkg("train_loop")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    eager = model.train()
    graphed = M.Airv2xWhere2com(cfg["model_args"])
    graphed.load_state_dict(sd)
    graphed.cuda().train()
    assigner = a2x_import.pkg("labels").TargetAssigner(hypes["postprocess"], "cuda")
    kw = dict(cameras=False, obj_span=(20.0, 9.0), agent_spread=0.15, pts_sigma=(12.0, 6.0), n_veh=2, n_rsu=1, n_drone=1)
    for step, (seed, npts) in enumerate([(71, 6000), (72, 6000), (73, 5200)]):
        scenes = [DC.synth_scene(DS, seed=seed, n_pts=npts, **kw), DC.synth_scene(DS, seed=seed + 10, n_pts=npts, **kw)]
        _, _, batch = MGD.run_ours(DS, hypes, True, scenes, seed=step)
        ours = batch["ego"]
        assert ours["record_len"].tolist() == [4, 4] and ours["raw_points"]["transforms"] is not None
        labels = assigner(ours["object_bbx_center"], ours["object_bbx_mask"], ours["object_class_ids"])
        random.seed(200 + step)
        l_e = eager.train_step(ours, labels, 1.0, 2.0).clone()
        random.seed(200 + step)
        l_g = graphed.train_step_graphed(ours, labels, 1.0, 2.0).clone()
        assert torch.allclose(l_e, l_g, rtol=1e-6, atol=1e-9), (step, l_e, l_g)
        for (n, pe), (_, pg) in zip(eager.named_parameters(), graphed.named_parameters()):
            if pe.grad is not None:
                assert torch.allclose(pe.grad, pg.grad, rtol=1e-4, atol=1e-6 * float(pe.grad.abs().max()) + 1e-12), (step, n)
    assert len(graphed._graphs) == 1 and graphed.launches_per_step > 200
    # and the Trainer takes that path for dataset batches
    th = dict(hypes, loss={"args": cfg["loss_args"]},
              optimizer={"core_method": "Adam", "lr": 0.002, "args": {"eps": 1e-10, "weight_decay": 1e-4}},
              lr_scheduler={"core_method": "multistep", "gamma": 0.1, "step_size": [10, 25, 40]})
    tr = TL.Trainer(graphed, th)
    assert bool(torch.isfinite(tr.step(ours)).all()) and len(graphed._graphs) >= 1


@pytest.mark.parametrize("name", ["cobevt", "v2xvit"])
def test_dataset_batch_feeds_the_transformer_fusion_models(name):
    """the same batch layout drives Airv2xCoBEVT / Airv2xV2XVit (they read `prior_encoding` / `spatial_correction_matrix` of
    the dataset's collate as well): logits equal to the reference-layout voxel dict of the same clouds"""
    import a2x_import

    if name == "cobevt":
        import cobevt_common as CC
        cfg, gold = CC.load_small()
        model = a2x_import.pkg("opencood.models.airv2x_cobevt").Airv2xCoBEVT(cfg["model_args"])
        model.load_state_dict(CC.golden_state_dict(model, gold))
    else:
        import v2xvit_common as VC
        cfg, gold = VC.load_small()
        model = a2x_import.pkg("opencood.models.airv2x_v2xvit").Airv2xV2XVit(cfg["model_args"])
        model.load_state_dict(VC.golden_state_dict(model, gold))
    model.cuda().eval()
    DS = a2x_import.pkg("intermediate_fusion_dataset")
    hypes = json.load(open(os.path.join(ROOT, "tests", "golden", "dataset_config.json")))
    hypes = dict(hypes, preprocess=cfg["preprocess"], postprocess=cfg["postprocess"],
                 train_params=dict(hypes["train_params"], max_cav=cfg["model_args"]["max_cav"]))
    scene = DC.synth_scene(DS, seed=81, n_veh=2, n_rsu=1, n_drone=1, cameras=False, n_pts=6000, obj_span=(20.0, 9.0),
                           agent_spread=0.15, pts_sigma=(12.0, 6.0))
    _, _, batch = MGD.run_ours(DS, hypes, False, [scene], seed=3)
    ours = batch["ego"]
    L = sum(cfg["model_args"]["max_cav"].values())
    assert ours["prior_encoding"].shape == (1, L, 3) and ours["spatial_correction_matrix"].shape == (1, L, 4, 4)
    ref = _voxel_dict(ours, hypes, False)
    for k in ("prior_encoding", "spatial_correction_matrix", "pairwise_t_matrix_collab"):
        ref[k] = ours[k]
    with torch.no_grad():
        a = model(ours)
        b = model(C.to_device(ref, "cuda"))
    for k in ("psm", "rm", "obj"):
        assert torch.equal(a[k], b[k]) and bool(torch.isfinite(a[k]).all()), k
