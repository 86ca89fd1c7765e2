"""CPU: host-side logic of the drop-in module (scene-major agent layout) against the oracle's regrouping."""
import json
import os

import torch

import w2c_common as C
from oracle import w2c_oracle as O


def test_layout_matches_reference_regrouping(pkg):
    import a2x_import

    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    cfg, _ = C.load_small()
    args = cfg["model_args"]
    model = M.Airv2xWhere2com(args)
    # 3 scenes: scene 1 has no RSU, scene 2 has no drone
    dd = {"vehicle": {"record_len": torch.tensor([2, 1, 1]), "batch_idxs": [0, 1, 2]},
          "rsu": {"record_len": torch.tensor([1, 0, 2]), "batch_idxs": [0, 2]},
          "drone": {"record_len": torch.tensor([1, 1, 0]), "batch_idxs": [0, 1]}}
    for t, n in (("vehicle", 4), ("rsu", 3), ("drone", 2)):
        dd[t]["batch_merged_lidar_features_torch"] = {"n_agents": n}
    lay = model._layout(dd, torch.device("cpu"))
    assert lay["record_len"] == [4, 2, 3] and lay["n_total"] == 9
    # oracle: tag every agent map with a unique id and regroup the way airv2x_base_model.py:179-248 does
    nx, ny = lay["nx"], lay["ny"]
    sd, tags = {}, {}
    per_type = {}
    uid = 0
    for t, n in (("vehicle", 4), ("rsu", 3), ("drone", 2)):
        per_type[t] = list(range(uid, uid + n))
        uid += n
    order = []
    for b in range(3):
        for t in O.AGENT_TYPES:
            idxs = dd[t]["batch_idxs"]
            if b not in idxs:
                continue
            rl = dd[t]["record_len"]
            rl = rl[rl > 0]
            cs = torch.cumsum(rl, 0)
            ti = idxs.index(b)
            start = 0 if ti == 0 else int(cs[ti - 1])
            order += per_type[t][start:int(cs[ti])]
    got = [None] * 9
    for t in per_type:
        for j, row in enumerate(lay["agent_map"][t].tolist()):
            got[row] = per_type[t][j]
    assert got == order
    assert lay["scene_start"].tolist() == [0, 4, 6]
    assert lay["ego_flags"].tolist() == [1, 0, 0, 0, 1, 0, 1, 0, 0]


def test_config_json_matches_yaml_keys():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = json.load(open(os.path.join(root, "configs", "airv2x_intermediate_where2com.json")))
    a = cfg["model_args"]
    assert a["vehicle"]["lidar"]["point_pillar_scatter"]["grid_size"] == [704, 200, 1]
    assert a["modality_fusion"]["base_bev_backbone"]["layer_nums"] == [3, 5, 8]
    assert a["where2com_fusion"]["communication"]["threshold"] == 0.01
    assert cfg["preprocess"]["args"]["max_voxel_train"] == 32000


def test_agent_parallel_exchange_regions():
    """layout of the one buffer an agent publishes (header | cell indices | selected level-0 rows | dense deeper levels)"""
    import json
    import os

    import a2x_import

    E = a2x_import.pkg("w2c_engine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = json.load(open(os.path.join(root, "configs", "airv2x_intermediate_where2com.json")))
    eng = E.W2CEngine(cfg["model_args"], "cpu")
    reg = eng.ap_regions(100, 352)
    assert reg["hdr"] == (0, 64) and reg["idx"] == (64, 35200) and reg["vals"] == (64 + 35200, 35200 * 64)
    assert reg["lvl1"][2] == (50, 176, 128) and reg["lvl2"][2] == (25, 88, 256)
    assert reg["total"] == 64 + 35200 + 35200 * 64 + 50 * 176 * 128 + 25 * 88 * 256
    for k in ("idx", "vals", "lvl1", "lvl2"):
        assert reg[k][0] % 4 == 0          # 16-byte aligned regions (float4 / int4 access)


def test_checkpoint_resume_helpers(tmp_path):
    """train_loop.find_last_checkpoint / load_saved_model: the latest `net_epochN.pth` is found (the reference's
    findLastCheckpoint returns an undefined name, train_utils.py:54-63) and BOTH layouts load — the dict train.py writes
    (which the reference's loader silently drops key by key, :88-116) and a flat / DataParallel-prefixed state_dict."""
    import torch
    import torch.nn as nn

    import a2x_import

    TL = a2x_import.pkg("train_loop")
    d = str(tmp_path)
    assert TL.find_last_checkpoint(d) == 0
    net = nn.Sequential(nn.Linear(4, 3), nn.BatchNorm1d(3))
    opt = torch.optim.Adam(net.parameters(), lr=0.002, eps=1e-10, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[1, 3], gamma=0.1)
    net(torch.randn(5, 4)).sum().backward()
    opt.step()
    sch.step()
    torch.save({"epoch": 2, "model_state_dict": net.state_dict(), "optimizer_state_dict": opt.state_dict(),
                "scheduler_state_dict": sch.state_dict()}, d + "/net_epoch3.pth")
    torch.save({"module." + k: v for k, v in net.state_dict().items()}, d + "/net_epoch12.pth")
    torch.save({"x": 1}, d + "/net_epoch_bestval_at7.pth")                      # not an epoch checkpoint
    assert TL.find_last_checkpoint(d) == 12
    fresh = nn.Sequential(nn.Linear(4, 3), nn.BatchNorm1d(3))
    ep, _ = TL.load_saved_model(d, fresh)                                        # flat, `module.`-prefixed layout
    assert ep == 12 and all(torch.equal(a, b) for a, b in zip(fresh.state_dict().values(), net.state_dict().values()))
    fresh = nn.Sequential(nn.Linear(4, 3), nn.BatchNorm1d(3))
    opt2 = torch.optim.Adam(fresh.parameters(), lr=0.5)
    sch2 = torch.optim.lr_scheduler.MultiStepLR(opt2, milestones=[1, 3], gamma=0.1)
    ep, _ = TL.load_saved_model(d, fresh, epoch=3, optimizer=opt2, scheduler=sch2)  # the dict layout of train.py
    assert ep == 3 and all(torch.equal(a, b) for a, b in zip(fresh.state_dict().values(), net.state_dict().values()))
    assert opt2.param_groups[0]["lr"] == opt.param_groups[0]["lr"] and sch2.last_epoch == sch.last_epoch
    assert len(opt2.state) == len(opt.state)
    hypes = {"optimizer": {"core_method": "Adam", "lr": 0.002, "args": {"eps": 1e-10, "weight_decay": 1e-4}},
             "lr_scheduler": {"core_method": "multistep", "gamma": 0.1, "step_size": [10, 25, 40]}}
    o = TL.setup_optimizer(hypes, net)
    assert type(o).__name__ == "Adam" and o.defaults["eps"] == 1e-10 and o.defaults["weight_decay"] == 1e-4
    s = TL.setup_lr_scheduler(hypes, o, init_epoch=11)
    assert abs(o.param_groups[0]["lr"] - 0.0002) < 1e-12 and s.last_epoch == 11


def test_cosine_warmup_schedule_of_the_legacy_yamls(pkg):
    """`lr_scheduler: cosineannealwarm` (V2XR_*.yaml): the update-counting cosine schedule with linear warm-up that
    `setup_lr_schedular` builds on timm (train_utils.py:430-447) — closed-form values, timm's construction side effect (the
    rate starts at `warmup_lr`), its no-op epoch step, state round trip"""
    import math

    import pytest
    import torch

    import a2x_import

    TL = a2x_import.pkg("train_loop")
    hypes = {"lr_scheduler": {"core_method": "cosineannealwarm", "epoches": 4, "warmup_lr": 2e-5, "warmup_epoches": 1, "lr_min": 5e-6}}
    p = torch.nn.Parameter(torch.zeros(3))
    opt = torch.optim.Adam([p], lr=2e-3)
    with pytest.raises(ValueError):
        TL.setup_lr_scheduler(hypes, opt)
    sch = TL.setup_lr_scheduler(hypes, opt, n_iter_per_epoch=10)
    lr = lambda: opt.param_groups[0]["lr"]  # noqa: E731
    assert lr() == 2e-5 and opt.param_groups[0]["initial_lr"] == 2e-3          # constructing it drops the rate to warmup_lr
    for e in range(3):
        sch.step(e)                                                          # the epoch-level call of tools/train.py:289
    assert lr() == 2e-5
    sch.step_update(5)
    assert abs(lr() - (2e-5 + 5 * (2e-3 - 2e-5) / 10)) < 1e-12                 # linear warm-up over 10 updates
    sch.step_update(10)
    assert abs(lr() - (5e-6 + (2e-3 - 5e-6) * 0.5 * (1 + math.cos(math.pi * 10 / 40)))) < 1e-12
    sch.step_update(20)
    assert abs(lr() - (5e-6 + (2e-3 - 5e-6) * 0.5)) < 1e-12                    # half way: the mean of lr and lr_min
    sch.step_update(39)
    assert 5e-6 < lr() < 2e-5
    sch.step_update(40)
    assert lr() == 5e-6
    sch.step_update(400)
    assert lr() == 5e-6                                                      # one cycle only
    lrs = []
    for t in range(41):
        sch.step_update(t)
        lrs.append(lr())
    # rises through the warm-up, then (no warm-up prefix: the cosine is evaluated at t, not t - warmup_t) falls to lr_min
    assert all(a < b for a, b in zip(lrs[:9], lrs[1:10])) and all(a > b for a, b in zip(lrs[10:40], lrs[11:41]))
    state = sch.state_dict()
    assert "optimizer" not in state and state["num_updates"] == 40
    opt2 = torch.optim.Adam([torch.nn.Parameter(torch.zeros(3))], lr=2e-3)
    sch2 = TL.setup_lr_scheduler(hypes, opt2, n_iter_per_epoch=10)
    sch2.load_state_dict(state)
    sch2.step_update(20)
    assert abs(opt2.param_groups[0]["lr"] - (5e-6 + (2e-3 - 5e-6) * 0.5)) < 1e-12
    # the other schedules are torch's own
    for m, cls in (("step", "StepLR"), ("multistep", "MultiStepLR"), ("exponential", "ExponentialLR")):
        s = TL.setup_lr_scheduler({"lr_scheduler": {"core_method": m, "gamma": 0.1, "step_size": [1, 2] if m == "multistep" else 2}},
                                  torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1e-3))
        assert type(s).__name__ == cls
