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
