"""CPU (-m "not gpu"): the oracle restatement at the FULL size BASELINE.json quotes (config 2: 5 agents x 60 000 points,
200 x 704) against the fixture recorded from the real reference (scripts/make_golden_full.py) — the oracle is pinned at
the size the GPU parity tests (tests/test_gpu_fullsize.py) and the headline use it, not only at 128 x 64."""
import json
import os

import numpy as np
import torch

import fullsize_common as FC
from oracle import w2c_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_equals_reference_at_config2_size():
    import a2x_import

    cfg = json.load(open(os.path.join(ROOT, "configs", "airv2x_intermediate_where2com.json")))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "full_w2c.npz"), allow_pickle=False)
    M = a2x_import.pkg("opencood.models.airv2x_where2com")
    sd = FC.seeded_state_dict(M.Airv2xWhere2com(cfg["model_args"]), FC.W2C_PARAM_SEED, cls_shift=FC.W2C_CLS_SHIFT)
    dd = FC.scene(cfg["preprocess"], training=False)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        out, _ = O.where2com_forward(sd, cfg["model_args"], dd, training=False)
    assert FC.compare_with_golden(out, gold, "eval_", 1e-5) < 1e-5
    assert out["comm_rate"] == int(gold["eval_comm_rate"])
    assert abs(float(out["com"]) - float(gold["eval_com"])) < 1e-7
