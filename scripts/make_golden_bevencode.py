"""Pin oracle/bevencode_oracle.py against the REAL reference `BevEncode` (opencood/models/sub_modules/lss_submodule.py)
on the CPU, eval and train mode, seeded weights, and write tests/golden/bevencode_small.npz.

    python scripts/make_golden_bevencode.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bevencode_oracle as BO, ref_import, w2c_oracle as O  # noqa: E402

IN_C, OUT_C, H, W, SEED = 64, 64, 48, 80, 77


def main():
    ref_import.install()
    from opencood.models.sub_modules.lss_submodule import BevEncode

    torch.manual_seed(0)
    torch.set_num_threads(8)
    m = BevEncode(IN_C, OUT_C)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = m.state_dict()
    sd.update(O.det_init_state_dict(shapes, seed=SEED))
    m.load_state_dict(sd)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    print("BevEncode params", sum(p.numel() for p in m.parameters()), "state_dict entries", len(sd))
    x = torch.randn(2, IN_C, H, W, generator=torch.Generator().manual_seed(SEED + 1))
    out = {"seed": SEED, "in_c": IN_C, "out_c": OUT_C, "h": H, "w": W}
    m.eval()
    with torch.no_grad():
        ref = m(x)
        ora = BO.bev_encode(sd, x, training=False)
    print("eval: ref-vs-oracle %.3e (|ref| max %.3f)" % (float((ref - ora).abs().max()), float(ref.abs().max())))
    assert float((ref - ora).abs().max()) < 1e-5
    out["eval_out"] = ref[:, ::8, ::4, ::4].numpy()
    m.train()
    bufs = {}
    with torch.no_grad():
        ref_t = m(x)
        ora_t = BO.bev_encode(sd, x, training=True, buffers=bufs)
    print("train: ref-vs-oracle %.3e" % float((ref_t - ora_t).abs().max()))
    assert float((ref_t - ora_t).abs().max()) < 1e-4
    worst = max(float((m.state_dict()[k] - v).abs().max()) for k, v in bufs.items())
    print("running statistics after one train-mode forward: max diff %.3e over %d buffers" % (worst, len(bufs)))
    assert worst < 1e-5
    out["train_out"] = ref_t[:, ::8, ::4, ::4].numpy()
    dst = os.path.join(ROOT, "tests", "golden", "bevencode_small.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, "%.1f KB" % (os.path.getsize(dst) / 1024))


if __name__ == "__main__":
    main()
