"""Pin the oracle's NaiveCompressor (oracle/cobevt_oracle.py:naive_compressor) in TRAIN mode against the REAL reference
module (opencood/models/common_modules/naive_compress.py): forward with batch statistics, running-stat updates, input /
parameter gradients. Groundwork for training with `compression > 0`. Writes tests/golden/compressor_train.npz.

    python scripts/make_golden_compressor_train.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cobevt_oracle as CO, ref_import, w2c_oracle as O  # noqa: E402

C, R, SEED = 256, 2, 55


def inputs():
    g = torch.Generator().manual_seed(SEED + 1)
    return torch.randn(3, C, 12, 20, generator=g), torch.randn(3, C, 12, 20, generator=g)


def main():
    ref_import.install()
    from opencood.models.common_modules.naive_compress import NaiveCompressor

    torch.manual_seed(0)
    m = NaiveCompressor(C, R)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = m.state_dict()
    sd.update(O.det_init_state_dict(shapes, seed=SEED))
    m.load_state_dict(sd)
    sd = {"naive_compressor." + k: v.clone() for k, v in m.state_dict().items()}
    x, w = inputs()
    m.train()
    a = x.clone().requires_grad_(True)
    ya = m(a)
    (ya * w).sum().backward()
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    b = x.clone().requires_grad_(True)
    bufs = {}
    yb = CO.naive_compressor(p, b, True, bufs)
    (yb * w).sum().backward()
    print("train forward ref-vs-oracle %.3e, dx %.3e" % (float((ya - yb).abs().max()), float((a.grad - b.grad).abs().max())))
    assert float((ya - yb).abs().max()) < 1e-5 and float((a.grad - b.grad).abs().max()) < 1e-5
    worst = 0.0
    for n, q in m.named_parameters():
        worst = max(worst, float((q.grad - p["naive_compressor." + n].grad).abs().max() / (q.grad.abs().max() + 1e-30)))
    print("parameter gradients: worst relative difference %.3e; running stats %.3e" % (
        worst, max(float((m.state_dict()[k[len("naive_compressor."):]] - v).abs().max()) for k, v in bufs.items())))
    assert worst < 1e-4
    dst = os.path.join(ROOT, "tests", "golden", "compressor_train.npz")
    np.savez_compressed(dst, seed=SEED, c=C, r=R, y=ya.detach()[:, ::16, ::2, ::2].numpy(), dx=a.grad[:, ::16, ::2, ::2].numpy(),
                        dw_enc=m.encoder[0].weight.grad[::8, ::16].numpy())
    print("wrote", dst)


if __name__ == "__main__":
    main()
