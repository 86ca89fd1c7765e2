"""Pin oracle/lss_oracle.py against the REAL reference's Lift-Splat geometry and voxel pooling
(LiftSplatShootEncoder.create_frustum / get_geometry / voxel_pooling called unbound on a parameter namespace: the class
itself hard-codes .to("cuda") in __init__) and write tests/golden/lss_small.npz.

    python scripts/make_golden_lss.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lss_oracle as LO, ref_import  # noqa: E402

GRID = {"xbound": [-25.6, 25.6, 0.4], "ybound": [-12.8, 12.8, 0.4], "zbound": [-10, 10, 20.0], "ddiscr": [2, 50, 48], "mode": "LID"}
FINAL_DIM, DOWNSAMPLE, CAMC, B, N, SEED = [96, 160], 8, 64, 2, 4, 11


def main():
    ref_import.install()
    from opencood.models.common_modules.airv2x_encoder import LiftSplatShootEncoder as E
    from opencood.utils import camera_utils as CU

    dx, bx, nx = CU.gen_dx_bx(GRID["xbound"], GRID["ybound"], GRID["zbound"])
    ns = types.SimpleNamespace(grid_conf=GRID, data_aug_conf={"final_dim": FINAL_DIM}, downsample=DOWNSAMPLE, dx=dx, bx=bx,
                               nx=nx, use_quickcumsum=True)
    ns.frustum = E.create_frustum(ns)
    rig = LO.synth_cameras(B, N, SEED, FINAL_DIM)
    geom = E.get_geometry(ns, *rig)
    fr = LO.create_frustum(FINAL_DIM, DOWNSAMPLE, GRID["ddiscr"], GRID["mode"])
    assert torch.equal(fr, ns.frustum)
    o_dx, o_bx, o_nx = LO.gen_dx_bx(GRID["xbound"], GRID["ybound"], GRID["zbound"])
    assert torch.equal(o_dx, dx) and torch.equal(o_bx, bx) and torch.equal(o_nx, nx)
    assert torch.equal(LO.get_geometry(fr, *rig), geom)
    D, fH, fW = fr.shape[:3]
    g = torch.Generator().manual_seed(SEED + 1)
    depth = torch.softmax(torch.randn(B * N, D, fH, fW, generator=g) * 2, 1)
    feat = torch.randn(B * N, CAMC, fH, fW, generator=g)
    x = LO.lift(depth, feat).view(B, N, CAMC, D, fH, fW).permute(0, 1, 3, 4, 5, 2)       # get_cam_feats' layout
    ref = E.voxel_pooling(ns, geom, x)
    ora = LO.voxel_pooling(geom, x, dx, bx, nx)
    assert torch.equal(ref, ora), float((ref - ora).abs().max())
    exact = LO.voxel_pooling_exact(geom, x, dx, bx, nx)
    err = float((ref.double() - exact).abs().max())
    cells, kept = LO.voxel_cells(geom, dx, bx, nx)
    print("frustum %s, %d of %d points inside the grid, BEV %s, |reference - float64 sum| max %.3e (|BEV| max %.3f)"
          % (tuple(fr.shape), int(kept.sum()), kept.numel(), tuple(ref.shape), err, float(ref.abs().max())))
    nzi = torch.nonzero(ref.abs().sum(1).view(-1)).view(-1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lss_small.npz"), seed=SEED, B=B, N=N, camC=CAMC,
                        final_dim=np.array(FINAL_DIM), downsample=DOWNSAMPLE, xbound=np.array(GRID["xbound"]),
                        ybound=np.array(GRID["ybound"]), zbound=np.array(GRID["zbound"]), ddiscr=np.array(GRID["ddiscr"]),
                        geom_sample=geom.view(-1, 3)[::997].numpy(), nonzero_cells=nzi.numpy().astype(np.int32),
                        bev_sample=ref.permute(0, 2, 3, 1).reshape(-1, ref.shape[1])[nzi[::7]].numpy())
    print("oracle == reference; wrote tests/golden/lss_small.npz")


if __name__ == "__main__":
    main()
