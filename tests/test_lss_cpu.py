"""CPU (-m "not gpu"): the Lift-Splat oracle against the golden vectors recorded from the REAL reference functions
(scripts/make_golden_lss.py), and the host half of lss.LiftSplat (frustum, grid) against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import lss_oracle as LO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def setup():
    g = np.load(os.path.join(ROOT, "tests", "golden", "lss_small.npz"))
    grid = {"xbound": g["xbound"].tolist(), "ybound": g["ybound"].tolist(), "zbound": g["zbound"].tolist(),
            "ddiscr": [int(g["ddiscr"][0]), int(g["ddiscr"][1]), int(g["ddiscr"][2])], "mode": "LID"}
    return g, grid, [int(v) for v in g["final_dim"]], int(g["downsample"])


def inputs(g, grid, final_dim, down):
    B, N, C, seed = int(g["B"]), int(g["N"]), int(g["camC"]), int(g["seed"])
    fr = LO.create_frustum(final_dim, down, grid["ddiscr"], grid["mode"])
    rig = LO.synth_cameras(B, N, seed, final_dim)
    geom = LO.get_geometry(fr, *rig)
    D, fH, fW = fr.shape[:3]
    gen = torch.Generator().manual_seed(seed + 1)
    depth = torch.softmax(torch.randn(B * N, D, fH, fW, generator=gen) * 2, 1)
    feat = torch.randn(B * N, C, fH, fW, generator=gen)
    return B, N, C, rig, geom, depth, feat


def test_oracle_reproduces_reference_geometry_and_pooling():
    g, grid, final_dim, down = setup()
    B, N, C, rig, geom, depth, feat = inputs(g, grid, final_dim, down)
    assert np.array_equal(geom.view(-1, 3)[::997].numpy(), g["geom_sample"])
    dx, bx, nx = LO.gen_dx_bx(grid["xbound"], grid["ybound"], grid["zbound"])
    D, fH, fW = depth.shape[1:]
    x = LO.lift(depth, feat).view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2)
    bev = LO.voxel_pooling(geom, x, dx, bx, nx)
    nzi = torch.nonzero(bev.abs().sum(1).view(-1)).view(-1)
    assert np.array_equal(nzi.numpy().astype(np.int32), g["nonzero_cells"])
    assert np.array_equal(bev.permute(0, 2, 3, 1).reshape(-1, bev.shape[1])[nzi[::7]].numpy(), g["bev_sample"])
    exact = LO.voxel_pooling_exact(geom, x, dx, bx, nx)
    assert float((bev.double() - exact).abs().max()) < 1e-4          # the cumulative-sum trick's own cancellation error


def test_host_half_matches_oracle():
    import a2x_import

    L = a2x_import.pkg("lss")
    g, grid, final_dim, down = setup()
    dx, bx, nx = LO.gen_dx_bx(grid["xbound"], grid["ybound"], grid["zbound"])
    dx2, bx2, nx2 = L.gen_dx_bx(grid["xbound"], grid["ybound"], grid["zbound"])
    assert torch.equal(dx, dx2) and torch.equal(bx, bx2) and [int(v) for v in nx] == nx2
    assert np.array_equal(L.depth_bins(*grid["ddiscr"], "LID"), LO.depth_discretization(*grid["ddiscr"], "LID"))
    with pytest.raises(RuntimeError, match="CUDA"):
        L.LiftSplat(grid, final_dim, down, "cpu")
