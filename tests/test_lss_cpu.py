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


def test_bevencode_oracle_matches_reference_golden():
    """the BEV encoder that consumes the pooled camera features (groundwork for its kernels): oracle == the recorded
    outputs of the real reference BevEncode, eval and train mode"""
    from oracle import bevencode_oracle as BO, w2c_oracle as O

    g = np.load(os.path.join(ROOT, "tests", "golden", "bevencode_small.npz"))
    in_c, out_c, H, W, seed = int(g["in_c"]), int(g["out_c"]), int(g["h"]), int(g["w"]), int(g["seed"])
    import torchvision  # noqa: F401  (state_dict key layout = torchvision resnet18 layer1-3)

    shapes = {"conv1.weight": (64, in_c, 7, 7)}
    for n in ("bn1",):
        shapes.update({n + ".weight": (64,), n + ".bias": (64,), n + ".running_mean": (64,), n + ".running_var": (64,),
                       n + ".num_batches_tracked": ()})

    def bn(n, c):
        return {n + ".weight": (c,), n + ".bias": (c,), n + ".running_mean": (c,), n + ".running_var": (c,),
                n + ".num_batches_tracked": ()}

    cin = 64
    for li, c in ((1, 64), (2, 128), (3, 256)):
        for b in range(2):
            p = "layer%d.%d" % (li, b)
            shapes[p + ".conv1.weight"] = (c, cin if b == 0 else c, 3, 3)
            shapes.update(bn(p + ".bn1", c))
            shapes[p + ".conv2.weight"] = (c, c, 3, 3)
            shapes.update(bn(p + ".bn2", c))
            if b == 0 and li > 1:
                shapes[p + ".downsample.0.weight"] = (c, cin, 1, 1)
                shapes.update(bn(p + ".downsample.1", c))
        cin = c
    shapes["up1.conv.0.weight"] = (256, 320, 3, 3)
    shapes.update(bn("up1.conv.1", 256))
    shapes["up1.conv.3.weight"] = (256, 256, 3, 3)
    shapes.update(bn("up1.conv.4", 256))
    shapes["up2.1.weight"] = (128, 256, 3, 3)
    shapes.update(bn("up2.2", 128))
    shapes["up2.4.weight"] = (out_c, 128, 1, 1)
    shapes["up2.4.bias"] = (out_c,)
    assert len(shapes) == 110
    sd = O.det_init_state_dict(shapes, seed=seed)
    x = torch.randn(2, in_c, H, W, generator=torch.Generator().manual_seed(seed + 1))
    torch.set_num_threads(8)
    with torch.no_grad():
        ev = BO.bev_encode(sd, x, training=False)
        tr = BO.bev_encode(sd, x, training=True, buffers={})
    assert ev.shape == (2, out_c, H, W)
    assert np.abs(ev[:, ::8, ::4, ::4].numpy() - g["eval_out"]).max() < 1e-5
    assert np.abs(tr[:, ::8, ::4, ::4].numpy() - g["train_out"]).max() < 1e-4
