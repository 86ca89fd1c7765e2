"""GPU (-m gpu): the fused lift + voxel-pooling kernel (lss.LiftSplat -> a2x_lift_splat_fwd / _bwd) against the oracle
(pinned to the real reference's voxel_pooling): the same cells are hit, and the sums agree with the float64 pooling at
least as well as the reference's own cumulative-sum trick does; backward == torch autograd of the exact pooling."""
import numpy as np
import pytest
import torch

import test_lss_cpu as T
from oracle import lss_oracle as LO

pytestmark = pytest.mark.gpu


def test_forward_and_backward_match_the_reference_pooling():
    import a2x_import

    L = a2x_import.pkg("lss")
    g, grid, final_dim, down = T.setup()
    B, N, C, rig, geom, depth, feat = T.inputs(g, grid, final_dim, down)
    ls = L.LiftSplat(grid, final_dim, down, "cuda")
    assert torch.equal(ls.frustum.cpu(), LO.create_frustum(final_dim, down, grid["ddiscr"], grid["mode"]))
    dx, bx, nx = LO.gen_dx_bx(grid["xbound"], grid["ybound"], grid["zbound"])
    D, fH, fW = depth.shape[1:]
    x = LO.lift(depth, feat).view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2)
    ref = LO.voxel_pooling(geom, x, dx, bx, nx)                      # == the real reference (scripts/make_golden_lss.py)
    exact = LO.voxel_pooling_exact(geom, x, dx, bx, nx)
    dg, fg = depth.cuda().requires_grad_(True), feat.cuda().requires_grad_(True)
    bev = ls(dg, fg, geom.cuda())
    got = bev.detach().cpu()
    assert got.shape == ref.shape
    assert torch.equal(got.abs().sum(1) > 0, ref.abs().sum(1) > 0)   # exactly the reference's cells are populated
    err_ref = float((ref.double() - exact).abs().max())
    err_got = float((got.double() - exact).abs().max())
    print("max |reference - f64| %.2e, |kernel - f64| %.2e, |kernel - reference| %.2e" % (err_ref, err_got, float((got - ref).abs().max())))
    assert err_got <= 2e-6 * float(exact.abs().max()) + 1e-6 and err_got <= err_ref + 1e-6
    assert float((got - ref).abs().max()) <= err_ref + err_got + 1e-6
    # geometry on the device == the oracle's up to fp32 rounding of the 3x3 products
    geom_dev = ls.geometry(*rig).cpu()
    assert float((geom_dev - geom).abs().max()) < 1e-3
    # backward vs autograd through the exact pooling
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    (bev * w.cuda()).sum().backward()
    d64, f64 = depth.double().requires_grad_(True), feat.double().requires_grad_(True)
    x64 = LO.lift(d64, f64).view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2)
    (LO.voxel_pooling_exact(geom, x64, dx, bx, nx) * w.double()).sum().backward()
    rel = lambda a, b: float((a.double().cpu() - b).abs().max() / (b.abs().max() + 1e-30))
    assert rel(dg.grad, d64.grad) < 1e-5 and rel(fg.grad, f64.grad) < 1e-5


def test_full_size_properties():
    """config-5-like size (6 cameras, 45 x 80 feature map, 48 bins, 704 x 200 BEV): linearity in the features and mass
    conservation — sum over the BEV == sum over the kept frustum points of depth * feature"""
    import a2x_import

    L = a2x_import.pkg("lss")
    grid = {"xbound": [-140.8, 140.8, 0.4], "ybound": [-40, 40, 0.4], "zbound": [-10, 10, 20.0], "ddiscr": [2, 50, 48], "mode": "LID"}
    final_dim, down, B, N, C = [360, 640], 8, 1, 6, 64
    ls = L.LiftSplat(grid, final_dim, down, "cuda")
    rig = LO.synth_cameras(B, N, 3, final_dim)
    geom = ls.geometry(*rig)
    gen = torch.Generator().manual_seed(2)
    depth = torch.softmax(torch.randn(B * N, ls.D, ls.fH, ls.fW, generator=gen), 1).cuda()
    f1 = torch.randn(B * N, C, ls.fH, ls.fW, generator=gen).cuda()
    f2 = torch.randn(B * N, C, ls.fH, ls.fW, generator=gen).cuda()
    a, b, ab = ls(depth, f1, geom), ls(depth, f2, geom), ls(depth, f1 + 2 * f2, geom)
    assert a.shape == (1, 64, 200, 704)
    assert float((ab - (a + 2 * b)).abs().max()) < 1e-4 * float(ab.abs().max())
    dx, bx, nx = LO.gen_dx_bx(grid["xbound"], grid["ybound"], grid["zbound"])
    _, kept = LO.voxel_cells(geom.cpu(), dx, bx, nx)
    kept = kept.view(B * N, ls.D, ls.fH, ls.fW).cuda()
    mass = torch.einsum("ndhw,nchw->c", (depth * kept).double(), f1.double())
    assert float((a.double().sum((0, 2, 3)) - mass).abs().max()) < 1e-6 * float(mass.abs().max()) + 1e-3


def test_bevencode_eval_matches_reference_golden():
    """BevEncode (sub_modules/lss_submodule.py:312-349) on the tap-GEMM kernels — 7x7 stride-2 stem as 49 taps over the parity
    views, BasicBlock residual adds in the GEMM epilogue, bilinear x4 / x2 (align_corners=True) on the ego-warp kernel —
    against the output recorded from the REAL reference module and, everywhere, against the oracle. Tolerance 1e-3."""
    import os

    import a2x_import
    from oracle import bevencode_oracle as BO, w2c_oracle as O

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = np.load(os.path.join(root, "tests", "golden", "bevencode_small.npz"))
    in_c, out_c, H, W, seed = int(g["in_c"]), int(g["out_c"]), int(g["h"]), int(g["w"]), int(g["seed"])
    L = a2x_import.pkg("lss")
    m = L.BevEncode(in_c, out_c)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = m.state_dict()
    sd.update(O.det_init_state_dict(shapes, seed=seed))
    m.load_state_dict(sd)                       # same keys and shapes as the reference module (strict)
    x = torch.randn(2, in_c, H, W, generator=torch.Generator().manual_seed(seed + 1))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.eval()(x)
    m.cuda().eval()
    out = m(x.cuda())
    assert tuple(out.shape) == (2, out_c, H, W)
    err_g = float(np.abs(out[:, ::8, ::4, ::4].cpu().numpy() - g["eval_out"]).max())
    with torch.no_grad():
        ora = BO.bev_encode({k: v.cpu() for k, v in m.state_dict().items()}, x, training=False)
    err_o = float((out.cpu() - ora).abs().max())
    print("BevEncode: vs golden %.2e, vs oracle %.2e (|out| max %.2f)" % (err_g, err_o, float(ora.abs().max())))
    assert err_g < 1e-3 and err_o < 1e-3
    out2 = L.BevEncode(in_c, 5).cuda().eval()(x.cuda())          # an output width that is not a multiple of 32
    assert tuple(out2.shape) == (2, 5, H, W)
    m.train()
    with pytest.raises(NotImplementedError):
        m(x.cuda())


def test_camera_branch_matches_oracle_composition():
    """lss.CameraBranch = LiftSplatShootEncoder.forward (common_modules/airv2x_encoder.py:308-335) with a pluggable trunk:
    geometry -> lift + voxel pooling -> BevEncode, against the oracle's composition of the same pieces (each pinned to the
    real reference). The stub trunk returns a seeded depth distribution and feature map."""
    import a2x_import
    from oracle import bevencode_oracle as BO, w2c_oracle as O

    L = a2x_import.pkg("lss")
    grid = {"xbound": [-16.0, 16.0, 0.5], "ybound": [-12.0, 12.0, 0.5], "zbound": [-10.0, 10.0, 20.0], "ddiscr": [2, 30, 24], "mode": "LID"}
    final_dim, down, camC, outC, B, N = [64, 96], 8, 64, 64, 2, 3
    args = {"grid_conf": grid, "data_aug_conf": {"final_dim": final_dim}, "img_downsample": down, "img_features": camC,
            "bevout_feature": outC}
    fr = LO.create_frustum(final_dim, down, grid["ddiscr"], grid["mode"])
    D, fH, fW = fr.shape[:3]
    gen = torch.Generator().manual_seed(5)
    depth = torch.softmax(torch.randn(B * N, D, fH, fW, generator=gen) * 2, 1)
    feat = torch.randn(B * N, camC, fH, fW, generator=gen)
    rig = LO.synth_cameras(B, N, 9, final_dim)
    branch = L.CameraBranch(args, "vehicle", lambda imgs: (depth.cuda(), feat.cuda()))
    shapes = {k: tuple(v.shape) for k, v in branch.bevencode.state_dict().items()}
    sd = branch.bevencode.state_dict()
    sd.update(O.det_init_state_dict(shapes, seed=3))
    branch.bevencode.load_state_dict(sd)
    branch.cuda().eval()
    dd = {"vehicle": {"batch_merged_cam_inputs": {"imgs": torch.zeros(B, N, 3, *final_dim), "rots": rig[0], "trans": rig[1],
                                                  "intrinsics": rig[2], "post_rots": rig[3], "post_trans": rig[4]}}}
    out = branch(dd)["spatial_features"]
    dx, bx, nx = LO.gen_dx_bx(grid["xbound"], grid["ybound"], grid["zbound"])
    geom = LO.get_geometry(fr, *rig)
    x = LO.lift(depth, feat).view(B, N, camC, D, fH, fW).permute(0, 1, 3, 4, 5, 2)
    pooled = LO.voxel_pooling_exact(geom, x, dx, bx, nx).float()
    with torch.no_grad():
        ref = BO.bev_encode({k: v.cpu() for k, v in branch.bevencode.state_dict().items()}, pooled, training=False)
    assert tuple(out.shape) == tuple(ref.shape) == (B, outC, 48, 64)
    err = float((out.cpu() - ref).abs().max())
    print("CameraBranch vs oracle composition: %.2e (|ref| max %.2f)" % (err, float(ref.abs().max())))
    assert err < 1e-3
