"""Timing of the fused lift + voxel-pooling kernels at config-5-like size (6 cameras, 360 x 640 images / 8 -> 45 x 80
features, 48 depth bins, C = 64, 704 x 200 x 1 BEV) next to the reference formulation in stock PyTorch on the same GPU
(materialised product, rank sort, cumulative-sum pooling — oracle/lss_oracle.py moved to the device). One JSON line.

    python scripts/bench_lift_splat.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import a2x_import
from oracle import lss_oracle as LO


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def torch_pooling(geom, x, dx, bx, nx):
    """airv2x_encoder.py:208-275 as written, on the device"""
    B, N, D, H, W, C = x.shape
    Np = B * N * D * H * W
    x = x.reshape(Np, C)
    cells = ((geom - (bx - dx / 2.0)) / dx).long().view(Np, 3)
    bix = torch.cat([torch.full([Np // B, 1], ix, device=x.device, dtype=torch.long) for ix in range(B)])
    g = torch.cat((cells, bix), 1)
    kept = (g[:, 0] >= 0) & (g[:, 0] < nx[0]) & (g[:, 1] >= 0) & (g[:, 1] < nx[1]) & (g[:, 2] >= 0) & (g[:, 2] < nx[2])
    x, g = x[kept], g[kept]
    ranks = g[:, 0] * (nx[1] * nx[2] * B) + g[:, 1] * (nx[2] * B) + g[:, 2] * B + g[:, 3]
    s = ranks.argsort()
    x, g, ranks = x[s], g[s], ranks[s]
    x = x.cumsum(0)
    last = torch.ones(x.shape[0], device=x.device, dtype=torch.bool)
    last[:-1] = ranks[1:] != ranks[:-1]
    x, g = x[last], g[last]
    x = torch.cat((x[:1], x[1:] - x[:-1]))
    final = torch.zeros((B, C, int(nx[2]), int(nx[1]), int(nx[0])), device=x.device)
    final[g[:, 3], :, g[:, 2], g[:, 1], g[:, 0]] = x
    return torch.cat(final.unbind(dim=2), 1)


def main():
    L = a2x_import.pkg("lss")
    grid = {"xbound": [-140.8, 140.8, 0.4], "ybound": [-40, 40, 0.4], "zbound": [-10, 10, 20.0], "ddiscr": [2, 50, 48], "mode": "LID"}
    final_dim, down, B, N, C = [360, 640], 8, 1, 6, 64
    ls = L.LiftSplat(grid, final_dim, down, "cuda")
    geom = ls.geometry(*LO.synth_cameras(B, N, 3, final_dim))
    gen = torch.Generator().manual_seed(2)
    depth = torch.softmax(torch.randn(B * N, ls.D, ls.fH, ls.fW, generator=gen), 1).cuda().requires_grad_(True)
    feat = torch.randn(B * N, C, ls.fH, ls.fW, generator=gen).cuda().requires_grad_(True)
    w = torch.randn(B, C, 200, 704, device="cuda")
    dx, bx, nx = [t.cuda() for t in LO.gen_dx_bx(grid["xbound"], grid["ybound"], grid["zbound"])]

    def ours_fwd():
        with torch.no_grad():
            return ls(depth, feat, geom)

    def ours_fb():
        depth.grad = feat.grad = None
        (ls(depth, feat, geom) * w).sum().backward()

    def ref_fwd():
        with torch.no_grad():
            x = LO.lift(depth, feat).view(B, N, C, ls.D, ls.fH, ls.fW).permute(0, 1, 3, 4, 5, 2)
            return torch_pooling(geom, x, dx, bx, nx)

    a, b = ours_fwd(), ref_fwd()
    err = float((a - b).abs().max())
    pts = B * N * ls.D * ls.fH * ls.fW
    print(json.dumps({"metric": "lift + voxel pooling, %d cameras, %d frustum points, C = %d" % (N, pts, C),
                      "ours_fwd_ms": timed(ours_fwd), "ours_fwd_bwd_ms": timed(ours_fb), "stock_pytorch_fwd_ms": timed(ref_fwd, 5),
                      "max_abs_diff_vs_stock": err,
                      "algorithmic_bytes_fwd": int(pts * 4 * (1 + 3) + B * N * C * ls.fH * ls.fW * 4 + B * 200 * 704 * C * 4)}))


if __name__ == "__main__":
    main()
