"""TEST INFRASTRUCTURE ONLY: CPU restatement of the Lift-Splat camera branch's geometry + voxel pooling (SURVEY §8f-4,
BASELINE config 5's camera half). Only tests/ and scripts/ may import it; the product path never does.

Follows:
  opencood/models/common_modules/airv2x_encoder.py:94-131   create_frustum
  opencood/models/common_modules/airv2x_encoder.py:133-168  get_geometry
  opencood/models/common_modules/airv2x_encoder.py:208-275  voxel_pooling (truncating `.long()` cell index, rank sort,
                                                             cumulative-sum pooling, z folded into the channels)
  opencood/models/sub_modules/lss_submodule.py:170-186      the "lift": depth.unsqueeze(1) * x_img.unsqueeze(2)
  opencood/utils/camera_utils.py:238-244, :310-326, :328-365  gen_dx_bx, depth_discretization, cumsum_trick / QuickCumsum
Pinned against the REAL reference functions by scripts/make_golden_lss.py (the encoder class hard-codes .to("cuda") in
__init__, airv2x_encoder.py:47-61, so its methods are called unbound on a parameter namespace): geometry and pooled BEV
identical. The EfficientNet trunk that produces `depth` / `x_img` is a library call outside this oracle (its pretrained
weights are not available offline): parity of the trunk is unpinned, parity of lift + splat is pinned.
"""
import numpy as np
import torch


def gen_dx_bx(xbound, ybound, zbound):
    dx = torch.Tensor([row[2] for row in [xbound, ybound, zbound]])
    bx = torch.Tensor([row[0] + row[2] / 2.0 for row in [xbound, ybound, zbound]])
    nx = torch.LongTensor([int((row[1] - row[0]) / row[2] + 0.5) for row in [xbound, ybound, zbound]])
    return dx, bx, nx


def depth_discretization(depth_min, depth_max, num_bins, mode):
    """camera_utils.py:310-326"""
    if mode == "UD":
        return np.linspace(depth_min, depth_max, num_bins, endpoint=False)
    indices = np.arange(0, num_bins)
    if mode == "LID":
        bin_size = 2 * (depth_max - depth_min) / (num_bins * (1 + num_bins))
        return depth_min + bin_size * (indices * (indices + 1)) / 2
    raise NotImplementedError(mode)


def create_frustum(final_dim, downsample, ddiscr, mode):
    ogfH, ogfW = final_dim
    fH, fW = ogfH // downsample, ogfW // downsample
    ds = torch.tensor(depth_discretization(*ddiscr, mode), dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = ds.shape[0]
    xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    return torch.stack((xs, ys, ds), -1)


def get_geometry(frustum, rots, trans, intrins, post_rots, post_trans):
    B, N, _ = trans.shape
    points = frustum - post_trans.view(B, N, 1, 1, 1, 3)
    points = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(points.unsqueeze(-1))
    points = torch.cat((points[:, :, :, :, :, :2] * points[:, :, :, :, :, 2:3], points[:, :, :, :, :, 2:3]), 5)
    combine = rots.matmul(torch.inverse(intrins))
    points = combine.view(B, N, 1, 1, 1, 3, 3).matmul(points).squeeze(-1)
    return points + trans.view(B, N, 1, 1, 1, 3)


def lift(depth, x_img):
    """depth [BN, D, fH, fW] (softmax over D), x_img [BN, C, fH, fW] -> [BN, C, D, fH, fW]"""
    return depth.unsqueeze(1) * x_img.unsqueeze(2)


def voxel_cells(geom, dx, bx, nx):
    """integer cell of every frustum point and the `kept` mask (airv2x_encoder.py:226-246); .long() truncates toward 0"""
    cells = ((geom - (bx - dx / 2.0)) / dx).long().view(-1, 3)
    kept = ((cells[:, 0] >= 0) & (cells[:, 0] < nx[0]) & (cells[:, 1] >= 0) & (cells[:, 1] < nx[1])
            & (cells[:, 2] >= 0) & (cells[:, 2] < nx[2]))
    return cells, kept


def voxel_pooling(geom, x, dx, bx, nx):
    """geom [B,N,D,H,W,3], x [B,N,D,H,W,C] -> [B, C*nz, ny, nx] with the reference's cumulative-sum pooling"""
    B, N, D, H, W, C = x.shape
    Nprime = B * N * D * H * W
    x = x.reshape(Nprime, C)
    cells, kept = voxel_cells(geom, dx, bx, nx)
    batch_ix = torch.cat([torch.full([Nprime // B, 1], ix, dtype=torch.long) for ix in range(B)])
    g = torch.cat((cells, batch_ix), 1)
    x, g = x[kept], g[kept]
    ranks = g[:, 0] * (nx[1] * nx[2] * B) + g[:, 1] * (nx[2] * B) + g[:, 2] * B + g[:, 3]
    sorts = ranks.argsort()
    x, g, ranks = x[sorts], g[sorts], ranks[sorts]
    x = x.cumsum(0)
    last = torch.ones(x.shape[0], dtype=torch.bool)
    last[:-1] = ranks[1:] != ranks[:-1]
    x, g = x[last], g[last]
    x = torch.cat((x[:1], x[1:] - x[:-1]))
    final = torch.zeros((B, C, int(nx[2]), int(nx[1]), int(nx[0])))
    final[g[:, 3], :, g[:, 2], g[:, 1], g[:, 0]] = x
    return torch.cat(final.unbind(dim=2), 1)


def voxel_pooling_exact(geom, x, dx, bx, nx):
    """the same pooling with every cell summed in float64 (what the cumulative-sum trick approximates)"""
    B, N, D, H, W, C = x.shape
    cells, kept = voxel_cells(geom, dx, bx, nx)
    batch_ix = torch.arange(B).view(B, 1).expand(B, N * D * H * W).reshape(-1)
    xf = x.reshape(-1, C).double()[kept]
    c, b = cells[kept], batch_ix[kept]
    flat = ((b * int(nx[2]) + c[:, 2]) * int(nx[1]) + c[:, 1]) * int(nx[0]) + c[:, 0]
    out = torch.zeros(B * int(nx[2]) * int(nx[1]) * int(nx[0]), C, dtype=torch.float64)
    out.index_add_(0, flat, xf)
    out = out.view(B, int(nx[2]), int(nx[1]), int(nx[0]), C).permute(0, 4, 1, 2, 3)
    return torch.cat(out.unbind(dim=2), 1)


def synth_cameras(B, N, seed, final_dim):
    """plausible pinhole rigs: N cameras around the vehicle looking outward, small augmentation transforms"""
    g = torch.Generator().manual_seed(seed)
    H, W = final_dim
    rots, trans, intr, prots, ptrans = [], [], [], [], []
    for b in range(B):
        for n in range(N):
            yaw = 2 * np.pi * n / N + float(torch.rand(1, generator=g)) * 0.2
            # camera frame (x right, y down, z forward) -> ego frame (x forward, y left, z up), then yaw about z
            base = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
            cy, sy = np.cos(yaw), np.sin(yaw)
            rz = torch.tensor([[cy, -sy, 0.0], [sy, cy, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32)
            rots.append(rz @ base)
            trans.append(torch.tensor([1.5 * cy, 1.5 * sy, 1.6]) + 0.1 * torch.randn(3, generator=g))
            f = 0.8 * W
            intr.append(torch.tensor([[f, 0.0, W / 2], [0.0, f, H / 2], [0.0, 0.0, 1.0]]))
            s = 1.0 + 0.05 * float(torch.randn(1, generator=g))
            prots.append(torch.tensor([[s, 0.0, 0.0], [0.0, s, 0.0], [0.0, 0.0, 1.0]]))
            ptrans.append(torch.tensor([float(torch.randn(1, generator=g)) * 4, float(torch.randn(1, generator=g)) * 3, 0.0]))
    st = lambda l, *shape: torch.stack(l).view(B, N, *shape).float()
    return st(rots, 3, 3), st(trans, 3), st(intr, 3, 3), st(prots, 3, 3), st(ptrans, 3)
