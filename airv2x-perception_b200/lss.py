"""Lift-Splat camera branch on the GPU (the camera half of BASELINE config 5): frustum geometry + the fused
lift / voxel-pooling kernel behind `LiftSplatShootEncoder.get_geometry` / `voxel_pooling`
(opencood/models/common_modules/airv2x_encoder.py:94-275) and the lift of `CamEncode.forward`
(opencood/models/sub_modules/lss_submodule.py:170-186).

    ls = LiftSplat(grid_conf, final_dim, img_downsample, device)
    geom = ls.geometry(rots, trans, intrins, post_rots, post_trans)         # [B, N, D, fH, fW, 3]
    bev = ls(depth, x_img, geom)                                            # [B, C * nz, ny, nx], differentiable

`depth` [B*N, D, fH, fW] is the softmax depth distribution and `x_img` [B*N, C, fH, fW] the image features — the two
heads of the camera trunk (EfficientNet: a library call, its pretrained weights are not available offline, so the trunk
is outside this module and its parity unpinned; lift + splat are pinned to the real reference, oracle/lss_oracle.py).
The [B, N, D, fH, fW, C] product the reference materialises, sorts and cumulative-sums is never formed. No CPU fallback.
"""
import numpy as np
import torch

from . import ops


def gen_dx_bx(xbound, ybound, zbound):
    """utils/camera_utils.py:238-244 (fp32 tensors like the reference's)"""
    rows = [xbound, ybound, zbound]
    dx = torch.tensor([r[2] for r in rows], dtype=torch.float32)
    bx = torch.tensor([r[0] + r[2] / 2.0 for r in rows], dtype=torch.float32)
    nx = [int((r[1] - r[0]) / r[2] + 0.5) for r in rows]
    return dx, bx, nx


def depth_bins(depth_min, depth_max, num_bins, mode):
    """utils/camera_utils.py:310-326 (depth_discretization)"""
    if mode == "UD":
        return np.linspace(depth_min, depth_max, num_bins, endpoint=False)
    if mode == "LID":
        idx = np.arange(0, num_bins)
        bin_size = 2 * (depth_max - depth_min) / (num_bins * (1 + num_bins))
        return depth_min + bin_size * (idx * (idx + 1)) / 2
    raise NotImplementedError("depth discretisation mode %r" % mode)


class _LiftSplatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, feat, geom, owner):
        B, N = geom.shape[0], geom.shape[1]
        nx, ny, nz = owner.nx
        C = feat.shape[1]
        depth, feat = depth.contiguous(), feat.contiguous()
        bev = torch.empty(B, ny, nx, nz * C, device=depth.device)
        cells = torch.empty(depth.numel(), dtype=torch.int32, device=depth.device)
        ops.lift_splat_fwd(depth, feat, geom.contiguous(), B, N, owner.origin, owner.dx, owner.nx, bev, cells)
        ctx.save_for_backward(depth, feat, cells)
        ctx.bn = (B, N)
        return bev.permute(0, 3, 1, 2)                      # the reference's [B, C * nz, ny, nx] (a view)

    @staticmethod
    def backward(ctx, dbev):
        depth, feat, cells = ctx.saved_tensors
        B, N = ctx.bn
        ddepth, dfeat = torch.empty_like(depth), torch.empty_like(feat)
        ops.lift_splat_bwd(depth, feat, cells, dbev.permute(0, 2, 3, 1).contiguous(), B, N, ddepth, dfeat)
        return ddepth, dfeat, None, None


class LiftSplat:
    def __init__(self, grid_conf, final_dim, downsample, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("LiftSplat (B200) needs a CUDA device; there is no CPU path")
        dx, bx, nx = gen_dx_bx(grid_conf["xbound"], grid_conf["ybound"], grid_conf["zbound"])
        self.dx, self.nx = [float(v) for v in dx], nx
        self.origin = [float(v) for v in (bx - dx / 2.0)]   # fp32 tensor arithmetic, as airv2x_encoder.py:226 evaluates it
        ogfH, ogfW = final_dim
        fH, fW = ogfH // downsample, ogfW // downsample
        ds = torch.tensor(depth_bins(*grid_conf["ddiscr"], grid_conf["mode"]), dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
        D = ds.shape[0]
        xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
        ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
        self.frustum = torch.stack((xs, ys, ds), -1).to(self.device)    # [D, fH, fW, 3]
        self.D, self.fH, self.fW = D, fH, fW

    def geometry(self, rots, trans, intrins, post_rots, post_trans):
        """frustum points in the ego frame, [B, N, D, fH, fW, 3] (airv2x_encoder.py:133-168), torch ops on the device"""
        dev = self.device
        rots, trans, intrins, post_rots, post_trans = [t.to(dev).float() for t in (rots, trans, intrins, post_rots, post_trans)]
        B, N, _ = trans.shape
        pts = self.frustum - post_trans.view(B, N, 1, 1, 1, 3)
        pts = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(pts.unsqueeze(-1))
        pts = torch.cat((pts[..., :2, :] * pts[..., 2:3, :], pts[..., 2:3, :]), 5)
        combine = rots.matmul(torch.inverse(intrins))
        pts = combine.view(B, N, 1, 1, 1, 3, 3).matmul(pts).squeeze(-1)
        return (pts + trans.view(B, N, 1, 1, 1, 3)).contiguous()

    def __call__(self, depth, feat, geom):
        assert depth.is_cuda and feat.is_cuda and geom.is_cuda, "LiftSplat inputs must live on the GPU"
        assert tuple(depth.shape[1:]) == (self.D, self.fH, self.fW) and tuple(feat.shape[2:]) == (self.fH, self.fW)
        assert tuple(geom.shape[2:]) == (self.D, self.fH, self.fW, 3) and geom.shape[0] * geom.shape[1] == depth.shape[0]
        return _LiftSplatFn.apply(depth, feat, geom, self)
