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


# ---------------------------------------------------------------------------------------------------------------------
# BevEncode: the BEV encoder that consumes the Lift-Splat pooling (sub_modules/lss_submodule.py:312-349): 7x7 stride-2 stem,
# ResNet-18 layer1-3 (torchvision BasicBlock), Up (bilinear x4 + skip concat + 2 x conv3x3-BN-ReLU), up2 (bilinear x2,
# conv3x3-BN-ReLU, conv1x1). Parameter containers with the reference's state_dict keys; the eval forward runs on the
# tap-GEMM kernels (BatchNorm folded into the GEMM epilogue, the residual add of a BasicBlock in the epilogue of its second
# conv, the 7x7 stem as 49 taps over the stride-2 parity views), bilinear resampling on the ego-warp kernel.
class _BasicBlockParams(torch.nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        nn = torch.nn
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.stride = stride
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


class BevEncode(torch.nn.Module):
    """Same constructor, state_dict keys and eval forward as the reference's BevEncode(inC, outC); input / output NCHW.
    inC a multiple of 64, H and W multiples of 8. Train mode is not implemented (the oracle's train mode is pinned for it)."""

    def __init__(self, inC, outC, precision="split3"):
        super().__init__()
        nn = torch.nn
        assert inC % 64 == 0, "BevEncode (B200): inC must be a multiple of 64 (split-precision GEMM operand)"
        self.conv1 = nn.Conv2d(inC, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layer1 = nn.Sequential(_BasicBlockParams(64, 64, 1), _BasicBlockParams(64, 64, 1))
        self.layer2 = nn.Sequential(_BasicBlockParams(64, 128, 2), _BasicBlockParams(128, 128, 1))
        self.layer3 = nn.Sequential(_BasicBlockParams(128, 256, 2), _BasicBlockParams(256, 256, 1))
        self.up1 = nn.Module()
        self.up1.conv = nn.Sequential(nn.Conv2d(64 + 256, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                                      nn.Conv2d(256, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True))
        self.up2 = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
                                 nn.Conv2d(256, 128, 3, padding=1, bias=False), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
                                 nn.Conv2d(128, outC, 1, padding=0))
        self.outC = outC
        assert precision == "split3"

    # ---- kernel plumbing
    @staticmethod
    def _affine(bn):
        scale, shift = torch.empty_like(bn.weight), torch.empty_like(bn.weight)
        ops.bn_eval_affine(bn.weight, bn.bias, bn.running_mean, bn.running_var, scale, shift, eps=bn.eps)
        return scale, shift

    def _conv_bn(self, x, conv, bn, relu, out=None, accumulate=False):
        k, s = conv.kernel_size[0], conv.stride[0]
        n, h, w, _ = x.shape
        ho, wo = (h - 1) // s + 1, (w - 1) // s + 1
        if out is None:
            out = ops.Act.empty((n, ho, wo, conv.out_channels), x.hi.device, True)
        scale, shift = self._affine(bn)
        ops.conv_fwd(x, ops.pack_conv_weight(conv.weight.detach()), k, s, out, scale=scale, shift=shift, relu=relu,
                     accumulate=accumulate)
        return out

    def _block(self, blk, x):
        """BasicBlock: relu(bn2(conv2(relu(bn1(conv1(x))))) + identity), the add + ReLU in the second GEMM's epilogue"""
        h = self._conv_bn(x, blk.conv1, blk.bn1, True)
        out = ops.Act.empty(h.shape, x.hi.device, True)
        if hasattr(blk, "downsample"):
            sub = ops.split(x.hi[:, ::blk.stride, ::blk.stride, :].contiguous())      # 1x1 stride-s conv = 1x1 on the subsampled map
            conv, bn = blk.downsample[0], blk.downsample[1]
            scale, shift = self._affine(bn)
            ops.conv_fwd(sub, ops.pack_conv_weight(conv.weight.detach()), 1, 1, out, scale=scale, shift=shift, relu=False)
        else:
            out.hi.copy_(x.hi)
        return self._conv_bn(h, blk.conv2, blk.bn2, True, out=out, accumulate=True)

    @staticmethod
    def _upsample(src, scale, out):
        """nn.Upsample(bilinear, align_corners=True) = identity affine resampling on the ego-warp kernel"""
        n = src.shape[0]
        theta = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], device=src.device).repeat(n, 1, 1)
        ops.warp_affine_fwd(src, theta, out, align_corners=True)
        return out

    def forward(self, x):
        if x.device.type != "cuda":
            raise RuntimeError("BevEncode (B200) needs CUDA tensors; there is no CPU path")
        if self.training:
            raise NotImplementedError("BevEncode (B200): eval forward only")
        B, C, H, W = x.shape
        assert H % 8 == 0 and W % 8 == 0, "BevEncode: H and W must be multiples of 8"
        dev = x.device
        with torch.no_grad():
            xa = ops.split(x.permute(0, 2, 3, 1).contiguous().float())
            # stem: raw 7x7 stride-2 conv, then BatchNorm + ReLU + operand split in one pass
            z = torch.empty(B, H // 2, W // 2, 64, device=dev)
            ops.conv_fwd(xa, ops.pack_conv_weight(self.conv1.weight.detach()), 7, 2, ops.Act(z))
            scale, shift = self._affine(self.bn1)
            y = ops.Act.empty(z.shape, dev, True)
            ops.affine_act(z, scale, shift, True, y)
            x1 = y
            for blk in self.layer1:
                x1 = self._block(blk, x1)
            t = x1
            for layer in (self.layer2, self.layer3):
                for blk in layer:
                    t = self._block(blk, t)
            # Up: cat([skip x1 (64), bilinear x4 of t (256)]) -> 2 x conv3x3-BN-ReLU
            cat = ops.Act.empty((B, H // 2, W // 2, 64 + 256), dev, True)
            ops.affine_act(x1.hi, None, None, False, cat.slice_c(0, 64))
            self._upsample(t.hi, 4, cat.slice_c(64, 320))
            u = self._conv_bn(cat, self.up1.conv[0], self.up1.conv[1], True)
            u = self._conv_bn(u, self.up1.conv[3], self.up1.conv[4], True)
            # up2: bilinear x2, conv3x3-BN-ReLU, conv1x1 + bias (output columns padded to a multiple of 32)
            v = self._upsample(u.hi, 2, ops.Act.empty((B, H, W, 256), dev, True))
            v = self._conv_bn(v, self.up2[1], self.up2[2], True)
            last = self.up2[4]
            cpad = (self.outC + 31) // 32 * 32
            bias = torch.zeros(cpad, device=dev)
            bias[:self.outC] = last.bias.detach()
            o = ops.Act(torch.empty(B, H, W, cpad, device=dev))
            ops.conv_fwd(v, ops.pack_conv_weight(last.weight.detach(), cout_pad=cpad), 1, 1, o, shift=bias)
            return o.hi[..., :self.outC].permute(0, 3, 1, 2)


class CameraBranch(torch.nn.Module):
    """LiftSplatShootEncoder (common_modules/airv2x_encoder.py:30-335) with the image trunk as a pluggable library call:
    imgs -> trunk -> (depth distribution [B*N, D, fH, fW], image features [B*N, C, fH, fW]) -> lift + voxel pooling
    (a2x_lift_splat_fwd) -> BevEncode (tap-GEMM kernels) -> {"spatial_features": [B, bevout, ny, nx]}.
    `trunk` is any callable / nn.Module with CamEncode's contract (sub_modules/lss_submodule.py:50-190): the reference's is
    EfficientNet-b0 with downloaded weights, not available offline, so it is not built here and its parity is unpinned.
    args: the yaml's camera block (grid_conf, data_aug_conf.final_dim, img_downsample, img_features, bevout_feature)."""

    def __init__(self, args, agent_type, trunk, device="cuda"):
        super().__init__()
        self.agent_type = agent_type
        self.trunk = trunk
        self.ls = LiftSplat(args["grid_conf"], args["data_aug_conf"]["final_dim"], args["img_downsample"], device)
        self.bevencode = BevEncode(args["img_features"], args["bevout_feature"])
        assert self.ls.nx[2] == 1, "one z bin: the pooled map has img_features channels (BevEncode's inC)"

    def forward(self, data_dict):
        cam = data_dict[self.agent_type]["batch_merged_cam_inputs"]
        imgs = cam["imgs"]
        depth, feat = self.trunk(imgs)
        geom = self.ls.geometry(cam["rots"], cam["trans"], cam["intrinsics"], cam["post_rots"], cam["post_trans"])
        bev = self.ls(depth, feat, geom)
        x = self.bevencode(bev)
        return {"spatial_features": x, "spatial_features_3d": x.unsqueeze(2)}
