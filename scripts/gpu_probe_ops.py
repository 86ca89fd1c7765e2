"""GPU probe: elementwise / fusion / BN-backward ops vs torch autograd. Report only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.nn.functional as F
import a2x_import
ops = a2x_import.pkg("ops"); Act = ops.Act
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(0)
def rnd(*s): return torch.randn(*s, generator=g).cuda()
def rel(a, b): return float((a - b).abs().max() / (b.abs().max() + 1e-12))
# attention fusion fwd/bwd
for n, C, H, W in [(4, 64, 9, 13), (3, 128, 5, 7), (5, 256, 4, 6), (1, 64, 3, 5)]:
    x = rnd(n, H, W, C).requires_grad_(True)
    q = x.view(n, H * W, C).permute(1, 0, 2)
    ctx = torch.bmm(F.softmax(torch.bmm(q, q.transpose(1, 2)) / np.sqrt(C), -1), q)
    ref = ctx[:, 0].view(1, H, W, C)
    dout = rnd(1, H, W, C)
    ref.backward(dout)
    out = Act(torch.empty(1, H, W, C, device="cuda"))
    ops.att_fuse_fwd(x.detach(), out)
    dx = torch.empty_like(x)
    ops.att_fuse_bwd(x.detach(), dout, dx)
    torch.cuda.synchronize()
    print("att n=%d C=%d fwd rel %.2e  bwd rel %.2e" % (n, C, rel(out.hi, ref.detach()), rel(dx, x.grad)))
# BN + ReLU train fwd/bwd
for N, H, W, C in [(3, 10, 12, 64), (2, 7, 9, 128)]:
    z = rnd(N, H, W, C).requires_grad_(True)
    gam = (torch.rand(C, generator=g) + 0.5).cuda().requires_grad_(True); bet = (torch.randn(C, generator=g) * 0.1).cuda().requires_grad_(True)
    y = F.relu(F.batch_norm(z.permute(0, 3, 1, 2), None, None, gam, bet, True, 0.01, 1e-3)).permute(0, 2, 3, 1)
    dy = rnd(N, H, W, C)
    y.backward(dy)
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    scale, shift, mean, invstd = [torch.empty(C, device="cuda") for _ in range(4)]
    ops.channel_stats(z.detach(), sums)
    ops.bn_finalize(sums, N * H * W, gam.detach(), bet.detach(), 0, None, None, scale, shift, mean, invstd)
    yo = Act(torch.empty(N, H, W, C, device="cuda"))
    ops.affine_act(z.detach(), scale, shift, True, yo)
    bs = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    dz = Act(torch.empty(N, H, W, C, device="cuda")); dg = torch.empty(C, device="cuda"); db = torch.empty(C, device="cuda")
    ops.bn_relu_bwd(dy, z.detach(), scale, shift, mean, invstd, bs, dz, dg, db)
    torch.cuda.synchronize()
    print("bn C=%d fwd %.2e dz %.2e dgamma %.2e dbeta %.2e" % (C, rel(yo.hi, y.detach()), rel(dz.hi, z.grad), rel(dg, gam.grad), rel(db, bet.grad)))
# split conv family in 3x mode vs fp32 torch (tight tolerance)
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
for (n, h, w, cin, cout, k, s) in [(2, 12, 40, 64, 64, 3, 1), (2, 20, 44, 64, 128, 3, 2), (1, 10, 36, 384, 256, 1, 1)]:
    x = rnd(n, cin, h, w); wt = rnd(cout, cin, k, k) * 0.1
    yref = F.conv2d(x, wt, stride=s, padding=k // 2); dy = rnd(*yref.shape)
    dxref = torch.nn.grad.conv2d_input(x.shape, wt, dy, stride=s, padding=k // 2)
    dwref = torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride=s, padding=k // 2)
    wf, wd = ops.pack_conv_weight(wt)
    xs, dys = ops.split_tf32(nhwc(x)), ops.split_tf32(nhwc(dy))
    y = Act(torch.empty_like(nhwc(yref)))
    ops.conv_fwd(xs, wf, k, s, y)
    dx = torch.empty_like(nhwc(x)); ops.conv_dgrad(dys, wd, k, s, dx)
    dwp = torch.zeros(k * k, cout, cin, device="cuda"); ops.conv_wgrad(xs, dys, k, s, dwp)
    dw = ops.unpack_conv_wgrad(dwp, cout, cin, k)
    torch.cuda.synchronize()
    print("3xTF32 conv k%d s%d %d->%d: fwd %.2e dgrad %.2e wgrad %.2e" % (k, s, cin, cout, rel(y.hi, nhwc(yref)), rel(dx, nhwc(dxref)), rel(dw, dwref)))
for (n, h, w, cin, cout, s) in [(1, 10, 18, 128, 128, 2), (2, 5, 9, 256, 128, 4), (1, 6, 8, 64, 128, 1)]:
    x = rnd(n, cin, h, w).requires_grad_(True); wt = (rnd(cin, cout, s, s) * 0.1).requires_grad_(True)
    yref = F.conv_transpose2d(x, wt, stride=s); dy = rnd(*yref.shape); yref.backward(dy)
    wf, wd = ops.pack_deconv_weight(wt.detach())
    xs, dys = ops.split_tf32(nhwc(x.detach())), ops.split_tf32(nhwc(dy))
    y = Act(torch.empty_like(nhwc(yref.detach()))); ops.deconv_fwd(xs, wf, cout, s, y)
    dx = torch.empty_like(nhwc(x.detach())); ops.deconv_dgrad(dys, wd, s, dx)
    dwp = torch.zeros(s * s, cin, cout, device="cuda"); ops.deconv_wgrad(xs, dys, s, dwp)
    dw = ops.unpack_deconv_wgrad(dwp, cin, cout, s)
    torch.cuda.synchronize()
    print("3xTF32 deconv s%d %d->%d: fwd %.2e dgrad %.2e wgrad %.2e" % (s, cin, cout, rel(y.hi, nhwc(yref.detach())), rel(dx, nhwc(x.grad)), rel(dw, wt.grad)))
