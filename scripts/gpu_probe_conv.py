"""GPU probe: every dense-contraction entry point vs torch (fp32, TF32 off). Prints a report, never raises."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import a2x_import
ops = a2x_import.pkg("ops")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)

def rnd(*s):
    return torch.randn(*s, generator=g).to(dev)

def nhwc(t):  # NCHW -> NHWC contiguous
    return t.permute(0, 2, 3, 1).contiguous()

def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))

def report(name, a, b):
    r = rel(a, b)
    print("%-40s rel_err=%.3e  %s" % (name, r, "OK" if r < 5e-3 else "FAIL"), flush=True)

def conv_case(n, h, w, cin, cout, k, s):
    x = rnd(n, cin, h, w); wt = rnd(cout, cin, k, k) * 0.1
    xh = nhwc(x)
    wf, wd = ops.pack_conv_weight(wt)
    tag = "n%d %dx%d %d->%d k%d s%d" % (n, h, w, cin, cout, k, s)
    yref = F.conv2d(x, wt, stride=s, padding=k // 2)
    try:
        y = ops.conv2d_fwd(xh, wf, k, s); torch.cuda.synchronize()
        report("conv_fwd " + tag, y, nhwc(yref))
    except Exception as e:
        print("conv_fwd", tag, "EXC", e)
    dy = rnd(*yref.shape)
    dxref = torch.nn.grad.conv2d_input(x.shape, wt, dy, stride=s, padding=k // 2)
    dwref = torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride=s, padding=k // 2)
    try:
        dx = ops.conv2d_dgrad(nhwc(dy), wd, k, s, h, w); torch.cuda.synchronize()
        report("conv_dgrad " + tag, dx, nhwc(dxref))
    except Exception as e:
        print("conv_dgrad", tag, "EXC", e)
    try:
        dwp = ops.conv2d_wgrad(xh, nhwc(dy), k, s); torch.cuda.synchronize()
        dw = ops.unpack_conv_wgrad(dwp, cout, cin, k); torch.cuda.synchronize()
        report("conv_wgrad " + tag, dw, dwref)
    except Exception as e:
        print("conv_wgrad", tag, "EXC", e); traceback.print_exc()

def deconv_case(n, h, w, cin, cout, s):
    x = rnd(n, cin, h, w); wt = rnd(cin, cout, s, s) * 0.1
    wf, wd = ops.pack_deconv_weight(wt)
    tag = "n%d %dx%d %d->%d s%d" % (n, h, w, cin, cout, s)
    xr = x.clone().requires_grad_(True); wr = wt.clone().requires_grad_(True)
    yref = F.conv_transpose2d(xr, wr, stride=s)
    dy = rnd(*yref.shape)
    yref.backward(dy)
    try:
        y = ops.deconv_fwd(nhwc(x), wf, cout, s); torch.cuda.synchronize()
        report("deconv_fwd " + tag, y, nhwc(yref.detach()))
    except Exception as e:
        print("deconv_fwd", tag, "EXC", e)
    try:
        dx = ops.deconv_dgrad(nhwc(dy), wd, s); torch.cuda.synchronize()
        report("deconv_dgrad " + tag, dx, nhwc(xr.grad))
    except Exception as e:
        print("deconv_dgrad", tag, "EXC", e)
    try:
        dwp = ops.deconv_wgrad(nhwc(x), nhwc(dy), s); torch.cuda.synchronize()
        dw = ops.unpack_deconv_wgrad(dwp, cin, cout, s); torch.cuda.synchronize()
        report("deconv_wgrad " + tag, dw, wr.grad)
    except Exception as e:
        print("deconv_wgrad", tag, "EXC", e)

print(torch.cuda.get_device_name(0), flush=True)
conv_case(1, 8, 16, 32, 32, 1, 1)
conv_case(2, 12, 40, 64, 64, 3, 1)
conv_case(2, 20, 44, 64, 64, 3, 2)
conv_case(1, 25, 88, 128, 256, 3, 1)
conv_case(2, 9, 21, 64, 128, 3, 2)
conv_case(1, 10, 36, 384, 256, 1, 1)
conv_case(1, 10, 36, 256, 32, 1, 1)
deconv_case(1, 10, 36, 64, 128, 1)
deconv_case(1, 10, 18, 128, 128, 2)
deconv_case(2, 5, 9, 256, 128, 4)
# channel-slice output / input (concat buffer)
buf = torch.zeros(1, 10, 36, 384, device=dev)
x = rnd(1, 64, 10, 36); wt = rnd(64, 128, 1, 1) * 0.1
wf, wd = ops.pack_deconv_weight(wt)
ops.deconv_fwd(nhwc(x), wf, 128, 1, out=buf[..., 128:256]); torch.cuda.synchronize()
report("deconv_fwd into channel slice", buf[..., 128:256], nhwc(F.conv_transpose2d(x, wt)))
print("other channels untouched:", float(buf[..., :128].abs().max()), float(buf[..., 256:].abs().max()))
# full-size timing of the level-0 conv
x = rnd(5, 64, 100, 352); wt = rnd(64, 64, 3, 3) * 0.05
xh = nhwc(x); wf, wd = ops.pack_conv_weight(wt)
y = ops.conv2d_fwd(xh, wf, 3, 1)
torch.cuda.synchronize()
for name, fn in [("conv64 L0", lambda: ops.conv2d_fwd(xh, wf, 3, 1, out=y))]:
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    fl = 2 * 5 * 100 * 352 * 576 * 64
    print("%s: %.3f ms  %.1f TFLOP/s" % (name, ms, fl / ms / 1e9))
report("conv_fwd full L0", y, nhwc(F.conv2d(x, wt, padding=1)))
x = rnd(5, 256, 25, 88); wt = rnd(256, 256, 3, 3) * 0.02
xh = nhwc(x); wf, wd = ops.pack_conv_weight(wt)
y = ops.conv2d_fwd(xh, wf, 3, 1); torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): ops.conv2d_fwd(xh, wf, 3, 1, out=y)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 20
print("conv256 L2: %.3f ms  %.1f TFLOP/s" % (ms, 2 * 5 * 25 * 88 * 2304 * 256 / ms / 1e9))
report("conv_fwd full L2", y, nhwc(F.conv2d(x, wt, padding=1)))
