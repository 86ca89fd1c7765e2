import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import a2x_import
ops = a2x_import.pkg("ops"); lib = a2x_import.pkg("_lib").load()
torch.backends.cudnn.allow_tf32 = False
g = torch.Generator().manual_seed(0)
def rnd(*s): return torch.randn(*s, generator=g).cuda()
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
def rel(a, b): return float((a - b).abs().max() / (b.abs().max() + 1e-12))
for mode in (0, 1):
    lib.a2x_debug_set(6, mode)
    for (n, h, w, cin, cout) in [(1, 16, 8, 64, 64), (2, 20, 44, 64, 64), (1, 25, 88, 128, 256), (2, 12, 40, 256, 128)]:
        x, wt = rnd(n, cin, h, w), rnd(cout, cin, 3, 3) * 0.1
        yref = F.conv2d(x, wt, padding=1); dy = rnd(*yref.shape)
        dxref = torch.nn.grad.conv2d_input(x.shape, wt, dy, padding=1)
        pw = ops.pack_conv_weight(wt)
        for split in (False, True):
            xs = ops.split(nhwc(x)) if split else ops.Act(nhwc(x))
            dys = ops.split(nhwc(dy)) if split else ops.Act(nhwc(dy))
            y = ops.Act(torch.empty_like(nhwc(yref))); ops.conv_fwd(xs, pw, 3, 1, y)
            dx = torch.empty_like(nhwc(x)); ops.conv_dgrad(dys, pw, 3, 1, dx)
            torch.cuda.synchronize()
            print("base_offset_mode=%d %s split=%d fwd %.2e dgrad %.2e" % (mode, (n, h, w, cin, cout), split, rel(y.hi, nhwc(yref)), rel(dx, nhwc(dxref))), flush=True)
lib.a2x_debug_set(6, 0)
# timing at full size, halo vs classic
for (n, h, w, cin, cout) in [(5, 100, 352, 64, 64), (5, 50, 176, 128, 128), (5, 25, 88, 256, 256), (5, 100, 352, 256, 256)]:
    x, wt = rnd(n, cin, h, w), rnd(cout, cin, 3, 3) * 0.05
    xs = ops.split(nhwc(x)); pw = ops.pack_conv_weight(wt); y = ops.Act(torch.empty(n, h, w, cout, device="cuda"))
    for dbg, nost in ((0, 0), (0, 1)):
        lib.a2x_debug_set(7, dbg); lib.a2x_debug_set(9, nost)
        ops.conv_fwd(xs, pw, 3, 1, y); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20): ops.conv_fwd(xs, pw, 3, 1, y)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        print("%s %s: %.3f ms  %.1f alg TFLOP/s" % ("halo no-store" if nost else "halo         ", (n, h, w, cin, cout), ms, 2 * n * h * w * 9 * cin * cout / ms / 1e9), flush=True)
lib.a2x_debug_set(7, 0); lib.a2x_debug_set(9, 0)
