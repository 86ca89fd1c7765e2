import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes
import a2x_import
ops = a2x_import.pkg("ops"); lib = a2x_import.pkg("_lib").load()
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
g = torch.Generator().manual_seed(0)
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
n, h, w, cin, cout = 1, 4, 32, 32, 32
x = torch.randn(n, cin, h, w, generator=g).to(dev); dy = torch.randn(n, cout, h, w, generator=g).to(dev)
ref = torch.einsum("nohw,nihw->oi", dy, x)
for dbg in [(0, 0), (4096, 512), (512, 4096), (4096, 1024), (128, 512)]:
    lib.a2x_debug_set(2, dbg[0]); lib.a2x_debug_set(3, dbg[1])
    dwp = ops.conv2d_wgrad(nhwc(x), nhwc(dy), 1, 1); torch.cuda.synchronize()
    d = dwp[0]
    print("lbo/sbo", dbg, "max|d|=%.4f" % float(d.abs().max()), "err=%.4e" % float((d - ref).abs().max()), "ref max %.3f" % float(ref.abs().max()), "errT %.4e" % float((d - ref.t()).abs().max()))
lib.a2x_debug_set(2, 0); lib.a2x_debug_set(3, 0)
n, h, w, cin, cout = 2, 12, 40, 128, 256
x = torch.randn(n, cin, h, w, generator=g).to(dev); dy = torch.randn(n, cout, h, w, generator=g).to(dev)
ref = torch.nn.grad.conv2d_weight(x, (cout, cin, 3, 3), dy, padding=1)
dwp = ops.conv2d_wgrad(nhwc(x), nhwc(dy), 3, 1); dw = ops.unpack_conv_wgrad(dwp, cout, cin, 3); torch.cuda.synchronize()
print("3x3 128->256 rel err", float((dw - ref).abs().max() / ref.abs().max()))
