import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import a2x_import
ops = a2x_import.pkg("ops")
g = torch.Generator().manual_seed(0)
n, h, w, cin, cout = 5, 100, 352, 64, 64
x = torch.randn(n, h, w, cin, generator=g).cuda(); wt = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05).cuda()
xs = ops.split(x); pw = ops.pack_conv_weight(wt); y = ops.Act.empty((n, h, w, cout), "cuda", True)
for _ in range(3): ops.conv_fwd(xs, pw, 3, 1, y)
torch.cuda.synchronize()
