"""Full-size launches of the HBM-bound kernels that the headline step does not run or runs only small (for `ncu --set full`
captures, profiles/): ego-warp fwd / bwd (100 x 352 x 256 maps, 4 non-ego agents), communication-mask compaction +
pointer-table decompaction (5 agents x 100 x 352 x 64, 30 % of the cells selected), AttentionFusion fwd / bwd at level 0.

    ncu --set full --clock-control none -k regex:"warp_affine|mask_compact|mask_decompact|att_fuse" -o out python scripts/ncu_small_kernels.py
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import a2x_import


def main():
    ops = a2x_import.pkg("ops")
    g = torch.Generator().manual_seed(0)
    n, H, W, C = 4, 100, 352, 256
    x = torch.randn(n, H, W, C, generator=g).cuda()
    th = torch.zeros(n, 2, 3)
    for i in range(n):
        a = 0.1 * (i + 1)
        th[i] = torch.tensor([[math.cos(a), -math.sin(a), 0.05 * i], [math.sin(a), math.cos(a), -0.03 * i]])
    th = th.cuda()
    out = ops.Act.empty(x.shape, "cuda", True)
    dsrc = torch.zeros_like(x)
    for _ in range(3):
        ops.warp_affine_fwd(x, th, out, align_corners=True)
        dsrc.zero_()
        ops.warp_affine_bwd(x, th, dsrc, align_corners=True)
    # sparse feature select
    n, C = 5, 64
    hw = H * W
    x = torch.randn(n, H, W, C, generator=g).cuda()
    mask = (torch.rand(n, H, W, generator=g) > 0.7).float().cuda()
    total = 64 + (hw + 3) // 4 * 4 + hw * C
    bufs = torch.zeros(n, total, device="cuda")
    table = torch.tensor([bufs[a].data_ptr() for a in range(n)], dtype=torch.int64, device="cuda")
    dst = torch.empty(n, H, W, C, device="cuda")
    for _ in range(3):
        for a in range(n):
            ops.mask_compact(x[a:a + 1], mask[a], a == 0, bufs[a, :64].view(torch.int32),
                             bufs[a, 64:64 + (hw + 3) // 4 * 4].view(torch.int32), bufs[a, 64 + (hw + 3) // 4 * 4:])
        ops.mask_decompact_ptrs(table, 64 * 4, (64 + (hw + 3) // 4 * 4) * 4, n, dst)
    # per-pixel attention fusion over the agents, level 0
    fused = ops.Act.empty((1, H, W, C), "cuda", True)
    dfused = torch.randn(1, H, W, C, generator=g).cuda()
    dx = torch.empty_like(x)
    for _ in range(3):
        ops.att_fuse_fwd(x, fused)
        ops.att_fuse_bwd(x, dfused, dx)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
