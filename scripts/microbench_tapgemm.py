"""Micro-benchmark of the non-halo tap-GEMM on the shapes of the headline step where it is far from the tensor peak
(1x1 shrink conv, stride-2 convs, transposed convs): time with / without the epilogue stores (relu = 77 debug switch) and
with / without the fp32 output plane, to separate MMA / operand-fetch time from epilogue time."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import a2x_import


def main():
    ops = a2x_import.pkg("ops")
    g = torch.Generator().manual_seed(0)
    res = []

    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for (n, h, w, cin, cout, k, s) in ((5, 100, 352, 384, 256, 1, 1), (5, 200, 704, 64, 64, 3, 2), (5, 100, 352, 64, 128, 3, 2),
                                       (5, 100, 352, 256, 256, 3, 1), (5, 25, 88, 256, 256, 3, 1)):
        x = ops.split(torch.randn(n, h, w, cin, generator=g).cuda())
        wt = torch.randn(cout, cin, k, k, generator=g).cuda() * 0.05
        pk = ops.pack_conv_weight(wt)
        ho, wo = (h - 1) // s + 1, (w - 1) // s + 1
        out = ops.Act.empty((n, ho, wo, cout), "cuda", True)
        bias = torch.zeros(cout, device="cuda")
        gf = 2.0 * n * ho * wo * cout * cin * k * k / 1e9
        row = {"shape": (n, h, w, cin, cout, k, s), "gflop": gf}
        lib = a2x_import.pkg("_lib").load()
        for tag, kw, dbg in (("full", dict(relu=1), 0), ("no_hi_plane", dict(relu=1, write_hi=False), 0), ("no_stores", dict(relu=1), 1)):
            lib.a2x_debug_set(9, dbg)
            ms = timeit(lambda: ops.conv_fwd(x, pk, k, s, out, shift=bias, **kw))
            lib.a2x_debug_set(9, 0)
            row[tag + "_ms"] = ms
            row[tag + "_tflops"] = gf / ms
        res.append(row)
        print(row)
    print("MICROBENCH " + json.dumps(res))


if __name__ == "__main__":
    main()
