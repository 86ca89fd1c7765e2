"""Per-call timing of the tap-GEMM family inside one eager headline step (CUDA events around every C-ABI call, side
stream off): shape, algorithmic GFLOP, ms, TFLOP/s (algorithmic; x3 for the bf16 MMAs executed). Prints a table sorted by
time and one JSON line. Run on the GPU box:  python scripts/profile_conv_calls.py [--config 2|5]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import a2x_import
import bench


def main():
    which = int(sys.argv[sys.argv.index("--config") + 1]) if "--config" in sys.argv else 2
    dev = torch.device("cuda", 0)
    libmod = a2x_import.pkg("_lib")
    w = bench.Workload(which, torch, dev, seed=0)
    w.model.train()
    for _ in range(3):
        w.train_step()
    torch.cuda.synchronize()
    libmod.PROFILE = []
    w.model.engine.use_side_stream = False
    w.train_step()
    torch.cuda.synchronize()
    prof, libmod.PROFILE = libmod.PROFILE, None
    rows = []
    total = 0.0
    for name, cargs, e0, e1 in prof:
        ms = e0.elapsed_time(e1)
        total += ms
        base = name[:-3] if name.endswith("_ex") else name
        if base in ("a2x_conv2d_fwd", "a2x_conv2d_dgrad", "a2x_conv2d_wgrad", "a2x_deconv_fwd", "a2x_deconv_dgrad", "a2x_deconv_wgrad"):
            s = cargs[0]._obj
            if base.startswith("a2x_conv2d"):
                ho, wo = (s.h - 1) // s.stride + 1, (s.w - 1) // s.stride + 1
                fl = 2.0 * s.n * ho * wo * s.cout * s.cin * s.ksize * s.ksize
            else:
                fl = 2.0 * s.n * s.h * s.w * s.cin * s.cout * s.stride * s.stride
            rows.append((base, (s.n, s.h, s.w, s.cin, s.cout, s.ksize, s.stride), fl / 1e9, ms))
    agg = {}
    for base, shp, gf, ms in rows:
        k = (base, shp)
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += gf
        a[2] += ms
    out = []
    print("%-18s %-34s %5s %9s %8s %8s" % ("call", "(n,h,w,cin,cout,k,s)", "calls", "GFLOP", "ms", "TFLOP/s"))
    for (base, shp), (c, gf, ms) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print("%-18s %-34s %5d %9.1f %8.3f %8.1f" % (base, shp, c, gf, ms, gf / ms))
        out.append({"call": base, "shape": shp, "calls": c, "gflop": gf, "ms": ms, "tflops": gf / ms})
    print("PROFILE_CONV " + json.dumps({"config": which, "step_ms_serialised": total, "rows": out}))


if __name__ == "__main__":
    main()
