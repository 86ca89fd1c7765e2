#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches / total us / share of the
LAST complete step found in the capture (a step starts at vox_key_kernel)."""
import collections
import csv
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        if x.get("Metric Name") == "gpu__time_duration.sum":
            v = float(x["Metric Value"].replace(",", ""))
            u = x["Metric Unit"]
            v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
            rows.append((x["Kernel Name"], v))
    idx = [i for i, r in enumerate(rows) if "vox_key" in r[0]]
    if len(idx) >= 2:
        step = rows[idx[-2]:idx[-1]]
    elif idx:
        step = rows[idx[-1]:]
    else:
        step = rows
    tot = sum(v for _, v in step)
    agg = collections.OrderedDict()
    for k, v in step:
        k = k.split("(")[0]
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += v
    o = ["%d launches, %.1f us total (cold-cache, serialised: compare shares)" % (len(step), tot), "",
         "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        o.append("| `%s` | %d | %.1f | %.1f%% |" % (k[:90], n, v, 100 * v / tot))
    txt = "\n".join(o)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:])
