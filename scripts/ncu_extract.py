"""Extract a per-launch summary (duration, DRAM bytes, achieved GB/s, tensor-pipe %, occupancy, issue-active %) from
`ncu --set full` reports into one JSON for profiles/.   python scripts/ncu_extract.py out.json rep1.ncu-rep rep2.ncu-rep ..."""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "regs",
}
MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
        "ms": 1e3, "msecond": 1e3}


def main(out, *reps):
    res = {}
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            res[rep] = {"error": "empty report"}
            continue
        h, u = rows[0], rows[1]
        for r in rows[2:]:
            name = r[h.index("Kernel Name")].split("(")[0]
            d = {}
            for k, short in KEYS.items():
                if k in h:
                    i = h.index(k)
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        continue
                    d[short] = v * MULT.get(u[i], 1.0)
            if "duration" in d and "dram_read" in d:
                d["dram_gbs"] = (d["dram_read"] + d.get("dram_write", 0.0)) / d["duration"] / 1e3   # bytes / us -> GB/s
            a = res.setdefault(name, [])
            a.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for k, v in res.items():
        if isinstance(v, list):
            print("%-50s x%d  %s" % (k[:50], len(v), {kk: round(vv, 1) for kk, vv in v[0].items()}))


if __name__ == "__main__":
    main(*sys.argv[1:])
