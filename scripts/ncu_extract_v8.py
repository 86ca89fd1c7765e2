import csv, json, subprocess, sys
out = {}
for tag, rep in (("bn_train_act_kernel", "gpurun_out/bn_v8.ncu-rep"), ("lift_splat_fwd_kernel", "gpurun_out/lss_v8.ncu-rep"),
                 ("bn_relu_bwd_apply_kernel", "gpurun_out/bnb_v8.ncu-rep")):
    try:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        h = rows[0]
        recs = []
        for r in rows[2:]:
            g = lambda k: r[h.index(k)] if k in h else None
            recs.append({"kernel": g("Kernel Name"), "grid": g("launch__grid_size"), "duration_us": g("gpu__time_duration.sum"),
                         "dram_read_bytes": g("dram__bytes_read.sum"), "dram_write_bytes": g("dram__bytes_write.sum"),
                         "dram_throughput_pct": g("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                         "sm_throughput_pct": g("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                         "l2_hit_pct": g("lts__t_sector_hit_rate.pct"), "achieved_occupancy_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active")})
        units = {k: rows[1][h.index(k)] for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum") if k in h}
        out[tag] = {"units": units, "launches": recs}
    except Exception as e:
        out[tag] = {"error": str(e)}
json.dump(out, open("gpurun_out/ncu_full_v8_summary.json", "w"), indent=1)
print(json.dumps(out)[:1500])
