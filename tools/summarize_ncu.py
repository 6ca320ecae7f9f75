#!/usr/bin/env python3
"""Summaries kept under profiles/: (1) per-kernel totals of an ncu launch list (csv),
(2) the key metrics of an `ncu --set full` report of k_msm (via `ncu -i ... --page raw --csv`)."""
import collections, csv, json, subprocess, sys

def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        k = row["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    step = {k: v for k, v in agg.items() if k.startswith("rk::") and not k.startswith(("rk::k_table", "rk::k_setup", "rk::k_roots", "rk::k_imad"))}
    tot = sum(v[1] for v in step.values())
    with open(out, "w") as f:
        f.write("# per-kernel totals of %s (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n" % path)
        f.write("# share = of the steady-state step kernels (table build / setup / peak microbenchmark listed but excluded)\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            sh = "%.4f" % (t / tot) if k in step else "   -  "
            f.write("%-44s launches %5d  total_ms %11.3f  avg_ms %9.3f  share %s\n" % (k[:44], n, t, t / n, sh))
    print(open(out).read())

def full(rep, out, note):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__icc_request_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio"] + \
           [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(out, "w") as f:
        f.write("# %s\n" % note)
        for k in keys:
            if k in d:
                f.write("%-92s %s %s\n" % (k, d[k][0], d[k][1]))
    print(open(out).read())
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd = float(d["dram__bytes_read.sum"][0]) * scale[d["dram__bytes_read.sum"][1]]
    wr = float(d["dram__bytes_write.sum"][0]) * scale[d["dram__bytes_write.sum"][1]]
    return rd + wr

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        t = full(sys.argv[2], sys.argv[3], sys.argv[4])
        if len(sys.argv) > 6:
            blobs = int(sys.argv[6])
            json.dump({"kernel": "k_msm", "dram_bytes_per_launch": t, "blobs_per_launch": blobs, "source": sys.argv[3],
                       "algorithmic_bytes_per_launch": blobs * (4096 * 17 * 96 + 131072)}, open(sys.argv[5], "w"), indent=1)
