#!/usr/bin/env python3
"""Turn an `ncu --set full --import-source on` report of one MSM kernel launch into the entry of
profiles/roofline_inputs.json that bench.py quotes its roofline from, so every number of the
bench line's `roofline` object can be re-derived from tracked files:

    tools/make_roofline_inputs.py <report.ncu-rep> <key> <units_per_launch> "<provenance note>"

key: k_msm_affine | k_msm | k_fr_eval_quot ...; units = point additions (or blobs) of the launch.
Writes profiles/roofline_inputs.json[key] and profiles/<round>/<key>_opcodes.json (the full dynamic
opcode histogram).  Needs only the report and `ncu` (no GPU).
"""
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_opcode_hist import histogram  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
TSCALE = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def raw_metrics(report, launch=0):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + launch]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    report, key, units, note = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    rnd = sys.argv[5] if len(sys.argv) > 5 else "r02"
    m = raw_metrics(report)

    def num(name):
        v, u = m[name]
        return float(v.replace(",", "")), u
    dur, du = num("gpu__time_duration.sum")
    dur *= TSCALE[du]
    rd, ru = num("dram__bytes_read.sum")
    wr, wu = num("dram__bytes_write.sum")
    dram = rd * SCALE[ru] + wr * SCALE[wu]
    h = histogram(report)
    h["note"] = note
    h["unit"] = "unit (point addition for MSM kernels, blob otherwise)"
    h["units_per_launch"] = units
    os.makedirs(os.path.join(ROOT, "profiles", rnd), exist_ok=True)
    hist_path = os.path.join("profiles", rnd, "%s_opcodes.json" % key)
    json.dump(h, open(os.path.join(ROOT, hist_path), "w"), indent=1)
    entry = {
        "source_report": os.path.basename(report), "opcode_histogram": hist_path, "note": note,
        "units_per_launch": units, "launch_ms_under_ncu": 1e3 * dur,
        "thread_inst_per_unit": 32.0 * h["warp_instructions_executed"] / units,
        "multiply_thread_inst_per_unit": 32.0 * h["multiply_warp_instructions"] / units,
        "mul_pipe_thread_slots_per_unit": 32.0 * h["mul_pipe_slots"] / units,
        "share_by_class": h["share_by_class"],
        "pipe_busy_fmaheavy": num("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed")[0] / 100.0,
        "pipe_busy_alu": num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed")[0] / 100.0,
        "issue_active": num("smsp__issue_active.avg.pct_of_peak_sustained_active")[0] / 100.0,
        "dram_bytes_per_launch": dram, "dram_bytes_per_unit": dram / units,
        "registers_per_thread": int(num("launch__registers_per_thread")[0]),
    }
    path = os.path.join(ROOT, "profiles", "roofline_inputs.json")
    allv = json.load(open(path)) if os.path.exists(path) else {}
    allv[key] = entry
    json.dump(allv, open(path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
