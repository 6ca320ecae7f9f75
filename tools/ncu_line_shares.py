#!/usr/bin/env python3
"""Executed instructions and stall samples per source line of one kernel launch, from an
`ncu --set full --import-source on` report (needs -lineinfo builds):

    tools/ncu_line_shares.py <report.ncu-rep> <out.txt> [units_per_launch]

Writes the 40 heaviest lines (share of executed warp instructions, share of stall samples, long-scoreboard
and barrier samples) and per-file totals; with units_per_launch also instructions per unit."""
import collections
import csv
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, hdr = None, None
    inst, smp, lsb, bar, text = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter(), {}
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ie, isamp, il, ib = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("stall_long_sb"), hdr.index("stall_barrier")
            continue
        if hdr is None or r[0] == "Function Name":
            continue
        try:
            ln, n, s, l, b = int(r[0]), int(r[ie]), int(r[isamp]), int(r[il]), int(r[ib])
        except (ValueError, IndexError):
            continue
        k = (cur, ln)
        inst[k] += n; smp[k] += s; lsb[k] += l; bar[k] += b; text[k] = r[1].strip()[:100]
    ti, ts = sum(inst.values()), max(1, sum(smp.values()))
    with open(out, "w") as f:
        f.write("# %s: executed warp instructions %d, stall samples %d (long_scoreboard %.1f %%, barrier %.1f %%)\n" % (rep, ti, ts, 100 * sum(lsb.values()) / ts, 100 * sum(bar.values()) / ts))
        files = collections.Counter()
        for (fn, _), v in inst.items():
            files[fn] += v
        for fn, v in files.most_common():
            f.write("# file %-28s %5.1f %% of instructions%s\n" % (fn, 100 * v / ti, "  %8.1f per unit" % (32 * v / units) if units else ""))
        f.write("# inst%  samples%  long_sb%  barrier%  file:line  source\n")
        for k, v in inst.most_common(40):
            f.write("%6.2f %8.2f %9.2f %8.2f  %s:%d  %s\n" % (100 * v / ti, 100 * smp[k] / ts, 100 * lsb[k] / ts, 100 * bar[k] / ts, k[0], k[1], text[k]))
    print(open(out).read()[:1500])


if __name__ == "__main__":
    main()
