#!/usr/bin/env python3
"""Dynamic SASS opcode histogram of one kernel launch from an `ncu --set full --import-source on`
report (per-instruction "Instructions Executed" of the source page), so that the work the
roofline is quoted against is a measured number and not a constant typed into bench.py.

    tools/ncu_opcode_hist.py <report.ncu-rep> <out.json> [--units N --unit-name point_addition] [--note "..."]

Classes (B200, measured rates in profiles/r01/*_r01.txt):
  imad_wide_rr64   IMAD.WIDE[.U32] Rd, Ra, Rb, Rc     two register multiplicands + 64-bit register addend: HALF rate (29/clk/SM)
  imad_wide_rz     IMAD.WIDE[.U32] Rd, Ra, Rb|imm, RZ no addend: full rate
  imad_wide_imm    IMAD.WIDE[.U32] Rd, Ra, imm|c[], Rc constant multiplicand + 64-bit register addend: HALF rate too
                   (31/clk/SM, profiles/r02/wide_operands3_ubench.txt; round 1 took it for full rate from a
                   microbenchmark whose repeated immediate products the compiler had merged)
  imad32           IMAD / IMAD.X / IMAD.SHL / IMAD.IADD / IMAD.HI ... 32-bit forms (IMAD.MOV counted as mov)
  lop_shf          LOP3 / SHF (share the multiply pipe's issue port: 58/clk/SM)
  iadd             IADD3 / IADD3.X / IADD
  ldst_global, ldst_shared, ldst_local, barrier, control, other
`mul_pipe_slots` = 2 * (imad_wide_rr64 + imad_wide_imm) + imad_wide_rz + imad32 (issue slots on the
integer-multiply pipe: an IMAD.WIDE with a 64-bit register addend occupies it twice as long), counted per
warp instruction.
"""
import argparse
import collections
import csv
import json
import re
import subprocess

PRED = re.compile(r"^@!?U?P\d+\s+")


def classify(sass: str) -> str:
    s = PRED.sub("", sass.strip())
    op = s.split()[0] if s else "?"
    base = op.split(".")[0]
    if base == "IMAD":
        if ".MOV" in op:
            return "mov"
        if ".WIDE" in op:
            ops = [o.strip() for o in s[len(op):].split(",")]
            b = ops[2] if len(ops) > 2 else ""          # Rd, Ra, Rb, Rc
            c = ops[3] if len(ops) > 3 else ""
            if c.startswith("RZ"):
                return "imad_wide_rz"                    # no addend (any multiplicand): full rate
            if not b.lstrip("-~|").startswith("R") or b.startswith("RZ"):
                return "imad_wide_imm"                   # immediate, constant bank or uniform multiplicand
            return "imad_wide_rr64"
        return "imad32"
    if base in ("LOP3", "SHF", "LOP", "SHL", "SHR", "PRMT", "BMSK", "SGXT", "LEA"):
        return "lop_shf"
    if base in ("IADD3", "IADD", "IABS", "IMNMX", "VIADD", "VIMNMX"):
        return "iadd"
    if base in ("ISETP", "PLOP3", "SEL", "ICMP", "P2R", "R2P"):
        return "pred_sel"
    if base in ("MOV", "S2R", "CS2R", "S2UR", "LDC", "LDCU", "R2UR", "UMOV") or base.startswith("U"):
        return "mov"
    if base in ("LDG", "STG", "LD", "ST", "ATOMG", "REDG", "ATOM", "RED"):
        return "ldst_global"
    if base in ("LDS", "STS", "LDSM", "ATOMS"):
        return "ldst_shared"
    if base in ("LDL", "STL"):
        return "ldst_local"
    if base in ("BAR", "WARPSYNC", "BSSY", "BSYNC", "DEPBAR", "MEMBAR", "ERRBAR", "NANOSLEEP"):
        return "barrier"
    if base in ("BRA", "BRX", "JMP", "EXIT", "RET", "CALL", "BREAK", "NOP", "YIELD"):
        return "control"
    if base in ("SHFL", "VOTE", "MATCH", "REDUX"):
        return "shuffle"
    if base in ("DFMA", "DADD", "DMUL", "DSETP"):
        return "fp64"
    return "other"


def histogram(report: str, launch: int = 0) -> dict:
    raw = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for line in raw.splitlines():                        # a report may hold several launches
        if line.startswith('"Kernel Name"'):
            cur = {"name": next(csv.reader([line]))[1], "lines": []}
            blocks.append(cur)
        elif cur is not None:
            cur["lines"].append(line)
    blk = blocks[launch]
    rows = list(csv.reader(blk["lines"]))
    hdr = rows[0]
    i_src, i_exec, i_thr = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    hist, ops = collections.Counter(), collections.Counter()
    total = thread_total = 0
    for r in rows[1:]:
        if len(r) <= i_thr:
            continue
        n = int(r[i_exec] or 0)
        if n == 0:
            continue
        total += n
        thread_total += int(r[i_thr] or 0)
        hist[classify(r[i_src])] += n
        ops[PRED.sub("", r[i_src].strip()).split()[0]] += n
    slots = 2 * (hist["imad_wide_rr64"] + hist["imad_wide_imm"]) + hist["imad_wide_rz"] + hist["imad32"]
    mul_inst = hist["imad_wide_rr64"] + hist["imad_wide_rz"] + hist["imad_wide_imm"] + hist["imad32"]
    return {
        "kernel": blk["name"], "report": report,
        "warp_instructions_executed": total, "thread_instructions_executed": thread_total,
        "by_class": dict(sorted(hist.items(), key=lambda x: -x[1])),
        "share_by_class": {k: round(v / total, 4) for k, v in sorted(hist.items(), key=lambda x: -x[1])},
        "multiply_warp_instructions": mul_inst, "mul_pipe_slots": slots,
        "top_opcodes": dict(ops.most_common(24)),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--units", type=float, default=0.0, help="algorithmic units one launch processes (e.g. point additions)")
    ap.add_argument("--unit-name", default="unit")
    ap.add_argument("--note", default="")
    ap.add_argument("--launch", type=int, default=0, help="which profiled launch of the report (0 = first)")
    a = ap.parse_args()
    out = histogram(a.report, a.launch)
    out["note"] = a.note
    if a.units > 0:
        total, mul_inst, slots = out["warp_instructions_executed"], out["multiply_warp_instructions"], out["mul_pipe_slots"]
        out["unit"] = a.unit_name
        out["units_per_launch"] = a.units
        out["thread_inst_per_unit"] = 32.0 * total / a.units          # one unit is one lane's work
        out["multiply_thread_inst_per_unit"] = 32.0 * mul_inst / a.units
        out["mul_pipe_thread_slots_per_unit"] = 32.0 * slots / a.units
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
