#!/usr/bin/env python
"""Per-region instruction / stall-sample accounting of one kernel from an ncu source-page CSV
(ncu -i rep --page source --csv --print-source sass).  Regions are split at BAR / SYNCS / back-edges.
Usage: tools/ncu_regions.py <source.csv> [n_warp_cells]"""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
body = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
iA, iS, iN, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(body[0][iA], 16)
norm = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tot_s = sum(int(r[iN]) for r in body); tot_e = sum(int(r[iE]) for r in body)
print(f"total samples {tot_s}, warp-inst {tot_e} ({tot_e / norm:.1f} per unit)")
# split where the executed count changes by a large factor or at barriers
regs, cur = [], None
prev_e = None
for r in body:
    e = int(r[iE]); src = r[iS].strip()
    newreg = cur is None or "BAR.SYNC" in src or (prev_e is not None and (e > 1.5 * prev_e + 10 or e * 1.5 + 10 < prev_e))
    if newreg:
        cur = {"start": int(r[iA], 16) - base, "n": 0, "s": 0, "e": 0, "ops": {}}
        regs.append(cur)
    cur["n"] += 1; cur["s"] += int(r[iN]); cur["e"] += e
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    cur["ops"][op] = cur["ops"].get(op, 0) + e
    cur["end"] = int(r[iA], 16) - base
    prev_e = e
for g in regs:
    if g["e"] < 0.005 * tot_e and g["s"] < 0.005 * tot_s:
        continue
    top = sorted(g["ops"].items(), key=lambda kv: -kv[1])[:7]
    print(f"  [{g['start']:#06x}-{g['end']:#06x}] {g['n']:4d} instr  exec/instr {g['e'] / g['n']:9.0f}  warp-inst {100 * g['e'] / tot_e:5.1f}% ({g['e'] / norm:7.1f}/unit)  samples {100 * g['s'] / tot_s:5.1f}%   "
          + " ".join(f"{k}:{v / norm:.0f}" for k, v in top))
