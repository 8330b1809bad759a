#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline raw metrics, stall reasons, opcode mix and
the hot loop of one kernel.  Usage: tools/ncu_summary.py <report.ncu-rep> [kernel-substring]"""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
       "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "smsp__average_warp_latency_per_inst_issued.ratio", "lts__t_sector_hit_rate.pct"]


def ncu(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = ncu(rep, "raw")
    hdr = rows[0]
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat not in name or name in seen:
            continue
        seen.add(name)
        print("== kernel:", name[:110])
        for k in RAW:
            if k in hdr:
                print(f"  {k:72s} {r[hdr.index(k)]} {rows[1][hdr.index(k)]}")
    rows = ncu(rep, "source", ["--print-source", "sass"])
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "body": []}
            blocks.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["body"].append(r)
    done = set()
    for b in blocks:
        if pat not in b["name"] or b["name"] in done or not b["hdr"]:
            continue
        done.add(b["name"])
        ix = {h: i for i, h in enumerate(b["hdr"])}
        stalls = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
        g = lambda r, k: int(float(r[ix[k]] or 0))
        samples = sum(g(r, "# Samples") for r in b["body"]) or 1
        tote = sum(g(r, "Instructions Executed") for r in b["body"]) or 1
        print("== source page:", b["name"][:100], "| SASS lines", len(b["body"]), "| samples", samples, "| warp-inst", tote)
        tot = {s: sum(g(r, s) for r in b["body"]) for s in stalls}
        print("  stalls: " + ", ".join(f"{s[6:]} {100 * v / samples:.1f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
        byop = {}
        for r in b["body"]:
            toks = r[ix["Source"]].split()
            op = (toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")).split(".")[0]
            d = byop.setdefault(op, [0, 0])
            d[0] += g(r, "# Samples")
            d[1] += g(r, "Instructions Executed")
        print("  opcode: %samples / %warp-inst")
        for op, (n, e) in sorted(byop.items(), key=lambda kv: -kv[1][1])[:14]:
            print(f"    {op:10s} {100 * n / samples:5.1f}% {100 * e / tote:5.1f}%")
        mx = max(g(r, "Instructions Executed") for r in b["body"])
        hot = [r for r in b["body"] if g(r, "Instructions Executed") >= 0.9 * mx]
        print(f"  hot loop: {len(hot)} SASS instr x {mx} warp-executions = {100 * len(hot) * mx / tote:.0f}% of warp-inst, "
              f"{100 * sum(g(r, '# Samples') for r in hot) / samples:.0f}% of samples")
        hs = {s: sum(g(r, s) for r in hot) for s in stalls}
        hsum = sum(g(r, "# Samples") for r in hot) or 1
        print("  hot-loop stalls: " + ", ".join(f"{s[6:]} {100 * v / hsum:.1f}%" for s, v in sorted(hs.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == "__main__":
    main()
