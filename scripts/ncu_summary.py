#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full --import-source on) into the text files committed under profiles/:
    python scripts/ncu_summary.py REPORT.ncu-rep OUT_PREFIX
writes OUT_PREFIX_raw.csv (ncu --page raw --csv of the first kernel in the report) and OUT_PREFIX_sass_regions.txt
(the SASS of the kernel cut into regions of 80 instructions: share of the warp-stall samples, share of the executed
instructions, the dominant instruction mnemonics and the top stall reasons of each region)."""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(out + "_raw.csv", "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = ["kernel: %s" % vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""]
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            lines.append("%-90s %-14s %s" % (h, u, v))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    # find the header row of the SASS table
    hi = next((i for i, r in enumerate(srows) if "Source" in r and any("Sampling" in c for c in r)), None)
    if hi is not None:
        h = srows[hi]
        c_src = h.index("Source")
        c_smp = next(i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)")
        c_exe = next((i for i, c in enumerate(h) if c.startswith("Instructions Executed")), None)
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_")]
        body = [r for r in srows[hi + 1:] if len(r) == len(h)]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        tot_s = sum(num(r[c_smp]) for r in body) or 1.0
        tot_e = sum(num(r[c_exe]) for r in body) if c_exe is not None else 1.0
        lines.append("")
        lines.append("instr %d samples %d exec %d" % (len(body), tot_s, tot_e))
        for b in range(0, len(body), 80):
            chunk = body[b:b + 80]
            smp = sum(num(r[c_smp]) for r in chunk)
            exe = sum(num(r[c_exe]) for r in chunk) if c_exe is not None else 0.0
            ops = Counter()
            for r in chunk:
                m = r[c_src].strip().split()
                if m:
                    op = m[1] if m[0].startswith("@") and len(m) > 1 else m[0]
                    if op.split(".")[0] in ("DMMA", "DFMA", "DMUL", "DADD", "SHFL", "LDS", "STS", "LDG", "STG", "LDGSTS", "BAR", "MUFU", "ATOMG", "ATOMS", "UTMALDG", "SYNCS"):
                        ops[".".join(op.split(".")[:2]) if op.startswith("MUFU") else op.split(".")[0]] += 1
            st = sorted(((sum(num(r[i]) for r in chunk), c[6:]) for i, c in stall_cols), reverse=True)[:3]
            lines.append("%5d smp %5.2f%% ex %5.2f%% | %s | %s" % (b, 100 * smp / tot_s, 100 * exe / (tot_e or 1), " ".join("%s:%d" % kv for kv in ops.most_common(5)),
                                                                 ", ".join("%s %.1f" % (n, 100 * v / tot_s) for v, n in st)))
    open(out + "_sass_regions.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
