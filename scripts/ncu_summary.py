#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): one line per launch with the metrics the roofline needs.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--md]
"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("smsp__cycles_active.avg", "cyc_active"), ("sm__cycles_elapsed.max", "cyc_elapsed")]


def main():
    rep = sys.argv[1]
    md = "--md" in sys.argv
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in WANT if m in idx]
    head = ["kernel"] + [n for _, n in cols]
    print(("| " + " | ".join(head) + " |") if md else "\t".join(head))
    if md:
        print("|" + "---|" * len(head))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("et::", "")
        vals = []
        for m, _ in cols:
            v, u = r[idx[m]], units[idx[m]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {u}".strip() if u not in ("", "%") else v)
        print(("| " + " | ".join([name] + vals) + " |") if md else "\t".join([name] + vals))


if __name__ == "__main__":
    main()
