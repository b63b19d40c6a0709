#!/usr/bin/env python
"""Copies the artefacts of scripts/gpu_final.sh from gpurun_out/ (scratch) into profiles/ (tracked), summarising the
ncu reports on the way.   python scripts/collect_profiles.py <run tag> [<profiles prefix>]"""
import csv
import collections
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
pre = sys.argv[2] if len(sys.argv) > 2 else "r02"


def last_json_line(path):
    for line in reversed(open(path).read().splitlines()):
        if line.startswith("{"):
            return line
    raise SystemExit(f"no JSON line in {path}")


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[2:]


# bench lines
open(os.path.join(PROF, f"{pre}_bench_default.jsonl"), "w").write(last_json_line(os.path.join(OUT, f"bench_default_{tag}.log")) + "\n")
open(os.path.join(PROF, f"{pre}_bench_reference.jsonl"), "w").write(last_json_line(os.path.join(OUT, f"bench_reference_{tag}.log")) + "\n")
shutil.copy(os.path.join(OUT, f"kernels_{tag}.json"), os.path.join(PROF, f"{pre}_kernels.json"))
fl = [json.loads(x) for x in open(os.path.join(OUT, f"forward_latency_{tag}.log")) if x.startswith("{")]
json.dump(fl, open(os.path.join(PROF, f"{pre}_forward_latency.json"), "w"), indent=1)

# config 4, k-means / init quick numbers, sanitizer summaries, full test log tail
for src, dst in ((f"config4_{tag}.log", f"{pre}_config4.jsonl"),):
    if os.path.exists(os.path.join(OUT, src)):
        open(os.path.join(PROF, dst), "w").write(last_json_line(os.path.join(OUT, src)) + "\n")
for src, dst in ((f"km_quick_{tag}.log", f"{pre}_kmeans_quick.txt"), (f"init_quick_{tag}.log", f"{pre}_init_quick.txt")):
    if os.path.exists(os.path.join(OUT, src)):
        shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
if os.path.exists(os.path.join(OUT, f"t_all_{tag}.log")):
    keep = [x for x in open(os.path.join(OUT, f"t_all_{tag}.log")) if any(k in x for k in ("passed", "failed", "label mismatches", "worst relative", "scenes"))]
    open(os.path.join(PROF, f"{pre}_gpu_tests.txt"), "w").write("".join(keep))
san = []
for tool in ("memcheck", "racecheck", "synccheck"):
    path = os.path.join(OUT, f"r2_sanitize_{tool}.log")
    if os.path.exists(path):
        lines = [x.strip() for x in open(path) if "SUMMARY" in x or "sanitize target done" in x or "Internal Sanitizer Error" in x]
        internal = sum("Internal Sanitizer Error" in x for x in lines)
        lines = [x for x in lines if "Internal Sanitizer Error" not in x]
        san.append(f"* `compute-sanitizer --tool {tool}` over `scripts/sanitize_target.py` (one small-N pass through every CUDA entry "
                   f"point): {' / '.join(lines)}" + (f" ({internal} launches not tracked by the tool)" if internal else ""))
if san:
    open(os.path.join(PROF, f"{pre}_sanitizer.md"), "w").write(f"# {pre} compute-sanitizer summaries (scripts/gpu_sanitize.sh)\n\n" + "\n".join(san) + "\n")

# launch list of the bench command
src = os.path.join(OUT, f"launches_{tag}.csv")
shutil.copy(src, os.path.join(PROF, f"{pre}_launches_bench.csv"))
lines = [x for x in open(src) if x.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ik].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) / 1e3
total = sum(v[1] for v in agg.values())
with open(os.path.join(PROF, f"{pre}_launches_bench_summary.md"), "w") as f:
    f.write(f"# {pre} launch list of `python bench.py --steps 20 --warmup 3` under ncu (serialised, cold cache: compare shares)\n\n")
    f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {name} | {cnt} | {us:.1f} | {100 * us / total:.1f}% |\n")

# ncu summaries
for rep, dst in ((f"prof_pr_{tag}.ncu-rep", f"{pre}_ncu_project_reconstruct_tma.md"), (f"prof_ops_{tag}.ncu-rep", f"{pre}_ncu_ops.md"),
                 (f"prof_lloyd_{tag}.ncu-rep", f"{pre}_ncu_kmeans_whole_fit.md")):
    if not os.path.exists(os.path.join(OUT, rep)):
        continue
    md = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), os.path.join(OUT, rep), "--md"],
                        capture_output=True, text=True).stdout
    open(os.path.join(PROF, dst), "w").write(md)

# DRAM traffic of the headline kernel
hdr, rows = ncu_raw(os.path.join(OUT, f"prof_pr_{tag}.ncu-rep"))
ir, iw, inm = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
units = list(csv.reader(subprocess.run(["ncu", "-i", os.path.join(OUT, f"prof_pr_{tag}.ncu-rep"), "--page", "raw", "--csv"],
                                       capture_output=True, text=True).stdout.splitlines()))[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = [float(r[ir].replace(",", "")) * scale[units[ir]] for r in rows]
wr = [float(r[iw].replace(",", "")) * scale[units[iw]] for r in rows]
json.dump({"kernel": rows[0][inm].split("(")[0], "launches": len(rows),
           "dram_bytes_read_per_launch": sum(rd) / len(rd), "dram_bytes_write_per_launch": sum(wr) / len(wr),
           "traffic_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd), "algorithmic_bytes_per_launch": 368000000,
           "note": "ncu --set full, N=1e6; writes below the algorithmic 208 MB because part of the output is still dirty in "
                   "the 126 MB L2 when the kernel ends",
           "source": f"profiles/{pre}_ncu_project_reconstruct_tma.md (gpurun_out/prof_pr_{tag}.ncu-rep)"},
          open(os.path.join(PROF, f"{pre}_traffic.json"), "w"), indent=1)
print("profiles updated:", sorted(x for x in os.listdir(PROF) if x.startswith(pre)))
