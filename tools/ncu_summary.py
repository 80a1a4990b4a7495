"""Markdown summary of the `ncu --set full` captures and the launch list of one evidence pass
(tools/gpu_evidence_r2b.sh).  Runs on the build host (ncu -i reads the reports, no GPU).
   python tools/ncu_summary.py gpurun_out prof_r02b_ > profiles/r02b_final_summary.md"""
import csv
import io
import os
import subprocess
import sys
from collections import OrderedDict

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
prefix = sys.argv[2] if len(sys.argv) > 2 else "prof_r02b_"
TITLES = OrderedDict([
    ("c2", "C2 double-Gauss, 9 997 351 rays x 13 entries, resident inputs (the headline kernel)"),
    ("c2spot", "C2, the bench step: trace + spot sums of the image plane in ONE launch (POLICY bit 32)"),
    ("c2gen", "C2, bundle generated in the kernel prologue (no input arrays)"),
    ("c3", "C3 even asphere, 9 997 351 rays x 4 entries (asphere-only instantiation), resident inputs"),
    ("c4", "C4 birefringent doublet, 1 000 519 -> 4 002 076 rays, the ONE complex-stretch launch"),
    ("c5", "C5 GRIN, 1 000 519 rays x 4 entries, 204 integrator steps"),
])
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "no_instruction", "math_pipe_throttle"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}

print("# Round-2 closing evidence (tools/gpu_evidence_r2b.sh, one B200, build of the commit that adds this file)\n")
print("Captures: `ncu --set full --clock-control none --import-source on`, one launch each after warm-up launches "
      "(`tools/profile_target.py`).  Times under ncu are cold-cache and serialised: the CUDA-event medians are in "
      "`r02b_final_timings.txt`.\n")
for (tag, title) in TITLES.items():
    rep = os.path.join(src, prefix + tag + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    (hdr, units, vals) = (rows[0], rows[1], rows[2])
    col = {h: i for (i, h) in enumerate(hdr)}
    print("## %s\n" % title)
    print("`%s`\n" % vals[col["Kernel Name"]])
    print("| metric | unit | value |\n|---|---|---|")
    for m in WANT:
        if m in col:
            print("| %s | %s | %s |" % (m, units[col[m]], vals[col[m]]))
    for st in STALLS:
        m = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % st
        if m in col:
            print("| %s | %s | %s |" % (m, units[col[m]], vals[col[m]]))
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(vals[col[m]]) * UNIT.get(units[col[m]], 1.0)
    print("\nDRAM traffic of the launch: %.3f GB\n" % (tot / 1e9))

lp = os.path.join(src, "launches.csv")
if os.path.exists(lp):
    text = open(lp).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        a = agg.setdefault(r["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    print("## Launch list of `python bench.py --steps 2 --warmup 3 --no-cpu` (ncu `gpu__time_duration.sum`, first 400 "
          "launches: `r02b_final_launches_bench.csv`)\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for (k, a) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f %% |" % (k[:110], a[0], a[1], 100.0 * a[1] / total))
