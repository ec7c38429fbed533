"""key metrics of every kernel in an .ncu-rep (or in the CSV of `ncu -i report.ncu-rep --page raw --csv`, which is what travels
back from the GPU box: reports exceed the 64 MiB gpurun_out cap) as a markdown table: python tools/ncu_summary.py report.ncu-rep|raw.csv"""
import csv
import subprocess
import sys

if sys.argv[1].endswith(".csv"):
    raw = open(sys.argv[1]).read()
else:
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines()) if len(r) > 20]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
M = [("time us", "gpu__time_duration.sum", 1), ("warp inst M", "smsp__inst_executed.sum", 1e-6),
     ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1), ("tensor %", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1),
     ("xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1), ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
     ("fma %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1), ("lsu %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
     ("dram rd MB", "dram__bytes_read.sum", None), ("dram wr MB", "dram__bytes_write.sum", None),
     ("dram %", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1),
     ("warps/SMSP", "smsp__warps_active.avg.per_cycle_active", 1), ("regs", "launch__registers_per_thread", 1),
     ("long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1),
     ("short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1),
     ("wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1),
     ("barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1)]
units = rows[1]
print("| kernel | grid | " + " | ".join(m[0] for m in M) + " |")
print("|---|---|" + "---:|" * len(M))
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:44]
    vals = []
    for lab, key, sc in M:
        if key not in col:
            vals.append("-"); continue
        v = float(r[col[key]].replace(",", "") or 0)
        u = units[col[key]]
        if sc is None:      # bytes -> MB whatever the unit ncu picked
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
            sc = 1
        if lab == "time us":
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        vals.append(f"{v * sc:.1f}" if abs(v * sc) < 1000 else f"{v * sc:.0f}")
    print(f"| `{name}` | {r[col['Grid Size']] if 'Grid Size' in col else ''} | " + " | ".join(vals) + " |")
