"""Summarise an ncu report (one kernel): headline metrics + the SASS instructions holding most stall samples.
usage: python tools/ncu_src_summary.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

# optional: --skip N selects the N-th captured launch of a multi-kernel report
SEL = []
if "--skip" in sys.argv:
    i = sys.argv.index("--skip")
    SEL = ["--launch-skip", sys.argv[i + 1], "--launch-count", "1"]
    del sys.argv[i:i + 2]

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", *SEL], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if h in KEYS:
            print(f"  {h} = {vals[i]} {units[i]}")
    st = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(vals[i]) for i, h in enumerate(hdr)
          if "average_warps_issue_stalled" in h and "per_issue_active" in h and vals[i]}
    print("  stalls/issue:", ", ".join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *SEL], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot)
names = ["stall_long_sb", "stall_short_sb", "stall_mio", "stall_barrier", "stall_wait", "stall_math", "stall_not_selected",
         "stall_lg", "stall_dispatch", "stall_branch_resolving"]
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:top_n]:
    n = int(r[ix["# Samples"]] or 0)
    st = {k[6:]: int(r[ix[k]] or 0) for k in names if k in ix}
    st = {k: v for k, v in st.items() if v > 0.2 * n}
    print(f"{r[ix['Address']][-5:]} {100.0 * n / tot:5.1f}% {r[ix['Instructions Executed']]:>9} {r[ix['Source']][:64]:64} {st}")
