"""Aggregate the stall samples of an ncu report by kernel phase: segments are delimited by barrier / MMA / mbarrier
instructions in the SASS.  usage: python tools/ncu_phase_summary.py report.ncu-rep"""
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

src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", *SEL], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
tot_i = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
names = ["stall_long_sb", "stall_short_sb", "stall_mio", "stall_barrier", "stall_wait", "stall_math", "stall_not_selected",
         "stall_selected", "stall_lg"]
acc, ninstr, st = 0, 0, {k: 0 for k in names}


def flush():
    global acc, ninstr, st
    if acc > 0.004 * tot:
        print(f"      segment: samples {100 * acc / tot:5.1f}%  instr {100 * ninstr / tot_i:5.1f}%  " +
              " ".join(f"{k[6:]}={100 * v / tot:.1f}" for k, v in st.items() if v > 0.008 * tot))
    acc, ninstr, st = 0, 0, {k: 0 for k in names}


for r in data:
    s = r[ix["Source"]]
    n = int(r[ix["# Samples"]] or 0)
    if any(m in s for m in ["BAR.SYNC", "UTCHMMA", "SYNCS.PHASECHK", "LDGDEPBAR", "UTCBAR"]):
        flush()
        print(f"---- {r[ix['Address']][-5:]} {s[:60]:60} samples {100 * n / tot:.1f}%")
    acc += n
    ninstr += int(r[ix["Instructions Executed"]] or 0)
    for k in names:
        st[k] += int(r[ix[k]] or 0)
flush()
print("total samples", tot, "instructions", tot_i)
