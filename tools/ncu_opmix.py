"""Opcode mix of a kernel from the source page of an ncu report: share of executed warp instructions and of stall samples
per SASS mnemonic.  usage: python tools/ncu_opmix.py report.ncu-rep [--skip N]"""
import collections
import csv
import io
import re
import subprocess
import sys

sel = []
if "--skip" in sys.argv:
    i = sys.argv.index("--skip")
    sel = ["--launch-skip", sys.argv[i + 1], "--launch-count", "1"]
    del sys.argv[i:i + 2]
src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", *sel], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot_i = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
tot_s = sum(int(r[ix["# Samples"]] or 0) for r in data)
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += int(r[ix["Instructions Executed"]] or 0)
    samp[op] += int(r[ix["# Samples"]] or 0)
print(f"warp instructions {tot_i}, stall samples {tot_s}")
for op, c in ops.most_common(24):
    print(f"{op:10s} instr {100 * c / tot_i:5.1f}%  samples {100 * samp[op] / tot_s:5.1f}%")
