"""SASS mnemonic counts per kernel of the built library (the proof that tcgen05 / TMEM / TMA / packed math are what
runs).  usage: python tools/sass_summary.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "geometry_rl_b200", "libgrl_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "FFMA2", "HFMA2", "MUFU.TANH", "FFMA"]
counts, fn, total = collections.OrderedDict(), None, collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[fn] = collections.Counter()
        continue
    if fn and re.search(r"/\*[0-9a-f]{4,5}\*/", line):
        counts[fn]["instr"] += 1
        for k in MN:
            if re.search(r"\b" + re.escape(k) + r"(\b|\.)", line) and not (k == "FFMA" and "FFMA2" in line):
                counts[fn][k] += 1
                total[k] += 1
out = ["# SASS mnemonic counts of geometry_rl_b200/libgrl_b200.so (cuobjdump -sass, sm_100a)\n",
       "UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (tensor-map TMA),",
       "UBLKCP / UBLKPF = 1-D bulk copy / L2 prefetch, SYNCS = mbarrier ops, FFMA2 / HFMA2 = packed fp32 / fp16 FMA.\n",
       "| kernel | instr | " + " | ".join(MN) + " |", "|---|---:|" + "---:|" * len(MN)]
for fn, c in counts.items():
    if any(c[k] for k in MN[:6]) or c["FFMA2"] or c["HFMA2"]:
        out.append(f"| `{fn[:60]}` | {c['instr']} | " + " | ".join(str(c[k]) if c[k] else "" for k in MN) + " |")
out.append("\nlibrary totals: " + ", ".join(f"{k} {total[k]}" for k in MN))
text = "\n".join(out) + "\n"
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
