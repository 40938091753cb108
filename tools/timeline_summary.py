"""Summarise a device timeline written by bench.py (GRL_TIMELINE=...) or tools/step_timeline.py: per-stream busy time,
idle gaps on the main stream with the activities either side, time under NCCL kernels.
usage: python tools/timeline_summary.py timeline.json [out.md]"""
import collections
import json
import sys

rows = json.load(open(sys.argv[1]))
by_stream = collections.defaultdict(list)
for r in rows:
    by_stream[r[1]].append(r)
# a replayed CUDA graph spreads its nodes over several internal streams: work on the UNION of all streams
allr = sorted(rows, key=lambda r: r[2])
i0 = next((i for i, r in enumerate(allr) if r[0].startswith("grl::")), 0)
allr = allr[i0:]
t0 = allr[0][2]
t1 = max(r[2] + r[3] for r in allr)
out = []
out.append(f"# device timeline of one update step ({sys.argv[1]})\n")
busy, gaps, cur_end, last = 0.0, [], allr[0][2], allr[0]
for r in allr:
    if r[2] > cur_end:
        if r[2] - cur_end > 1.0:
            gaps.append((r[2] - cur_end, cur_end - t0, last[0].split("(")[0][:50], r[0].split("(")[0][:50]))
        busy += 0.0
        cur_start = r[2]
    if r[2] + r[3] > cur_end:
        busy += (r[2] + r[3]) - max(cur_end, r[2])
        cur_end = r[2] + r[3]
        last = r
out.append(f"first library kernel to last activity: {(t1 - t0) / 1e3:.3f} ms; some stream busy for {busy / 1e3:.3f} ms; "
           f"sum over streams {sum(r[3] for r in allr) / 1e3:.3f} ms\n")
out.append("| stream | activities | busy ms | NCCL ms |\n|---|---:|---:|---:|")
for k, v in sorted(by_stream.items(), key=lambda kv: -sum(x[3] for x in kv[1])):
    nccl = sum(x[3] for x in v if "nccl" in x[0].lower())
    out.append(f"| {k} | {len(v)} | {sum(x[3] for x in v) / 1e3:.3f} | {nccl / 1e3:.3f} |")
out.append(f"\nidle gaps > 1 us (no stream active): {len(gaps)}, {sum(g[0] for g in gaps) / 1e3:.3f} ms in total\n")
out.append("| gap us | at ms | after | before |\n|---:|---:|---|---|")
for g in sorted(gaps, reverse=True)[:15]:
    out.append(f"| {g[0]:.1f} | {g[1] / 1e3:.2f} | `{g[2]}` | `{g[3]}` |")
nccl = [r for r in allr if "nccl" in r[0].lower()]
out.append(f"\nNCCL kernels: {len(nccl)}, {sum(r[3] for r in nccl) / 1e3:.3f} ms in total; longest " +
           ", ".join(f"{r[3]:.0f} us @{(r[2] - t0) / 1e3:.2f} ms" for r in sorted(nccl, key=lambda r: -r[3])[:4]))
tot = collections.Counter()
for r in rows:
    tot[r[0].split("(")[0][:60]] += r[3]
out.append("\n| ms | activity (all streams) |\n|---:|---|")
for k, v in tot.most_common(16):
    out.append(f"| {v / 1e3:.3f} | `{k}` |")
text = "\n".join(out) + "\n"
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
