"""Summarise ncu outputs brought back in gpurun_out/ into small tracked text files under profiles/.
  python profiles/summarise.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
  python profiles/summarise.py full     gpurun_out/prof_r01_conv.ncu-rep profiles/r01_conv_full.md
  python profiles/summarise.py traffic  gpurun_out/prof_r01_conv.ncu-rep profiles/r01_traffic.json WORKLOAD MINIBATCH
     (DRAM bytes per launch of every C-ABI entry point: what bench.py reports as roofline.traffic)"""
import json
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__inst_executed_pipe_tensor.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg, order, n = collections.OrderedDict(), [], 0
    for row in csv.DictReader(lines):
        if "Kernel Name" not in row:
            continue
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += ns
        n += 1
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): {n} launches, {tot / 1e6:.3f} ms of kernel time (cold-cache, serialised)\n\n")
        f.write("| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% | {a[0]} | `{k[:110]}` |\n")
    print(open(dst).read()[:3000])


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n")
        for r in data:
            f.write(f"\n## {r[idx['Kernel Name']][:100]}  (grid {r[idx.get('Grid Size', 0)]}, block {r[idx.get('Block Size', 0)]})\n\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"- {k} = {r[idx[k]]} {units[idx[k]]}\n")
    print(open(dst).read()[:6000])


# C-ABI entry point -> the kernels one call launches
ABI_KERNELS = {
    "grl_fbconv_node_bwd_tc": ["fbconv_node_bwd_tc3_kernel", "fbconv_fiber_bwd_kernel"],
    "grl_fbconv_edge_fused_fwd": ["edge_fused_fwd_kernel"],
    "grl_fbconv_edge_fused_bwd": ["edge_fused_bwd_ws3_kernel"],
    "grl_fbconv_node_fwd_tc": ["fbconv_node_fwd_tc2_kernel"],
    "grl_fbconv_edge_fwd_tc": ["fbconv_edge_fwd_tc_kernel"],
    "grl_fbconv_edge_bwd_tc": ["fbconv_edge_bwd_tc2_kernel"],
    "grl_edge_basis_fwd_tc": ["edge_basis_fwd_tc_kernel"],
    "grl_edge_basis_bwd_tc": ["edge_basis_bwd_tc_kernel"],
    "grl_embed_fwd": ["embed_fwd_kernel"],
    "grl_embed_bwd": ["embed_bwd_kernel"],
    "grl_absmax": ["absmax_kernel"],
}


def traffic(src, dst, workload, minibatch):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_kernel = collections.defaultdict(list)
    for r in data:
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[idx[k]].replace(",", "")) * scale[units[idx[k]]]
        per_kernel[r[idx["Kernel Name"]]].append(b)
    res = {}
    for abi, kernels in ABI_KERNELS.items():
        tot, ok = 0.0, True
        for kn in kernels:
            v = [x for name, xs in per_kernel.items() if kn in name for x in xs]
            if not v:
                ok = False
                break
            tot += sum(v) / len(v)
        if ok:
            res[abi] = tot
    json.dump({"source": src, "workload": workload, "minibatch_per_gpu": int(minibatch),
               "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the C-ABI entry point (mean over the captured launches)",
               "kernels": res}, open(dst, "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
