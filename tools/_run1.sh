cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
timeout 300 python bench.py --steps 20 --warmup 3 --single-precision --no-cpu-baseline --no-side-workloads --repeats 3 > gpurun_out/r2_ws2_hepi_8192.json 2> gpurun_out/r2_ws2_hepi_8192.err; tail -c 300 gpurun_out/r2_ws2_hepi_8192.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_ws2_hepi_8192.json"))
print(round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac_by_kernel"])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:edge_fused_bwd -s 4 -c 2 -o gpurun_out/r2_ws2_bwd python bench.py --config rigid_insertion_multi_hepi_trpl_cfg --minibatch 4096 --steps 2 --warmup 1 --no-graph --single-precision --no-cpu-baseline --no-side-workloads > gpurun_out/ncu_ws2.log 2>&1; tail -2 gpurun_out/ncu_ws2.log
