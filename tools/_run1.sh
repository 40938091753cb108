cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; tail -15 gpurun_out/t_all.log
timeout 300 python bench.py --steps 20 --warmup 3 --single-precision --no-cpu-baseline --no-side-workloads --repeats 3 > gpurun_out/r2_ws1_hepi_8192.json 2> gpurun_out/r2_ws1_hepi_8192.err; tail -c 400 gpurun_out/r2_ws1_hepi_8192.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_ws1_hepi_8192.json"))
print(round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac_by_kernel"])
PY
