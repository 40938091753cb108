cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_critic.py tests/test_gpu_parity.py tests/test_gpu_boundary.py tests/test_gpu_step.py tests/test_gpu_step16.py tests/test_gpu_dp.py -q -m gpu --tb=short > gpurun_out/t_critic.log 2>&1; grep -E "^E  |passed|failed|FAILED" gpurun_out/t_critic.log | head -30
timeout 300 python bench.py --steps 20 --warmup 3 --single-precision --no-cpu-baseline --no-side-workloads --repeats 3 > gpurun_out/r2_critic_hepi.json 2> gpurun_out/r2_critic_hepi.err; tail -c 300 gpurun_out/r2_critic_hepi.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_critic_hepi.json"))
print(round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms_per_step"])
PY
