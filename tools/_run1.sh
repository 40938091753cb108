cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_step16.py -q -m gpu --tb=short -x > gpurun_out/t_fused.log 2>&1; tail -2 gpurun_out/t_fused.log
for v in ws3 ws3; do
GRL_FUSED_BWD=$v timeout 300 python bench.py --steps 20 --warmup 3 --single-precision --no-cpu-baseline --no-side-workloads --repeats 3 > gpurun_out/r2_cmp_$v.json 2> gpurun_out/r2_cmp_$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_cmp_$v.json"))
k=d["roofline"]["kernel_ms_per_step"]
print("$v", round(d["value"]), round(d["ms_per_step"],3), "bwd", k["grl_fbconv_edge_fused_bwd"], "fwd", k["grl_fbconv_edge_fused_fwd"])
PY
done
