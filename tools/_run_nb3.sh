cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fused.py tests/test_gpu_step16.py -q -m gpu --tb=short -x > gpurun_out/t_nb3.log 2>&1; echo rc=$?; tail -5 gpurun_out/t_nb3.log
for v in tc3; do
GRL_NODE_BWD=$v timeout 300 python bench.py --steps 20 --warmup 3 --single-precision --no-cpu-baseline --no-side-workloads --repeats 3 > gpurun_out/r2_nb_$v.json 2> gpurun_out/r2_nb_$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_nb_$v.json"))
k=d["roofline"]["kernel_ms_per_step"]
print("$v", round(d["value"]), round(d["ms_per_step"],3), "node_bwd", k["grl_fbconv_node_bwd_tc"], "node_fwd", k["grl_fbconv_node_fwd_tc"])
PY
done
