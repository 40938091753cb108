cd $GRAFT_REPO_ROOT
[ -n "$SKIP_TESTS" ] || { timeout 1200 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/t_all.log 2>&1; echo rc=$?; tail -3 gpurun_out/t_all.log; }
for v in ${VARIANTS:-tc3}; do
GRL_NODE_BWD=$v timeout 300 python bench.py --steps 20 --warmup 3 --single-precision --no-cpu-baseline --no-side-workloads --repeats 3 > gpurun_out/r2_all_$v.json 2> gpurun_out/r2_all_$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_all_$v.json"))
k=d["roofline"]["kernel_ms_per_step"]
print("$v", round(d["value"]), round(d["ms_per_step"],3), {a:b for a,b in list(k.items())[:6]})
PY
done
