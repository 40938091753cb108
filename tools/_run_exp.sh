cd $GRAFT_REPO_ROOT
for v in tc4; do
GRL_NODE_BWD=$v timeout 300 python bench.py --steps 20 --warmup 3 --single-precision --no-cpu-baseline --no-side-workloads --repeats 3 > gpurun_out/r2_exp_$v.json 2> gpurun_out/r2_exp_$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_exp_$v.json"))
k=d["roofline"]["kernel_ms_per_step"]
print("$v", round(d["value"]), round(d["ms_per_step"],3), {a:b for a,b in list(k.items())[:4]})
PY
done
