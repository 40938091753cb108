cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_dp2.json 2> gpurun_out/r2_dp2.err; echo rc=$?; tail -c 600 gpurun_out/r2_dp2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_dp2.json"))
print(d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["details"]["regions_ms"])
for k,v in d["configs"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("error"))
PY
