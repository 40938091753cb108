cd $GRAFT_REPO_ROOT
N=${N:-2}
GRL_TIMELINE=gpurun_out/r02_timeline_n$N.json timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-side-workloads --single-precision > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err; echo rc=$?; tail -c 400 gpurun_out/r02_scale_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_scale_n$N.json"))
print(d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["details"]["regions_ms"][:5])

PY
