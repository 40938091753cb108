cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -x -k "topology" > gpurun_out/t_r1.log 2>&1; echo rc=$?; tail -4 gpurun_out/t_r1.log
timeout 600 python - <<'PY' > gpurun_out/r02_topology_rebuild.json 2> gpurun_out/r02_topology_rebuild.err
import json, torch, sys
sys.path.insert(0, ".")
import bench
print(json.dumps(bench.topology_rebuild_cost(torch.device("cuda"), "bf16"), indent=1))
PY
cat gpurun_out/r02_topology_rebuild.json; tail -3 gpurun_out/r02_topology_rebuild.err
