cd $GRAFT_REPO_ROOT
# launch order per step (eager): fwd: edge_fused_fwd L0, node_fwd L0, edge_fused_fwd L1, node_fwd L1; bwd: L1 first, then L0
run() { # name regex skip
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$2 --launch-skip $3 --launch-count 1 -f -o gpurun_out/r02_$1 python bench.py --steps 2 --warmup 1 --single-precision --no-cpu-baseline --no-side-workloads --repeats 1 --no-graph > gpurun_out/ncu_$1.log 2>&1; echo $1 rc=$?
}
run node_fwd fbconv_node_fwd_tc2 ${SKIP_FWD:-26}
run edge_fwd edge_fused_fwd ${SKIP_FWD:-26}

