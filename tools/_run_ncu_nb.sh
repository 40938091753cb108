cd $GRAFT_REPO_ROOT
for v in tc4; do
GRL_NODE_BWD=$v timeout 900 ncu --set full --import-source on --clock-control none -k regex:fbconv_node_bwd_$v --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_nb_$v python bench.py --steps 2 --warmup 1 --single-precision --no-cpu-baseline --no-side-workloads --repeats 1 --no-graph > gpurun_out/ncu_nb_$v.log 2>&1; echo rc=$?
done
