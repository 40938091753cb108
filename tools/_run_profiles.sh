cd $GRAFT_REPO_ROOT
TAG=${TAG:-r02}
B="python bench.py --steps 2 --warmup 1 --single-precision --no-cpu-baseline --no-side-workloads --repeats 1 --no-graph"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_l.log 2>&1; echo launches rc=$?
# one update step's worth of the message-passing kernels (skip the 44 launches of set-up, warm-up and the first timed step)
timeout 900 ncu --set full --clock-control none -k 'regex:edge_fused|fbconv_node|fbconv_fiber|embed_|absmax' --launch-skip ${SKIP:-44} --launch-count 16 -f -o gpurun_out/${TAG}_full $B > gpurun_out/${TAG}_ncu_f.log 2>&1; echo full rc=$?
ls -la gpurun_out/
