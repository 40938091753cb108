cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$KERNEL --launch-skip $SKIP --launch-count 1 -f -o gpurun_out/r02_$NAME python bench.py --steps 2 --warmup 1 --single-precision --no-cpu-baseline --no-side-workloads --repeats 1 --no-graph > gpurun_out/ncu_$NAME.log 2>&1; echo $NAME rc=$?
