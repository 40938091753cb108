cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/r02a_tests.log 2>&1; tail -3 gpurun_out/r02a_tests.log
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo bench rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err; echo ref rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-workloads --single-precision --repeats 1 --no-graph > gpurun_out/r02a_ncu_l.log 2>&1; echo ncu rc=$?
python tools/step_timeline.py rigid_insertion_multi_hepi_trpl_cfg bf16 8192 > gpurun_out/r02a_timeline.log 2>&1; echo tl rc=$?
