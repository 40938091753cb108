cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r02_gputests.log 2>&1; echo tests rc=$?; tail -2 gpurun_out/r02_gputests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo bench rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo ref rc=$?
SKIP=70 TAG=r02 bash tools/_run_profiles.sh | head -2
