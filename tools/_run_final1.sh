cd $GRAFT_REPO_ROOT
SKIP=70 TAG=r02 bash tools/_run_profiles.sh | tail -3
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo bench rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo ref rc=$?
timeout 600 python tools/bf16_parity_report.py > gpurun_out/r02_bf16_parity.txt 2>&1; echo parity rc=$?; tail -5 gpurun_out/r02_bf16_parity.txt
