cd $GRAFT_REPO_ROOT
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_encoder.py -q -m gpu -x -k "matches_torch_fp64 and (3-7 or 5-50)" > gpurun_out/race_enc.log 2>&1; echo enc rc=$?; grep -i "hazard" gpurun_out/race_enc.log | head -5; tail -3 gpurun_out/race_enc.log
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_tc.py tests/test_gpu_encoder.py -q -m gpu -x > gpurun_out/sync_tc.log 2>&1; echo sync rc=$?; tail -3 gpurun_out/sync_tc.log
