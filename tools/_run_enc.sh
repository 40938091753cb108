cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_encoder.py -q -m gpu --tb=short > gpurun_out/t_enc.log 2>&1; echo rc=$?; tail -25 gpurun_out/t_enc.log
