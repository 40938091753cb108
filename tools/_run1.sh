cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_all.log 2>&1; grep -E "^E  .*(rel|vs)|passed|failed|FAILED" gpurun_out/t_all.log | head -40
