cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -150 ) > gpurun_out/r2_run7.log 2>&1
