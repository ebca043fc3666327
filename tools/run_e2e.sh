cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_e2e.log
timeout 900 python -m pytest tests/test_gpu_build.py -x -q 2>&1 | tail -3 > $L
timeout 300 python tools/exp_e2e.py 2>&1 | tail -9 >> $L
