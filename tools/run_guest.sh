cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_guest5.log
timeout 600 python -m pytest tests/test_gpu_guest.py -x -q 2>&1 | tail -15 > $L
timeout 300 python tools/exp_guest.py >> $L 2>&1
timeout 300 python tools/exp_guest.py >> $L 2>&1
