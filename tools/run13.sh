cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_guest.py tests/test_gpu_setup_raspa.py -m gpu -q --tb=short 2>&1 | tail -8 ) > gpurun_out/r2_run13.log 2>&1
python tools/exp_guest.py >> gpurun_out/r2_run13.log 2>&1
