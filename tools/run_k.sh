cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_k.log
timeout 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_device_resident.py -x -q 2>&1 | tail -3 > $L
timeout 300 python tools/exp_build.py 0.10 5 >> $L 2>&1
timeout 300 python tools/exp_build.py 0.15 5 >> $L 2>&1
timeout 300 python tools/exp_configs.py >> $L 2>&1
