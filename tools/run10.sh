cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_run10.log
( timeout 900 python -m pytest tests/test_gpu_build.py tests/test_gpu_device_resident.py tests/test_gpu_setup_raspa.py -m gpu -q --tb=short -s 2>&1 | grep -v "^$" | tail -12 ) > $L 2>&1
for v in nopw pw; do
  export CEG_LIB_PATH=$PWD/tools/variants/$v.so
  echo "== $v" >> $L
  timeout 300 python tools/exp_build.py 0.10 5 >> $L 2>&1
  timeout 300 python tools/exp_build.py 0.15 5 >> $L 2>&1
  timeout 600 python tools/exp_configs.py 2>&1 | grep "C5\|C3 MFI 0.15 (3" >> $L
done
