set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_run1.log 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) >> gpurun_out/r2_run1.log 2>&1
for v in base nopipe cur; do
  if [ $v = cur ]; then unset CEG_LIB_PATH; else export CEG_LIB_PATH=$PWD/tools/variants/$v.so; fi
  echo "== $v" >> gpurun_out/r2_run1.log
  timeout 300 python tools/exp_build.py 0.10 5 >> gpurun_out/r2_run1.log 2>&1
  timeout 300 python tools/exp_build.py 0.15 5 >> gpurun_out/r2_run1.log 2>&1
done
unset CEG_LIB_PATH
echo "== configs base" >> gpurun_out/r2_run1.log
CEG_LIB_PATH=$PWD/tools/variants/base.so timeout 600 python tools/exp_configs.py >> gpurun_out/r2_run1.log 2>&1
echo "== configs cur" >> gpurun_out/r2_run1.log
timeout 600 python tools/exp_configs.py >> gpurun_out/r2_run1.log 2>&1
