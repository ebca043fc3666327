set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_run2.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $L 2>&1
for v in cur pipe1 pipe0 nosplit; do
  if [ $v = cur ]; then unset CEG_LIB_PATH; else export CEG_LIB_PATH=$PWD/tools/variants/$v.so; fi
  echo "== $v" >> $L
  timeout 300 python tools/exp_build.py 0.10 5 >> $L 2>&1
  timeout 300 python tools/exp_build.py 0.15 5 >> $L 2>&1
  timeout 600 python tools/exp_configs.py 2>&1 | grep -v generat >> $L
done
