set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_run3.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $L 2>&1
for v in d0 d1 d2; do
  export CEG_LIB_PATH=$PWD/tools/variants/$v.so
  echo "== $v" >> $L
  timeout 300 python tools/exp_build.py 0.10 5 >> $L 2>&1
  timeout 300 python tools/exp_build.py 0.15 5 >> $L 2>&1
done
unset CEG_LIB_PATH
echo "== configs cur" >> $L
timeout 600 python tools/exp_configs.py 2>&1 | grep -v generat >> $L
bash tools/run_ncu_build.sh b9
