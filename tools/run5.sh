set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_run5.log
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > $L 2>&1
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
echo "bench rc $?" >> $L
tail -20 gpurun_out/r2_bench5.err >> $L
