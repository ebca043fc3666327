set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_run6.log
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $L 2>&1
timeout 300 python tools/exp_pair.py >> $L 2>&1
timeout 900 python bench.py --steps 4 --warmup 3 --no-c5 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
echo "bench rc $?" >> $L
tail -5 gpurun_out/r2_bench6.err >> $L
