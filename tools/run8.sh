cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v "^$" | tail -40 ) > gpurun_out/r2_run8.log 2>&1
timeout 900 python bench.py --steps 4 --warmup 3 --no-c5 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
echo "bench rc $?" >> gpurun_out/r2_run8.log
tail -5 gpurun_out/r2_bench8.err >> gpurun_out/r2_run8.log
