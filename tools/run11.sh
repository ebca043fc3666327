cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25 ) > gpurun_out/r2_run11.log 2>&1
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
echo "bench rc $?" >> gpurun_out/r2_run11.log
grep -v "^$" gpurun_out/r2_bench11.err | tail -6 >> gpurun_out/r2_run11.log
