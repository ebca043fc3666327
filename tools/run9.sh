cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 ) > gpurun_out/r2_run9.log 2>&1
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
echo "bench rc $?" >> gpurun_out/r2_run9.log
grep -v "^$" gpurun_out/r2_bench9.err | tail -8 >> gpurun_out/r2_run9.log
