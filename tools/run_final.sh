cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_final.log
nvidia-smi -L > $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "smoke rc $?" >> $L
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
echo "bench rc $?" >> $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref_n1.json 2>> $L
echo "ref rc $?" >> $L
