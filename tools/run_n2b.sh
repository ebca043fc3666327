cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_n2b.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $L
for p in 4 2; do
CEG_BENCH_PIECES=$p timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$p bench.py --gpus 2 --steps 8 --warmup 3 --no-c5 > gpurun_out/r2c_bench_n2_p$p.json 2> gpurun_out/r2c_bench_n2_p$p.err
echo "bench pieces=$p rc $?" >> $L
grep -v "^$" gpurun_out/r2c_bench_n2_p$p.err | tail -5 >> $L
done
