cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_n2d.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $L
CEG_BENCH_NO_FUSED=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_n2_nofused.json 2> gpurun_out/r2i_bench_n2_nofused.err
echo "bench nofused rc $?" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err
echo "bench rc $?" >> $L
tail -3 gpurun_out/r2i_bench_n2_nofused.err >> $L
