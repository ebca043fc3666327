cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_n8d.log
nvidia-smi -L | wc -l > $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus 8 --steps 8 --warmup 3 --no-c5 > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err
echo "bench n=8 rc $?" >> $L
grep -v "^$" gpurun_out/r2l_bench_n8.err | tail -4 >> $L
