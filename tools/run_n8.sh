cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_n8c.log
nvidia-smi -L | wc -l > $L
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2e_bench_n$n.json 2> gpurun_out/r2e_bench_n$n.err
echo "bench n=$n rc $?" >> $L
grep -v "^$" gpurun_out/r2e_bench_n$n.err | tail -6 >> $L
done
