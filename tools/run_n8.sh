cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_n8b.log
nvidia-smi -L | wc -l > $L
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2b_bench_n$n.json 2> gpurun_out/r2b_bench_n$n.err
echo "bench n=$n rc $?" >> $L
done
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
echo "bench n=1 rc $?" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref_n8.json 2>> $L
