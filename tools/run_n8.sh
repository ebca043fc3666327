cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_n8.log
nvidia-smi -L > $L
nvidia-smi topo -m >> $L 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|NUMA" >> $L
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/d2h_ceiling.py 2>&1 | grep "D2H ceiling" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n tools/d2h_ceiling.py --bind 2>&1 | grep "D2H ceiling" >> $L
done
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
echo "bench n=$n rc $?" >> $L
done
