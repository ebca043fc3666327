cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_fused.log
nvidia-smi -L | wc -l > $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $L
timeout 300 python tools/exp_build.py 0.10 5 >> $L 2>&1
timeout 300 python tools/exp_configs.py >> $L 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 --no-c5 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
echo "bench n=2 rc $?" >> $L
grep -v "^$" gpurun_out/r2d_bench_n2.err | tail -8 >> $L
