cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_n2f.log
: > $L
for pol in 0 1 2 0 1; do
echo "== CEG_CHUNKS=$pol" >> $L
CEG_CHUNKS=$pol timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$pol tools/exp_e2e_n.py 2>&1 | grep "^rank" | sort >> $L
done
