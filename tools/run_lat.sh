cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_lat2.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $L
timeout 600 python bench.py --steps 3 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
echo "bench rc $?" >> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d['latency_us'].items() if not isinstance(v,str)})
print(d['guest']['kernel_ms'], d['guest']['ms_per_step'], d['guest']['roofline']['frac'])
PY
