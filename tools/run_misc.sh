cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_misc.log
timeout 300 python bench.py --headline interp --steps 3 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/r2m_interp.json 2> gpurun_out/r2m_interp.err
echo "headline interp rc $?" > $L
python -c "
import json
d=json.loads(open('gpurun_out/r2m_interp.json').read().strip().splitlines()[-1])
print(d['metric'], d['value'], d['unit'], d['roofline']['frac'], d['e2e']['value'], list(d.keys())[:12])" >> $L 2>&1
timeout 300 python examples/mfi_co2_trials.py >> $L 2>&1
echo "example rc $?" >> $L
