cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ewald.py -x -q 2>&1 | tail -12 > gpurun_out/r2_tests.log
timeout 200 python - >> gpurun_out/r2_tests.log 2>&1 <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import ceg_b200 as ceg
from ceg_b200 import workloads as wl
sys.path.insert(0, 'tests')
from test_gpu_ewald import _framework_from_fixture
inp, cset, ex = wl.load_fixture("cha_na")
fw = _framework_from_fixture(ex)
ef = ceg.initialize_ewald(fw)
rng = np.random.default_rng(0)
for na in (3, 4, 8, 16):
    q = rng.standard_normal(na); offs = rng.standard_normal((na, 3))
    pos = wl.random_positions(fw.mat, 200000, 7)[:, None, :] + offs[None]
    ceg._lib.ewald_reciprocal(ef.handle, pos, q, 0.0, 0.0)
    t0 = time.perf_counter()
    for _ in range(3): ceg._lib.ewald_reciprocal(ef.handle, pos, q, 0.0, 0.0)
    print(f"{na} atoms, 2e5 placements: {(time.perf_counter()-t0)/3*1e3:.2f} ms per call (host arrays in/out)")
PY
