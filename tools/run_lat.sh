cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_lat3.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $L
timeout 200 python - >> $L 2>&1 <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
import ceg_b200 as ceg
from ceg_b200 import workloads as wl
inp, cset, ex = wl.load_fixture("cha_na")
cs = wl.respace(cset, 0.3)
nx, ny, nz = cs.nodes
ps = ceg.ProbeSystem(None, None, inputs=inp)
outs = [np.zeros((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
ceg._lib.grid_build(ps.handle, ceg._lib.make_gridspec(cs), 0, nx, *(ceg.vdw_scaling() + ceg.coulomb_scaling()), outs, None)
ps.close()
for g, prec in enumerate((np.inf, 1e-6)):
    eg = ceg.EnergyGrid(cs, (1, 1, 1), prec, True, (outs[g].astype(np.float64) * ceg.GRID_TO_KELVIN).astype(np.float32))
    pos = wl.random_positions(cs.cell.mat, 5000, 3)
    big = ceg._lib.interpolate(eg.handle, pos)                       # batch kernels
    for n in (1, 7, 256):
        small = ceg._lib.interpolate(eg.handle, pos[100:100 + n])    # latency kernel
        assert np.array_equal(small, big[100:100 + n]), (g, n)
    one = pos[:1]
    for _ in range(50): ceg._lib.interpolate(eg.handle, one)
    ts = []
    for _ in range(500):
        t0 = time.perf_counter(); ceg._lib.interpolate(eg.handle, one); ts.append(time.perf_counter() - t0)
    print("grid", g, "interpolate 1 point:", round(float(np.median(ts)) * 1e6, 1), "us; small == batch bitwise")
    eg.close()
PY
