cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_lat4.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $L
timeout 200 python examples/mfi_co2_trials.py >> $L 2>&1
timeout 200 python - >> $L 2>&1 <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import ceg_b200 as ceg
from ceg_b200 import workloads as wl
from test_gpu_device_resident import _two_grids
inp, cs, ex, outs, gK = _two_grids(0.5)
n_fw = len(ex["site_charge"])
fw = ceg.RASPASystem(ex["unit_mat"], ex["positions"][:n_fw], ["X"] * n_fw, np.ones(n_fw), ex["charges"][:n_fw], False)
q = float(ex["probe_charge"])
eg_v = ceg.EnergyGrid(cs, (1, 1, 1), np.inf, True, gK[0]); eg_c = ceg.EnergyGrid(cs, (1, 1, 1), 1e-6, True, gK[1])
ef = ceg.initialize_ewald(fw)
mol = ceg.RASPASystem(np.diag([np.inf] * 3), np.zeros((1, 3)), ["Na"], np.ones(1), np.array([q]), True)
setup = ceg.CrystalEnergySetup(fw, mol, eg_c, np.array([q]), [eg_v], [0], ef, None, ceg.blockfile_from_grid(eg_v))
one = wl.random_positions(fw.mat, 1, 3)[:, None, :]
for name, fn in (("energy_points 1 placement", lambda: ceg.energy_points(setup, one)),
                 ("framework_interactions 1 placement", lambda: ceg.energy_points(setup, one, reciprocal=False)),
                 ("energy_point", lambda: ceg.energy_point(setup, one[0]))):
    for _ in range(50): fn()
    ts = []
    for _ in range(1000):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print(name, round(float(np.median(ts)) * 1e6, 1), "us")
PY
