"""GPU experiment: build-kernel timing on the other BASELINE configs (C3 MFI-shaped multi-probe, C5 MOF-shaped)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl

def run(name, inp, cs, reps=3):
    nx, ny, nz = cs.nodes
    ng = inp.table.shape[0] + int(inp.want_coulomb)
    ps = ceg.ProbeSystem(None, None, inputs=inp)
    gs = ceg._lib.make_gridspec(cs)
    lam_v, thr_v = ceg.vdw_scaling(); lam_c, thr_c = ceg.coulomb_scaling()
    # slab of at most ~24 M points to bound memory
    planes = max(8, min(nx, (24_000_000 // (ny * nz)) // 8 * 8))
    d = [torch.empty((8, planes, ny, nz), dtype=torch.float32, device="cuda") for _ in range(ng)]
    ptrs = [t.data_ptr() for t in d]
    stream = torch.cuda.current_stream().cuda_stream
    i0 = (nx - planes) // 2
    ceg._lib.grid_build_dev(ps.handle, gs, i0, i0 + planes, lam_v, thr_v, lam_c, thr_c, ptrs, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ceg._lib.grid_build_dev(ps.handle, gs, i0, i0 + planes, lam_v, thr_v, lam_c, thr_c, ptrs, stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = ceg._lib.probe_last_stats(ps.handle)
    pts = planes * ny * nz
    print(f"{name}: grid {nx}x{ny}x{nz} ({nx*ny*nz/1e6:.1f} M pts), {ng} grids, N_sc {len(inp.positions)}, slab {planes} planes: "
          f"{ms:.2f} ms, {pts/ms/1e3:.1f} M points/s, full grid {ms*nx/planes:.1f} ms, in-cutoff {st['pairs_in_cutoff']/st['n_points']:.0f} tested {st['pairs_tested']/st['n_points']:.0f}")
    ps.close()

t0 = time.time()
inp, cs, fw = wl.synthetic_inputs("mfi", 0.15, ("C_ch4", "C_co2", "O_co2"), True)
run("C3 MFI 0.15 (3 LJ probes + Coulomb)", inp, cs)
run("C3 MFI 0.10 (3 LJ probes + Coulomb)", inp, wl.respace(cs, 0.10))
inp1, cs1, _ = wl.synthetic_inputs("mfi", 0.15, ("C_ch4",), False)
run("C3 MFI 0.15 (CH4 only)", inp1, cs1)
inp2, cs2, _ = wl.synthetic_inputs("mfi", 0.15, ("C_co2", "O_co2"), True)
run("C3 MFI 0.15 (CO2: 2 LJ + Coulomb)", inp2, cs2)
print("mof generation...", flush=True)
inp5, cs5, _ = wl.synthetic_inputs("mof", 0.10, ("C_ch4",), True)
print(f"generated in {time.time()-t0:.0f}s")
run("C5 MOF 0.10 (CH4 LJ + Coulomb)", inp5, cs5)
inpa, csa, _ = wl.load_fixture("chana_ar")
run("Ar in CHA+Na 0.15 (shifted LJ only)", inpa, csa)
