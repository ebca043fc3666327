"""GPU experiment: C4-stress (optional): 3x3x3 replication of the CHA test cell (85 Å, 26,244 atoms),
0.10 Å grid (974x916x848 = 757 M points).  Times a 32-plane slab of the fused Na+ build with the image list prepared
for (a) that slab only, (b) the whole grid — the list every tile scans grows with the slab's bounding box."""
import dataclasses, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl
from ceg_b200.cell import inv3

inp, cset, ex = wl.load_fixture("cha_na")
m = np.asarray(inp.mat)
rep = [(i, j, k) for i in range(3) for j in range(3) for k in range(3)]
pos = np.concatenate([inp.positions + (m @ np.array(r, dtype=float)) for r in rep])
big = dataclasses.replace(inp, mat=3 * m, invmat=inv3(3 * m), safemin2=9 * inp.safemin2, positions=pos,
                          types=np.tile(inp.types, 27), charges=np.tile(inp.charges, 27))
cs = ceg.GridCoordinatesSetup.from_cell(3 * m, 0.10)
nx, ny, nz = cs.nodes
print(f"{len(pos)} atoms, grid {nx}x{ny}x{nz} = {nx*ny*nz/1e6:.0f} M points")
gs = ceg._lib.make_gridspec(cs)
lam_v, thr_v = ceg.vdw_scaling(); lam_c, thr_c = ceg.coulomb_scaling()
planes = 32
i0 = nx // 2
d = [torch.empty((8, planes, ny, nz), dtype=torch.float32, device="cuda") for _ in range(2)]
stream = torch.cuda.current_stream().cuda_stream
ps = ceg.ProbeSystem(None, None, inputs=big)
for _ in range(2):
    ceg._lib.grid_build_dev(ps.handle, gs, i0, i0 + planes, lam_v, thr_v, lam_c, thr_c, [t.data_ptr() for t in d], stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ceg._lib.grid_build_dev(ps.handle, gs, i0, i0 + planes, lam_v, thr_v, lam_c, thr_c, [t.data_ptr() for t in d], stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
st = ceg._lib.probe_last_stats(ps.handle)
pts = planes * ny * nz
print(f"(a) list of the 32-plane slab: {ms:.1f} ms, {pts/ms/1e3:.1f} M points/s, tested {st['pairs_tested']/st['n_points']:.0f} "
      f"in-cutoff {st['pairs_in_cutoff']/st['n_points']:.0f}; full grid would take {ms*nx/planes/1e3:.2f} s")
ck = float(d[1][0].double().sum())
# (b) the same planes computed as part of a 512-plane slab: one launch, the image list of the big slab
wide = 512
i0b = i0 - (wide - planes) // 2
del d
dd = [torch.empty((8, wide, ny, nz), dtype=torch.float32, device="cuda") for _ in range(2)]
ceg._lib.grid_build_dev(ps.handle, gs, i0b, i0b + wide, lam_v, thr_v, lam_c, thr_c, [t.data_ptr() for t in dd], stream)
torch.cuda.synchronize()
e0.record()
ceg._lib.grid_build_dev(ps.handle, gs, i0b, i0b + wide, lam_v, thr_v, lam_c, thr_c, [t.data_ptr() for t in dd], stream)
e1.record(); torch.cuda.synchronize()
msb = e0.elapsed_time(e1)
ptsb = wide * ny * nz
print(f"(b) one launch over a {wide}-plane slab: {msb:.1f} ms, {ptsb/msb/1e3:.1f} M points/s; full grid would take {msb*nx/wide/1e3:.2f} s")
off = i0 - i0b
assert abs(float(dd[1][0, off:off + planes].double().sum()) - ck) <= 1e-9 * abs(ck)
ps.close()
