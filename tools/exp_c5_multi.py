"""GPU experiment (torchrun, N ranks): BASELINE config C5 end to end on the devices —
  1. every rank builds its x-slab of the 502^3 LJ + Coulomb grid pair (ceg_grid_build_dev, no exchange),
  2. the slabs are all-gathered over NVLink (NCCL) so that every rank holds both full grids (2 x 4.05 GB),
  3. each rank turns them into interpolation handles on the device (ceg_grid_upload_dev, parse_grid's scaling),
  4. the 10^9-query sweep is sharded by query index (ceg_interpolate_sweep) and the partial results are summed.
usage: torchrun --nproc-per-node N tools/exp_c5_multi.py [n_queries] [spacing]"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl
from ceg_b200.sharding import split_range

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
n_q = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
spacing = float(sys.argv[2]) if len(sys.argv) > 2 else 0.10

def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

inp, cs, _ = wl.synthetic_inputs("mof", spacing, ("C_ch4",), True)
nx, ny, nz = cs.nodes
gs = ceg._lib.make_gridspec(cs)
lam_v, thr_v = ceg.vdw_scaling(); lam_c, thr_c = ceg.coulomb_scaling()
slabs = split_range(nx, world)
a, b = slabs[rank]
pmax = max(e - s for s, e in slabs)
stream = torch.cuda.current_stream().cuda_stream
ps = ceg.ProbeSystem(None, None, inputs=inp, device=local)
mine = [torch.zeros((8, pmax, ny, nz), dtype=torch.float32, device="cuda") for _ in range(2)]   # padded to the widest slab
views = [m[:, :b - a].contiguous() if (b - a) != pmax else m for m in mine]
sync(); t0 = time.perf_counter()
ceg._lib.grid_build_dev(ps.handle, gs, a, b, lam_v, thr_v, lam_c, thr_c, [v.data_ptr() for v in views], stream)
sync(); t_build = time.perf_counter() - t0
for m, v in zip(mine, views):
    if v is not m:
        m[:, :b - a] = v
full = [torch.empty((8, nx, ny, nz), dtype=torch.float32, device="cuda") for _ in range(2)]
sync(); t0 = time.perf_counter()
for g in range(2):
    if world > 1:
        parts = [torch.empty_like(mine[g]) for _ in range(world)]
        dist.all_gather(parts, mine[g])
    else:
        parts = [mine[g]]
    for (s, e), p in zip(slabs, parts):
        full[g][:, s:e] = p[:, :e - s]
sync(); t_gather = time.perf_counter() - t0
t0 = time.perf_counter()
grids = [ceg.EnergyGrid.from_device(cs, inp.supercell, prec, full[g].data_ptr(), device=local)
         for g, prec in enumerate((np.inf, 1e-6))]
sync(); t_upload = time.perf_counter() - t0
del full, mine, views
q0, q1 = split_range(n_q, world)[rank]
for eg in grids:
    ceg.interpolate_sweep(eg, 10**6, 4)
sync(); t0 = time.perf_counter()
res = [ceg.interpolate_sweep(eg, q1 - q0, 4, first=q0, n_keep=1000 if rank == 0 else 0) for eg in grids]
sync(); t_sweep = time.perf_counter() - t0
tot = torch.tensor([[r["sum"], float(r["n_blocked"])] for r in res], dtype=torch.float64, device="cuda")
mn = torch.tensor([r["min"] for r in res], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(tot); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
if rank == 0:
    pts = nx * ny * nz
    print(f"C5 on {world} GPU(s): grid {nx}x{ny}x{nz} = {pts/1e6:.1f} M points x 2 grids; build {t_build*1e3:.1f} ms "
          f"({pts/t_build/1e6:.0f} M points/s), all-gather {t_gather*1e3:.1f} ms ({2*8*pts*4/t_gather/1e9:.0f} GB/s of grid data), "
          f"device upload {t_upload*1e3:.1f} ms, sweep of {n_q:.0e} queries x 2 grids {t_sweep*1e3:.1f} ms "
          f"({2*n_q/t_sweep/1e9:.1f} G interpolations/s)")
    print(f"  VdW grid: sum {tot[0,0].item():.9e} min {mn[0].item():.6f} blocked {int(tot[0,1].item())}; "
          f"Coulomb grid: sum {tot[1,0].item():.9e} min {mn[1].item():.6f}")
    # spot check of the first queries against the single-call path on this rank's replica
    pos = wl.random_positions(cs.cell.mat, 1000, 4)
    assert np.array_equal(ceg.interpolate_grid(grids[0], pos), res[0]["keep"])
for eg in grids:
    eg.close()
ps.close()
if world > 1:
    dist.destroy_process_group()
