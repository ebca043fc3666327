"""GPU experiment (torchrun): host-landing build of this rank's x-slab of C4 into registered host memory, phases per rank."""
import os, sys, time, datetime
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceg_b200 as ceg
from ceg_b200 import workloads as wl, sharding
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
inp, cset, ex = wl.load_fixture("cha_na")
cs = wl.respace(cset, 0.10)
nx, ny, nz = cs.nodes
a, b = sharding.slab_of_rank(nx, world, rank)
gs = ceg._lib.make_gridspec(cs)
lam = ceg.vdw_scaling() + ceg.coulomb_scaling()
host = [np.empty((8, nx, ny, nz), dtype=np.float32) for _ in range(2)]
for h in host:
    for c in range(8):
        ceg._lib.host_register(h[c, a:b])
ps = ceg.ProbeSystem(None, None, inputs=inp, device=local)
for it in range(6):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ceg._lib.grid_build(ps.handle, gs, a, b, *lam, host, None)
    ms = (time.perf_counter() - t0) * 1e3
    t = ceg._lib.probe_last_timing(ps.handle)
    if it >= 2:
        print(f"rank {rank} it {it}: call {ms:.2f} ms  span {t['kernel_span_ms']:.2f}  tail {t['copy_tail_ms']:.2f}", flush=True)
for h in host:
    for c in range(8):
        ceg._lib.host_unregister(h[c, a:b])
ps.close()
if world > 1:
    dist.destroy_process_group()
