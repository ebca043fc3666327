"""GPU experiment (torchrun, N ranks): the device-to-host ceiling of this box for the build's write-back.
Every rank copies slabs of the size it writes in the C4 end-to-end step (1.8 GB / N, as 16 runs) from HBM into
pinned host memory, all ranks at once; reports the aggregate GB/s for plain pinned and write-combined buffers.  That is
the ceiling the `e2e` figure of bench.py --gpus N (grids landed in host memory) must be judged against.
usage: torchrun --nproc-per-node N tools/d2h_ceiling.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
from cuda.bindings import runtime as rt

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bind = "--bind" in sys.argv
if bind:
    import bench
    bench.bind_to_gpu_cpus(local)

total = 2 * 8 * 326 * 306 * 284 * 4          # the C4 grid pair, bytes
mine = total // world // 16 * 16
runs = 16
src = torch.empty(mine, dtype=torch.uint8, device="cuda")
src.fill_(1)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


res = {}
for name, flags in (("pinned", rt.cudaHostAllocDefault), ("write-combined", rt.cudaHostAllocWriteCombined)):
    err, hptr = rt.cudaHostAlloc(mine, flags)
    assert err == rt.cudaError_t.cudaSuccess, err
    stream = torch.cuda.current_stream().cuda_stream
    per = mine // runs

    def step():
        for r in range(runs):
            rt.cudaMemcpyAsync(hptr + r * per, src.data_ptr() + r * per, per, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost, stream)

    for _ in range(2):
        step()
    sync()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        step()
    sync()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name] = (float(t.item()), mine * world / float(t.item()) / 1e9)
    rt.cudaFreeHost(hptr)
if rank == 0:
    print(f"D2H ceiling, {world} GPU(s){' (ranks bound to their GPU-local CPUs)' if bind else ''}: {mine * world / 1e9:.2f} GB per step; " +
          "; ".join(f"{k}: {v[0] * 1e3:.2f} ms = {v[1]:.1f} GB/s aggregate" for k, v in res.items()))
if world > 1:
    dist.destroy_process_group()
