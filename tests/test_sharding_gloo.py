"""N > 1 host path on CPU: two gloo ranks each build their x-slab (with the CPU oracle standing in for
the device), rank 0 assembles the grid from the slabs exactly as the multi-GPU driver does, and the
result must be bit-identical to the unsharded build.  Also the max-over-ranks timing reduction."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ceg_b200 as ceg
    import oracle as orc
    from ceg_b200 import sharding, workloads as wl
    inp, cset, _ = wl.load_fixture("chana_ar")
    cs = wl.respace(cset, 2.0)
    nx, ny, nz = cs.nodes
    a, b = sharding.slab_of_rank(nx, world, rank)
    outs, _ = orc.grid_build(inp, cs, a, b, 1, want_f32=True)
    slab = torch.from_numpy(np.ascontiguousarray(outs[0][:, a:b]))
    # slabs go to rank 0 (in production each rank writes its runs of the .grid file / pinned host array)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((a, slab.numpy()), gathered, dst=0)
    # max-over-ranks reduction used for every multi-GPU timing
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == 10.0 + world - 1
    # query sharding covers every index exactly once
    q = torch.zeros(1001, dtype=torch.int64)
    qa, qb = sharding.split_range(1001, world)[rank]
    q[qa:qb] += 1
    dist.all_reduce(q)
    assert bool((q == 1).all())
    if rank == 0:
        full = np.zeros((8, nx, ny, nz), dtype=np.float32)
        for i0, s in gathered:
            sharding.place_slab(full, s, i0)
        ref, _ = orc.grid_build(inp, cs, 0, nx, 1, want_f32=True)
        assert np.array_equal(full.view(np.uint32), ref[0].view(np.uint32))
        runs = sharding.slab_file_runs((nx, ny, nz), a, b)
        assert len(runs) == 8 and runs[1][0] - runs[0][0] == nx * ny * nz * 4
        open(os.path.join(outdir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slab_build_gloo(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_split_range():
    sys.path.insert(0, ROOT)
    from ceg_b200 import sharding
    for n in (1, 7, 218, 326):
        for p in (1, 2, 4, 8):
            r = sharding.split_range(n, p)
            assert r[0][0] == 0 and r[-1][1] == n and all(x[1] == y[0] for x, y in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
