"""world_size-2 gloo tests (CPU) of the data-parallel host logic in supernormal_b200/dp.py: live-range gradient
all-reduce, per-rank patch streams, slab-sharded mesh gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for port in [_free_port()] for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        return dict(ret)


def _grad_case(rank, world):
    from supernormal_b200 import dp
    offsets = [0, 100, 300, 700]
    n_live = dp.live_numel(16, offsets, 2)          # 16 + 2*300
    g = torch.full((16 + 2 * 700,), float(rank + 1))
    g[n_live:] = 0.0                                # inactive levels: exactly zero on every rank
    scale = dp.allreduce_live_gradients(g, n_live, world)
    return g.tolist(), n_live, scale, dp.rank_seed(5, rank)


def test_live_gradient_allreduce():
    out = _run(_grad_case)
    for r in (0, 1):
        g, n_live, scale, seed = out[r]
        assert n_live == 616 and scale == 0.5
        assert all(v == 3.0 for v in g[:n_live]) and all(v == 0.0 for v in g[n_live:])
    assert out[0][3] != out[1][3]


def _mesh_case(rank, world):
    from oracle import mc
    from supernormal_b200 import dp
    res = 30
    g = np.linspace(-1, 1, res).astype(np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    rng = np.random.RandomState(3)
    u = (0.6 - np.sqrt(X ** 2 + Y ** 2 + Z ** 2) + 0.03 * rng.randn(res, res, res)).astype(np.float32)
    c0, c1 = dp.slab_cells(res, rank, world)
    v, t, n_main, _ = mc.marching_cubes(u[c0:c1 + 1], 0.0, x_offset=c0)
    merged = dp.gather_slab_meshes(torch.from_numpy(v), torch.from_numpy(t).long(), n_main)
    if rank != 0:
        return merged is None
    vm, tm = merged
    vw, tw, _, _ = mc.marching_cubes(u)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_mesh_oracle import canonical_triangles
    same = np.array_equal(canonical_triangles(vm.numpy(), tm.numpy()), canonical_triangles(vw, tw))
    return same and vm.shape[0] == vw.shape[0] and mc.mesh_checks(vm.numpy(), tm.numpy().astype(np.int32))[:3] == mc.mesh_checks(vw, tw)[:3]


@pytest.mark.parametrize("world", [2, 3])
def test_slab_mesh_gather(world):
    out = _run(_mesh_case, world)
    assert all(out[r] is True for r in range(world))
