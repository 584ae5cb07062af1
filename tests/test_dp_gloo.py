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


# ---- peer-memory step tail (snb_train_tail_peer): static chunk ownership, sharded Adam, broadcast -- host-side restatement ----
def test_peer_chunk_ownership_partitions_live_range():
    from supernormal_b200 import dp
    for world in (1, 2, 3, 8):
        for n_live in (0, 4, 4096, 4100, 3 * 4096 + 8, 1_000_000, 23_744_000):
            assert sum(dp.owned_floats(n_live, r, world) for r in range(world)) == n_live
        # static: the owner of a chunk does not depend on how many levels are live
        assert [dp.chunk_owner(c, world) for c in range(2 * world)] == list(range(world)) * 2
    for world in (2, 3):   # owner_mask is the elementwise form of owned_floats
        masks = [dp.owner_mask(3 * 4096 + 8, r, world) for r in range(world)]
        assert torch.stack(masks).sum(0).eq(1).all()
        assert [int(mk.sum()) for mk in masks] == [dp.owned_floats(3 * 4096 + 8, r, world) for r in range(world)]
    offs, total = dp.carve_layout([10, 4096, 1, 128])
    assert offs == [0, 256, 4352, 4608] and total == 4864 and all(o % 256 == 0 for o in offs)


def _adam(p, g, m, v, lr, t, scale):
    g = g * scale
    m = 0.9 * m + 0.1 * g
    v = 0.999 * v + 0.001 * g * g
    bc1, bc2 = 1 - 0.9 ** t, 1 - 0.999 ** t
    return p - lr / bc1 * m / (v.sqrt() / bc2 ** 0.5 + 1e-8), m, v


def _peer_tail_case(rank, world):
    """What the peer kernel does, with torch collectives standing in for peer loads / stores: the owner of each 4096-float
    chunk sums that chunk's gradients over the ranks in rank order, runs Adam on ITS shard of the state and broadcasts the
    new parameters; compared with all-reduce + replicated Adam."""
    from supernormal_b200 import dp
    n_live, n = 3 * dp.PEER_CHUNK_FLOATS + 1000, 5 * dp.PEER_CHUNK_FLOATS
    gen = torch.Generator().manual_seed(7)
    p0 = torch.randn(n, generator=gen, dtype=torch.float64)
    m0, v0 = torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    p_ref, m_ref, v_ref = p0.clone(), m0.clone(), v0.clone()
    p, m, v = p0.clone(), m0.clone(), v0.clone()          # m, v: only the owned chunks are ever touched on this rank
    chunk = torch.arange(n) // dp.PEER_CHUNK_FLOATS
    mine = (chunk % world == rank) & (torch.arange(n) < n_live)
    for t in (1, 2, 3):
        g = torch.randn(n, generator=torch.Generator().manual_seed(100 * t + rank), dtype=torch.float64)
        g[n_live:] = 0.0
        # reference path: all-reduce, every rank repeats the whole Adam sweep
        gs = g.clone()
        dist.all_reduce(gs)
        pr, mr, vr = _adam(p_ref[:n_live], gs[:n_live], m_ref[:n_live], v_ref[:n_live], 1e-2, t, 1.0 / world)
        p_ref[:n_live], m_ref[:n_live], v_ref[:n_live] = pr, mr, vr
        # peer path: "peer loads" of every rank's gradient, rank-order sum on the owner, Adam on the shard, broadcast
        all_g = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(all_g, g)
        gsum = all_g[0].clone()
        for r in range(1, world):
            gsum += all_g[r]
        pn, mn, vn = _adam(p[mine], gsum[mine], m[mine], v[mine], 1e-2, t, 1.0 / world)
        m[mine], v[mine] = mn, vn
        contrib = torch.zeros_like(p)
        contrib[mine] = pn
        dist.all_reduce(contrib)                           # every element of the live range has exactly one owner
        p[:n_live] = contrib[:n_live]
    return p.tolist(), p_ref.tolist(), int(mine.sum())


@pytest.mark.parametrize("world", [2, 3])
def test_peer_tail_ownership_equals_allreduce_adam(world):
    from supernormal_b200 import dp
    out = _run(_peer_tail_case, world)
    n_live = 3 * dp.PEER_CHUNK_FLOATS + 1000
    assert sum(out[r][2] for r in range(world)) == n_live
    for r in range(world):
        assert out[r][0] == out[0][0]                      # replicas bit-identical
        np.testing.assert_allclose(out[r][0], out[r][1], rtol=1e-12, atol=1e-12)   # == all-reduce + replicated Adam
    assert out[0][0][n_live:] == out[0][1][n_live:]       # inactive levels untouched
