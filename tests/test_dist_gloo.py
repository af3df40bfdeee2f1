"""world_size-2 gloo tests (CPU) of the batch-sharding host logic: shard bounds, the fused gradient all-reduce,
and that sharded gradients equal the single-process gradient.  The wave loop itself is CUDA-only, so a small
differentiable stand-in with the same signature (x, c, b, rho) -> [B,T,P] is used for the compute."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wavetorch_b200.distributed import shard_bounds, shard_batch, sync_grads


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_cover_everything():
    for n in (1, 7, 64, 65):
        for ws in (1, 2, 3, 8):
            spans = [shard_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _toy_loop(x, c, rho):
    # stand-in for the wave loop: depends on c and (nonlinearly) on rho
    field = torch.tanh(x[:, :, None, None] * c[None, None]) + 0.1 * rho[None, None] * x[:, :, None, None] ** 2
    return field.sum(dim=(2, 3))


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    B, T = 6, 5
    x = torch.randn(B, T)
    rho0 = torch.rand(4, 3)
    w = torch.randn(B, T)
    # single-process reference
    rho = rho0.clone().requires_grad_(True)
    c = 1.0 - 0.5 * rho
    (_toy_loop(x, c, rho) * w).sum().backward()
    full = rho.grad.clone()
    # sharded: each rank sees its slice, gradients are summed through sync_grads
    rho = rho0.clone().requires_grad_(True)
    c = 1.0 - 0.5 * rho
    cs, rs = sync_grads(c, rho)
    xl, wl = shard_batch(x), shard_batch(w)
    (_toy_loop(xl, cs, rs) * wl).sum().backward()
    ok = torch.allclose(rho.grad, full, atol=1e-6)
    # averaged variant
    rho2 = rho0.clone().requires_grad_(True)
    (c2,) = sync_grads(1.0 - 0.5 * rho2, average=True)
    (_toy_loop(xl, c2, torch.zeros(4, 3)) * wl).sum().backward()
    g = [torch.zeros_like(rho2.grad) for _ in range(world)]
    dist.all_gather(g, rho2.grad)
    same = all(torch.equal(g[0], gi) for gi in g)
    ret[rank] = (bool(ok), bool(same), xl.shape[0])
    dist.destroy_process_group()


def test_two_rank_gradient_sync_matches_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0], "sharded gradient differs from the single-process gradient"
    assert ret[0][1] and ret[1][1], "ranks ended with different gradients"
    assert ret[0][2] + ret[1][2] == 6


# ---- row-slab domain decomposition: host logic (the kernels and the exchange run in tests/test_gpu_parity.py) ----------
def test_slab_geometry_partitions_grid_sources_and_probes():
    from wavetorch_b200.domain import _SlabRank, memory_model, slab_rows
    Nx, Ny, halo = 100, 32, 8
    src = torch.tensor([[3, 5], [49, 7], [50, 9], [57, 1]], dtype=torch.int32)
    prb = torch.tensor([[10, 1], [50, 2], [42, 3], [99, 4]], dtype=torch.int32)
    sq = torch.tensor([1, 0, 1, 0], dtype=torch.int32)
    for world in (1, 2, 3, 4):
        slabs = [_SlabRank(r, world, Nx, Ny, halo, src, prb, sq, "cpu") for r in range(world)]
        assert slabs[0].r0 == 0 and slabs[-1].r1 == Nx and all(a.r1 == b.r0 for a, b in zip(slabs, slabs[1:]))
        for s in slabs:
            assert s.up == (halo if s.rank > 0 else 0) and s.dn == (halo if s.rank < world - 1 else 0)
            assert (s.r0, s.r1, s.e0, s.e1) == slab_rows(Nx, world, s.rank, halo)
            # local coordinates are relative to the first extended row
            assert (s.src_ext[:, 0] >= 0).all() and (s.src_ext[:, 0] < s.rows).all()
        # every source / probe is OWNED by exactly one slab, and seen (ghosts included) by at least that one
        assert sum(s.src_own.shape[0] for s in slabs) == src.shape[0]
        assert sum(int(s.prb_owned.sum()) for s in slabs) == prb.shape[0]
        owned = torch.cat([s.prb_ids[s.prb_owned] for s in slabs]).sort().values
        assert owned.tolist() == [0, 1, 2, 3]
    # memory model: more ranks -> thinner slabs; chunks bound state and tape, not the number of checkpoints
    m1 = memory_model(4096, 4096, 32, 10000, 1, 16, 128, batch_chunk=4)
    m8 = memory_model(4096, 4096, 32, 10000, 8, 16, 128)
    assert m8["rows"] == 512 + 32 and m1["rows"] == 4096
    assert m8["tape"] == 32 * 544 * 4096 * 4 * 128 and m8["segments"] == 79
    assert m8["total"] < 100e9 and m1["total"] < 150e9       # fits a 180 GB B200 either way


def _gather_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wavetorch_b200.domain import gather_row_slabs
    Nx, Ny = 11, 4                                    # ragged: 6 + 5 rows
    full = torch.arange(Nx * Ny, dtype=torch.float32).view(Nx, Ny)
    lo, hi = shard_bounds(Nx, world, rank)
    got = gather_row_slabs(full[lo:hi].clone(), Nx)
    ret[rank] = bool(torch.equal(got, full))
    dist.destroy_process_group()


def test_two_rank_slab_gather_assembles_the_full_field():
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_gather_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]
