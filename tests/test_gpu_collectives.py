"""The two hand-written NVLink kernels -- wt_peer_allreduce (gradient all-reduce of the batch-sharded path) and the
ghost-row exchange of the slab decomposition -- exercised on ONE GPU: the "ranks" are streams of this process and the
"peer" pointers are plain device pointers, so the kernels, their flag / epoch protocols and their CUDA-graph replay run
exactly as they do across GPUs (tests/test_multi_gpu.py repeats them over real peers when the box has >= 2 GPUs)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

from wavetorch_b200 import _lib  # noqa: E402

DEV = "cuda"


class _VirtualPeers:
    """`world` ranks of wt_peer_allreduce inside one process: one exchange buffer, one state word pair and one stream each."""

    def __init__(self, world, nmax):
        self.world, self.nmax = world, nmax
        gather = 2 * world * nmax
        self.flags_off = ((gather * 4 + 255) // 256) * 256
        self.bufs = [torch.zeros(self.flags_off // 4 + 64, dtype=torch.float32, device=DEV) for _ in range(world)]
        self.base = (ctypes.c_uint64 * world)(*[b.data_ptr() for b in self.bufs])
        self.state = [torch.zeros(2, dtype=torch.int32, device=DEV) for _ in range(world)]
        self.streams = [torch.cuda.Stream() for _ in range(world)]

    def launch(self, rank, src, out, scale):
        lib = _lib.load()
        with torch.cuda.stream(self.streams[rank]):
            st = lib.wt_peer_allreduce(self.world, rank, src.numel(), self.nmax, scale, _lib.ptr(src), _lib.ptr(out), self.base,
                                       self.flags_off, _lib.ptr(self.state[rank]), torch.cuda.current_device(),
                                       _lib.stream_ptr(torch.device(DEV)))
        _lib.check(st, "wt_peer_allreduce")


@pytest.mark.parametrize("world,n", [(2, 15000), (4, 30000), (8, 1000)])
def test_peer_allreduce_kernel_many_calls(world, n):
    """60 back-to-back calls (epoch parity, double buffering) with fresh data each time: every rank gets the rank-ordered
    float32 sum, bitwise, and all ranks agree bitwise."""
    vp = _VirtualPeers(world, 2 * n)
    g = torch.Generator(device=DEV).manual_seed(world * 1000 + n)
    torch.cuda.synchronize()
    for call in range(60):
        m = n if call % 3 else n // 2 + 1                  # message length may change between calls
        srcs = [torch.randn(m, device=DEV, generator=g) * (10.0 ** (r - 1)) for r in range(world)]
        outs = [torch.empty(m, device=DEV) for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            vp.launch(r, srcs[r], outs[r], 0.5)
        torch.cuda.synchronize()
        want = torch.zeros(m, device=DEV)
        for r in range(world):
            want = want + srcs[r] * 0.5                    # the kernel's order: ranks 0..world-1
        for r in range(world):
            assert torch.equal(outs[r], want), (call, r)


def test_peer_allreduce_kernel_graph_replay():
    """The epoch lives in device memory: a captured launch can be replayed (what GraphedTrainStep does at N > 1)."""
    world, n = 2, 6000
    vp = _VirtualPeers(world, n)
    srcs = [torch.zeros(n, device=DEV) for _ in range(world)]
    outs = [torch.empty(n, device=DEV) for _ in range(world)]
    for r in range(world):       # warm-up outside capture
        vp.launch(r, srcs[r], outs[r], 1.0)
    torch.cuda.synchronize()
    graphs = []
    for r in range(world):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=vp.streams[r]):
            lib = _lib.load()
            st = lib.wt_peer_allreduce(world, r, n, n, 1.0, _lib.ptr(srcs[r]), _lib.ptr(outs[r]), vp.base, vp.flags_off,
                                       _lib.ptr(vp.state[r]), torch.cuda.current_device(), _lib.stream_ptr(torch.device(DEV)))
            _lib.check(st, "wt_peer_allreduce")
        graphs.append(gr)
    for it in range(20):
        for r in range(world):
            srcs[r].fill_(float(it + 1) * (r + 1))
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(vp.streams[r]):
                graphs[r].replay()
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(outs[r], torch.full((n,), float(it + 1) * 3.0, device=DEV)), (it, r)


def test_slab_exchange_kernel_moves_exactly_the_ghost_rows():
    """wt_slab_exchange on three slabs of one process: after one exchange every ghost row holds the neighbour's owned row,
    owned rows are untouched, both fields, all samples; repeated exchanges keep working (epoch bookkeeping)."""
    from wavetorch_b200.domain import WtSlab
    lib = _lib.load()
    B, Ny, halo = 3, 24, 8
    owned = [20, 16, 12]
    ups = [0, halo, halo]
    dns = [halo, halo, 0]
    rows = [o + u + d for o, u, d in zip(owned, ups, dns)]
    f = [[torch.zeros(B, r, Ny, device=DEV) for _ in range(2)] for r in rows]
    flags = [torch.zeros(4, dtype=torch.int32, device=DEV) for _ in rows]
    state = [torch.zeros(2, dtype=torch.int32, device=DEV) for _ in rows]
    streams = [torch.cuda.Stream() for _ in rows]
    descs = []
    for i in range(3):
        d = WtSlab()
        d.halo, d.up, d.dn = halo, ups[i], dns[i]
        if ups[i]:
            d.up_Nx, d.up_f1, d.up_f2, d.up_flags = rows[i - 1], f[i - 1][0].data_ptr(), f[i - 1][1].data_ptr(), flags[i - 1].data_ptr()
        if dns[i]:
            d.dn_Nx, d.dn_f1, d.dn_f2, d.dn_flags = rows[i + 1], f[i + 1][0].data_ptr(), f[i + 1][1].data_ptr(), flags[i + 1].data_ptr()
        d.flags, d.state = flags[i].data_ptr(), state[i].data_ptr()
        descs.append(d)
    g0 = [0, owned[0], owned[0] + owned[1]]                   # global index of each slab's first owned row
    for rep in range(5):
        # owned rows <- a function of (field, sample, global row, column, rep); ghost rows <- garbage
        for i in range(3):
            for k in range(2):
                f[i][k].fill_(-777.0)
                gr = torch.arange(g0[i], g0[i] + owned[i], device=DEV, dtype=torch.float32)[None, :, None]
                bb = torch.arange(B, device=DEV, dtype=torch.float32)[:, None, None]
                cc = torch.arange(Ny, device=DEV, dtype=torch.float32)[None, None, :]
                f[i][k][:, ups[i]:ups[i] + owned[i]] = 1000.0 * k + 100.0 * bb + gr + 0.01 * cc + 0.5 * rep
        torch.cuda.synchronize()
        for i in range(3):
            with torch.cuda.stream(streams[i]):
                st = lib.wt_slab_exchange(ctypes.byref(descs[i]), B, rows[i], Ny, _lib.ptr(f[i][0]), _lib.ptr(f[i][1]),
                                          torch.cuda.current_device(), _lib.stream_ptr(torch.device(DEV)))
            _lib.check(st, "wt_slab_exchange")
        torch.cuda.synchronize()
        for i in range(3):
            for k in range(2):
                lo = g0[i] - ups[i]
                gr = torch.arange(lo, lo + rows[i], device=DEV, dtype=torch.float32)[None, :, None]
                bb = torch.arange(B, device=DEV, dtype=torch.float32)[:, None, None]
                cc = torch.arange(Ny, device=DEV, dtype=torch.float32)[None, None, :]
                want = 1000.0 * k + 100.0 * bb + gr + 0.01 * cc + 0.5 * rep
                assert torch.equal(f[i][k], want.expand(B, rows[i], Ny)), (rep, i, k)


def test_slab_descriptor_validation():
    from wavetorch_b200.domain import WtSlab
    lib = _lib.load()
    f = torch.zeros(1, 32, 16, device=DEV)
    d = WtSlab()
    d.halo, d.up, d.dn = 12, 0, 0      # not a multiple of 8
    d.flags = d.state = f.data_ptr()
    st = lib.wt_slab_exchange(ctypes.byref(d), 1, 32, 16, _lib.ptr(f), _lib.ptr(f), 0, None)
    assert st == -1 and b"multiple of 8" in lib.wt_last_error()
    d.halo, d.up = 8, 8                # neighbour not mapped
    st = lib.wt_slab_exchange(ctypes.byref(d), 1, 32, 16, _lib.ptr(f), _lib.ptr(f), 0, None)
    assert st == -1 and b"not mapped" in lib.wt_last_error()
