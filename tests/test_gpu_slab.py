"""GPU parity of the slab-decomposed passes (csrc/slab.cu) against the single-GPU path and the oracle.

All ranks of a `world` are driven from one process on one GPU: each rank is a thread running the product
SlabPipeline + CudaSlabOps unchanged, and the two collectives are replaced by an in-process exchange with the same
semantics as all_to_all_single / all_gather_into_tensor. (The NCCL binding itself is exercised by bench.py --gpus N.)
"""
import ctypes as C
import os
import sys
import threading

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu
CAMB_FILE = os.path.join(ROOT, "tests", "golden", "camb_transfer_kmax40_z0.dat")


class LocalComm:
    """all_to_all / all_gather between threads of one process (one per emulated rank)."""

    class Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, shared, rank):
        import torch
        self.torch, self.sh, self.rank, self.world = torch, shared, rank, shared.world

    def _publish(self, t):
        self.torch.cuda.synchronize()
        self.sh.slots[self.rank] = t
        self.sh.barrier.wait()

    def all_to_all(self, recv, send):
        self._publish(send)
        blk = send.numel() // self.world
        for s in range(self.world):
            recv[s * blk:(s + 1) * blk].copy_(self.sh.slots[s][self.rank * blk:(self.rank + 1) * blk])
        self.torch.cuda.synchronize()
        self.sh.barrier.wait()

    def all_to_all_async(self, recv, send):
        self.all_to_all(recv, send)   # in-process stand-in: completes before returning
        return None

    def all_gather(self, out, inp):
        self._publish(inp)
        blk = inp.numel()
        for s in range(self.world):
            out[s * blk:(s + 1) * blk].copy_(self.sh.slots[s])
        self.torch.cuda.synchronize()
        self.sh.barrier.wait()

    def barrier(self):
        self.sh.barrier.wait()

    def stream_barrier(self):
        self.torch.cuda.synchronize()
        self.sh.barrier.wait()

    def peer_pointers(self, buf):
        """In one process every rank's buffer is directly addressable: the product path gets these through CUDA IPC."""
        self._publish(buf)
        ptrs = [t.data_ptr() for t in self.sh.slots]
        self.sh.barrier.wait()
        return ptrs


class LockedOps:
    """The library keeps per-process scratch state (noise workspace): emulated ranks take turns."""
    lock = threading.Lock()

    def __init__(self, ops):
        self._ops = ops
        self.geo = ops.geo

    def __getattr__(self, name):
        fn = getattr(self._ops, name)
        if not callable(fn):
            return fn

        def call(*a, **k):
            with LockedOps.lock:
                return fn(*a, **k)
        return call


def _run_ranks(world, body):
    import torch
    shared = LocalComm.Shared(world)
    results, errors = [None] * world, []

    def target(r):
        try:
            torch.cuda.set_device(0)
            results[r] = body(r, LocalComm(shared, r))
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            shared.barrier.abort()

    th = [threading.Thread(target=target, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errors:
        raise errors[0]
    return results


@pytest.mark.parametrize("world,n,overlap", [(1, 32, False), (2, 32, False), (4, 64, False), (8, 64, False), (4, 512, False),
                                             (4, 64, True), (8, 128, True)])
def test_slab_ic_matches_single_gpu(world, n, overlap):
    import torch
    import genetic_b200 as gb
    from genetic_b200 import lib as L, slab as SL
    box, seed = 50.0 * n / 32, 8896131
    cos = gb.CosmologicalParameters(0.279, 0.721, 0.817, 99)
    pos1, vel1, _, _ = gb.generate_unigrid_ic(n, box, seed, cos, CAMB_FILE, out_dtype=np.float64)
    spectrum = gb.CAMB(cos, CAMB_FILE, "all", np.float64)
    grid = gb.Grid(box, n)
    ratio = cos.zeldovichVelocityToOffsetRatio(np.float64)

    def body(rank, comm):
        geo = SL.SlabGeometry(n, world, rank, 8)
        ops = LockedOps(SL.CudaSlabOps(geo, np.float64))
        pipe = SL.SlabPipeline(geo, ops, comm, overlap=overlap)
        delta = ops.alloc(geo.slab_bytes)
        out_pos, out_vel = ops.alloc(3 * geo.particles * 8), ops.alloc(3 * geo.particles * 8)
        SL.unigrid_ic_slab(geo, ops, comm, seed, L.NOISE_FOURIER_PARALLEL, 0, 0, grid, spectrum, ratio, np.float64, delta,
                           pipe, out_pos, out_vel)
        torch.cuda.synchronize()
        return (out_pos.cpu().numpy().view(np.float64).reshape(3, -1), out_vel.cpu().numpy().view(np.float64).reshape(3, -1))

    res = _run_ranks(world, body)
    pos = np.concatenate([r[0] for r in res], axis=1).T
    vel = np.concatenate([r[1] for r in res], axis=1).T
    # same butterflies, different pass order (x first here, y first on one GPU): 1e-10 relative is the north-star bar
    assert np.abs(vel - vel1).max() <= 1e-12 * np.abs(vel1).max()
    d = np.abs(pos - pos1)
    assert np.minimum(d, box - d).max() <= 1e-12 * box


@pytest.mark.parametrize("world,n,mode,reverse", [(2, 16, 0, 0), (4, 32, 1, 0), (2, 32, 1, 1), (4, 32, 2, 0)])
def test_slab_noise_bit_exact_and_transform_roundtrip(world, n, mode, reverse):
    import torch
    import genetic_b200 as gb
    from genetic_b200 import lib as L, slab as SL
    lib = L.load()
    grid = gb.Grid(50.0, n)
    full = gb.Field(grid, fourier=True)
    off = C.c_ulonglong(0)
    lib.gx_noise_draw(full.ptr, n, L.F64, 1234, mode, reverse, C.byref(off), None)
    W = full.getDataVector()
    real_ref = None
    f2 = full.copy()
    f2.toReal()
    real_ref = f2.getDataVector()

    def body(rank, comm):
        geo = SL.SlabGeometry(n, world, rank, 8)
        ops = LockedOps(SL.CudaSlabOps(geo, np.float64))
        pipe = SL.SlabPipeline(geo, ops, comm)
        slab = ops.alloc(geo.slab_bytes)
        end = ops.noise_draw(slab, 1234, mode, reverse)
        if mode == L.NOISE_REAL:
            pipe.to_fourier(slab)
            if rank == 0:
                ops.zero_dc(slab)
        torch.cuda.synchronize()
        noise = slab.cpu().numpy().view(np.complex128).reshape(n, geo.P, geo.nc).copy()
        pipe.to_real(slab)
        torch.cuda.synchronize()
        real = slab.cpu().numpy().view(np.float64).reshape(geo.P, n, n + 2)[:, :, :n].copy()
        pipe.to_fourier(slab)
        torch.cuda.synchronize()
        back = slab.cpu().numpy().view(np.complex128).reshape(n, geo.P, geo.nc).copy()
        return noise, real, back, end

    res = _run_ranks(world, body)
    noise = np.concatenate([r[0] for r in res], axis=1)
    if mode == L.NOISE_REAL:
        assert np.abs(noise - W).max() <= 1e-13 * np.abs(W).max()  # the r2c ran as x|y|z passes in a different order
    else:
        assert np.array_equal(noise.view(np.float64), W.view(np.float64)), "slab noise differs from the single-GPU draw"
        assert all(r[3] == off.value for r in res) or mode == L.NOISE_FOURIER_PARALLEL
    real = np.concatenate([r[1] for r in res], axis=0)
    assert np.abs(real - real_ref).max() <= 1e-13 * np.abs(real_ref).max()
    back = np.concatenate([r[2] for r in res], axis=1)
    f2.toFourier()
    ref_back = f2.getDataVector()
    assert np.abs(back - ref_back).max() <= 1e-13 * np.abs(ref_back).max()


@pytest.mark.parametrize("world,n,reverse", [(2, 32, 0), (4, 64, 0), (8, 64, 1), (8, 256, 0)])
def test_slab_noise_staged_peer_delivery_bit_exact(world, n, reverse):
    """gx_slab_noise_staged: every shell generated on one rank, every mode stored through the owner table into the slab of
    the rank that holds its ky row -- identical to the single-GPU draw slot for slot."""
    import torch
    import genetic_b200 as gb
    from genetic_b200 import lib as L, slab as SL
    lib = L.load()
    full = gb.Field(gb.Grid(50.0, n), fourier=True)
    lib.gx_noise_draw(full.ptr, n, L.F64, 4321, L.NOISE_FOURIER_PARALLEL, reverse, None, None)
    W = full.getDataVector()

    def body(rank, comm):
        geo = SL.SlabGeometry(n, world, rank, 8)
        ops = LockedOps(SL.CudaSlabOps(geo, np.float64))
        slab = ops.alloc_shared(geo.slab_bytes)
        slab.fill_(0x7f)                        # stale contents: every slot must be written or zeroed by the draw
        peers = comm.peer_pointers(slab)
        comm.stream_barrier()
        ops.noise_staged(slab, peers, 4321, reverse, None, 50.0, None)
        comm.stream_barrier()
        return slab.cpu().numpy().view(np.complex128).reshape(n, geo.P, geo.nc).copy()

    res = _run_ranks(world, body)
    noise = np.concatenate(res, axis=1)
    assert np.array_equal(noise.view(np.float64), W.view(np.float64))


@pytest.mark.parametrize("world,n", [(2, 32), (4, 64), (8, 64), (8, 256), (2, 512)])
def test_slab_ic_peer_memory_mode_matches_single_gpu(world, n):
    """Peer-memory mode (SlabPipeline.ic_push): staged noise with the transfer folded in, inverse x-passes that store
    straight into the owners' exchange buffers, psi_y and psi_z from one exchanged x-transform."""
    import torch
    import genetic_b200 as gb
    from genetic_b200 import lib as L, slab as SL
    box, seed = 50.0 * n / 32, 8896131
    cos = gb.CosmologicalParameters(0.279, 0.721, 0.817, 99)
    pos1, vel1, _, _ = gb.generate_unigrid_ic(n, box, seed, cos, CAMB_FILE, out_dtype=np.float64)
    spectrum = gb.CAMB(cos, CAMB_FILE, "all", np.float64)
    grid = gb.Grid(box, n)
    ratio = cos.zeldovichVelocityToOffsetRatio(np.float64)

    def body(rank, comm):
        geo = SL.SlabGeometry(n, world, rank, 8)
        ops = LockedOps(SL.CudaSlabOps(geo, np.float64))
        pipe = SL.SlabPipeline(geo, ops, comm, lean=True)
        delta = ops.alloc_shared(geo.slab_bytes)
        assert pipe.enable_push(delta)
        out_pos, out_vel = ops.alloc(3 * geo.particles * 8), ops.alloc(3 * geo.particles * 8)
        for _ in range(2):   # twice: the second IC reuses every buffer the peers stored into
            SL.unigrid_ic_slab(geo, ops, comm, seed, L.NOISE_FOURIER_PARALLEL, 0, 0, grid, spectrum, ratio, np.float64, delta,
                               pipe, out_pos, out_vel)
        torch.cuda.synchronize()
        return (out_pos.cpu().numpy().view(np.float64).reshape(3, -1), out_vel.cpu().numpy().view(np.float64).reshape(3, -1))

    res = _run_ranks(world, body)
    pos = np.concatenate([r[0] for r in res], axis=1).T
    vel = np.concatenate([r[1] for r in res], axis=1).T
    assert np.abs(vel - vel1).max() <= 1e-12 * np.abs(vel1).max()
    d = np.abs(pos - pos1)
    assert np.minimum(d, box - d).max() <= 1e-12 * box
