"""Slab-decomposed (multi-GPU) Zeldovich IC and 3-D FFT: one process per GPU, torch.distributed for the exchange.

The reference is a single process; this is the SURVEY 8(e) decomposition of its unigrid path
(ICGenerator::done, ic.hpp:1633-1648; FieldFourierManager::performTransform, fourier.hpp:479-518):

    Fourier space  rank r holds the ky-slab [kx:n][ky_local:P][kz:n/2+1]     (P = n / world)
    real space     rank r holds the x-slab  [x_local:P][y:n][z:n+2]          (contiguous particle ids)
    inverse FFT  = local x-pass | all-to-all | local y-pass | local z-pass   (forward: the reverse)

The exchange buffer is [peer][x_local][ky_local][kz]: the x-pass leaves it in that order and the y-pass reads the
received blocks in place (csrc/slab.cu), so the only data movement besides the passes is ONE all-to-all per
transform (NCCL over NVLink) plus one all-gather of the two Hermitian-mirror planes per field.

`SlabPipeline` is written against two small interfaces so that the sequencing and the buffer/layout logic are the
same objects on a B200 node and in the world_size-2 gloo tests that run on CPU:
    ops  : the per-rank passes. Product: `CudaSlabOps` (C ABI, raises if the CUDA library is missing).
           The tests supply a numpy stand-in; nothing in this package computes field data on the CPU.
    comm : `TorchComm` (torch.distributed all_to_all_single / all_gather_into_tensor; nccl or gloo).
"""
import ctypes as C
import json
import os
import time

import numpy as np

from . import lib as L


class SlabGeometry:
    """Sizes and index ranges of one rank's slabs."""

    def __init__(self, n, world, rank, itemsize=8):
        if n & (n - 1) or world & (world - 1) or n % world:
            raise L.GxError("slab: n and world must be powers of two with world | n")
        self.n, self.world, self.rank = n, world, rank
        self.P = n // world
        self.nc = n // 2 + 1
        self.itemsize = itemsize                                  # bytes per real
        self.slab_complex = n * self.P * self.nc                   # complex elements per slab (either space)
        self.slab_bytes = 2 * self.slab_complex * itemsize
        self.block_complex = self.P * self.P * self.nc             # one peer block of the exchange buffer
        self.planes_local_complex = n * self.P * 2                 # kz = 0 and kz = n/2 of the ky-slab
        self.iy0 = rank * self.P                                   # first ky storage index owned (Fourier space)
        self.x0 = rank * self.P                                    # first x plane owned (real space)
        self.id0 = self.x0 * n * n                                 # first particle id owned
        self.particles = self.P * n * n


class TorchComm:
    """The two collectives of the slab path, on torch tensors (uint8 views of the slab buffers)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def all_to_all(self, recv, send):
        """Equal-split exchange: block s of `send` goes to rank s, block r of `recv` comes from rank r."""
        if self.world == 1:
            recv.copy_(send)
            return
        self.dist.all_to_all_single(recv, send, group=self.group)

    def all_to_all_async(self, recv, send):
        """Starts the exchange and returns a handle with wait(): the collective runs on the backend's own stream, so kernels
        issued on the current stream after this call overlap with it (the next component's x-pass)."""
        if self.world == 1:
            recv.copy_(send)
            return None
        return self.dist.all_to_all_single(recv, send, group=self.group, async_op=True)

    def all_gather(self, out, inp):
        if self.world == 1:
            out.copy_(inp)
            return
        self.dist.all_gather_into_tensor(out, inp, group=self.group)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def stream_barrier(self):
        """Barrier in stream order: work enqueued after it on the current stream starts only when every rank's stream has
        reached the same point (a one-element all-reduce; the kernels that store into peer memory need nothing else)."""
        if self.world == 1:
            return
        import torch
        if getattr(self, "_token", None) is None:
            self._token = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.dist.all_reduce(self._token, group=self.group)

    def peer_pointers(self, buf):
        """Device pointers, valid in THIS process, to the same buffer on every rank (buf from CudaSlabOps.alloc_shared):
        CUDA IPC handles exchanged once; the own entry is the local pointer."""
        import torch
        lib = L.load()
        if self.world == 1:
            return [buf.data_ptr()]
        h = C.create_string_buffer(64)
        lib.gx_ipc_get_handle(C.c_void_p(buf.data_ptr()), h)
        mine = torch.frombuffer(bytearray(h.raw), dtype=torch.uint8).cuda()
        allh = torch.empty(64 * self.world, dtype=torch.uint8, device="cuda")
        self.dist.all_gather_into_tensor(allh, mine, group=self.group)
        raw = bytes(allh.cpu().numpy().tobytes())
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(buf.data_ptr())
                continue
            out = C.c_void_p()
            lib.gx_ipc_open(C.create_string_buffer(raw[64 * r:64 * (r + 1)], 64), C.byref(out))
            ptrs.append(out.value)
        return ptrs


class _RawCuda:
    """A gx_malloc allocation seen by torch (torch.as_tensor reads __cuda_array_interface__): CUDA IPC needs the base of a
    cudaMalloc allocation, which a tensor from torch's caching allocator is not."""

    def __init__(self, lib, nbytes):
        self.lib, self.ptr = lib, C.c_void_p()
        lib.gx_malloc(C.byref(self.ptr), int(nbytes))
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (self.ptr.value, False),
                                         "version": 2}

    def __del__(self):
        try:
            import torch
            torch.cuda.synchronize()    # the buffer was used on non-blocking streams (and by peers): gx_free pools it unsynchronised
            self.lib.gx_free(self.ptr)
        except Exception:
            pass


class CudaSlabOps:
    """Per-rank passes through the C ABI (include/genetic_b200.h, 'slab decomposition'). Buffers are torch CUDA
    tensors (so that NCCL can see them); kernels are launched on torch's current stream."""

    def __init__(self, geo, dtype=np.float64, device=None):
        import torch
        self.torch = torch
        self.lib = L.load()  # raises if the CUDA library has not been built: there is no CPU fallback
        self.geo = geo
        self.dt = L.F64 if np.dtype(dtype) == np.float64 else L.F32
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())

    def alloc(self, nbytes):
        return self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.device)

    def alloc_shared(self, nbytes):
        """A buffer the other ranks' kernels may store into (peer-mapped through CUDA IPC)."""
        raw = _RawCuda(self.lib, nbytes)
        t = self.torch.as_tensor(raw, device=self.device)
        t._gx_raw = raw  # keeps the allocation alive as long as the tensor
        return t

    def _st(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _ptr_table(ptrs):
        return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])

    def noise_staged(self, slab, peer_slabs, seed, reverse, spectrum, grid_length, scratch):
        """Parallel Fourier order, every shell generated on one rank, modes stored straight into their owners' slabs."""
        g = self.geo
        self.lib.gx_slab_noise_staged(self._p(slab), g.n, self.dt, g.world, g.rank, seed, int(reverse),
                                      self._ptr_table(peer_slabs), spectrum.handle if spectrum is not None else None,
                                      float(grid_length), self._p(scratch) if scratch is not None else None,
                                      int(scratch.numel()) if scratch is not None else 0, self._st())

    def x_inverse_mirror_push(self, slab, planes_all, peer_exchange, grid_length, op, axis):
        g = self.geo
        self.lib.gx_slab_fft_x_mirror_push(self._p(slab), self._p(planes_all), self._ptr_table(peer_exchange), g.n, self.dt,
                                           g.world, g.rank, float(grid_length), op, axis, self._st())

    def y_pass_op(self, src, dst, src_exchange, grid_length, axis):
        g = self.geo
        self.lib.gx_slab_fft_y_op(self._p(src), self._p(dst), g.n, self.dt, g.world, int(src_exchange), float(grid_length),
                                  axis, self._st())

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def noise_draw(self, slab, seed, mode, reverse, offset=0):
        g = self.geo
        off = C.c_ulonglong(offset)
        self.lib.gx_slab_noise_draw(self._p(slab), g.n, self.dt, g.world, g.rank, seed, mode, int(reverse),
                                    C.byref(off), self._st())
        return off.value

    # parallel Fourier order in phases: pass A shared between the ranks (include/genetic_b200.h)
    def noise_fn_bytes(self):
        return self.lib.gx_slab_noise_fn_bytes(self.geo.n, self.geo.world)

    def noise_pass_a(self, fn_local, seed):
        g = self.geo
        self.lib.gx_slab_noise_pass_a(self._p(fn_local), g.n, self.dt, g.world, g.rank, seed, self._st())

    def noise_finish(self, slab, fn_all, seed, reverse):
        g = self.geo
        self.lib.gx_slab_noise_finish(self._p(slab), self._p(fn_all), g.n, self.dt, g.world, g.rank, seed, int(reverse),
                                      self._st())

    def unit_variance(self, slab):
        g = self.geo
        self.lib.gx_slab_unit_variance(self._p(slab), g.n, self.dt, g.world, g.rank, self._st())

    def apply_transfer(self, slab, grid_length, spectrum, power, divide_k2):
        g = self.geo
        self.lib.gx_slab_apply_transfer(self._p(slab), g.n, self.dt, g.world, g.rank, float(grid_length),
                                        spectrum.handle, float(power), int(divide_k2), self._st())

    def extract_planes(self, slab, planes_local):
        g = self.geo
        self.lib.gx_slab_extract_planes(self._p(slab), self._p(planes_local), g.n, self.dt, g.world, self._st())

    def x_inverse_mirror(self, slab, planes_all, out, grid_length, op, axis):
        g = self.geo
        self.lib.gx_slab_fft_x_mirror(self._p(slab), self._p(planes_all), self._p(out), g.n, self.dt, g.world, g.rank,
                                      float(grid_length), op, axis, self._st())

    def x_pass(self, slab, direction, scale):
        g = self.geo
        self.lib.gx_slab_fft_x(self._p(slab), g.n, self.dt, g.world, direction, int(scale), self._st())

    def y_pass(self, src, dst, direction, src_exchange, dst_exchange):
        g = self.geo
        self.lib.gx_slab_fft_y(self._p(src), self._p(dst), g.n, self.dt, g.world, direction, int(src_exchange),
                               int(dst_exchange), self._st())

    def z_r2c(self, slab):
        g = self.geo
        self.lib.gx_slab_fft_z_r2c(self._p(slab), g.n, self.dt, g.world, self._st())

    def z_c2r(self, slab):
        g = self.geo
        self.lib.gx_slab_fft_z_c2r(self._p(slab), g.n, self.dt, g.world, self._st())

    def z_emit(self, slab, axis, cell_size, offset_lower, domain_size, vel_ratio, plane0, nplanes, out_pos, out_vel,
               out_dtype):
        g = self.geo
        odt = L.F64 if np.dtype(out_dtype) == np.float64 else L.F32
        self.lib.gx_slab_z_emit(self._p(slab), g.n, self.dt, g.world, g.rank, axis, float(cell_size),
                                float(offset_lower), float(domain_size), float(vel_ratio), plane0, nplanes,
                                self._p(out_pos) if out_pos is not None else None,
                                self._p(out_vel) if out_vel is not None else None, odt, self._st())

    def zero_dc(self, slab):
        self.lib.gx_slab_zero_dc(self._p(slab), self.dt, self._st())


class SlabPipeline:
    """Sequencing of the slab path for one rank. All tensors are flat uint8 buffers of geo.slab_bytes."""

    def __init__(self, geo, ops, comm, overlap=False, lean=False):
        self.geo, self.ops, self.comm = geo, ops, comm
        g = geo
        self.bufA = ops.alloc(g.slab_bytes)
        # lean: peer-memory mode only (enable_push) -- the receive buffer of the all-to-all path is never allocated
        self.bufB = ops.alloc(g.slab_bytes) if not lean else None
        # overlap: a second pair of buffers so that the all-to-all of one displacement component runs while the next
        # component's x-pass and the previous component's y/z-passes compute (5 slab-sized buffers in total with delta)
        self.overlap = bool(overlap) and g.world > 1 and hasattr(comm, "all_to_all_async")
        self.bufA2 = ops.alloc(g.slab_bytes) if self.overlap else None
        self.bufB2 = ops.alloc(g.slab_bytes) if self.overlap else None
        self.planes_local = ops.alloc(2 * g.planes_local_complex * g.itemsize)
        self.planes_all = ops.alloc(2 * g.planes_local_complex * g.itemsize * g.world)
        self.fn_local = self.fn_all = None
        self.push = None

    def enable_push(self, delta, overlap=False):
        """Peer-memory mode (several GPUs of one node): the noise delivery and the inverse x-pass store straight into the
        owners' buffers over NVLink, so an IC needs no all-to-all at all. `delta` and the two exchange buffers must come
        from ops.alloc_shared. Returns False (and changes nothing) when ops / comm cannot do it (CPU stand-ins)."""
        ops, comm, g = self.ops, self.comm, self.geo
        if g.world == 1 or g.world > 8 or not (hasattr(ops, "x_inverse_mirror_push") and hasattr(comm, "peer_pointers")):
            return False
        ex0, ex1 = ops.alloc_shared(g.slab_bytes), ops.alloc_shared(g.slab_bytes)
        self.push = {"delta": delta, "ex": (ex0, ex1), "delta_peers": comm.peer_pointers(delta),
                     "ex_peers": (comm.peer_pointers(ex0), comm.peer_pointers(ex1)),
                     "overlap": bool(overlap) and hasattr(ops, "torch")}
        return True

    def _aux_stream(self):
        if getattr(self, "_aux", None) is None:
            self._aux = self.ops.torch.cuda.Stream()
        return self._aux

    def ic_push(self, seed, reverse, grid_length, spectrum, emit):
        """One IC in peer-memory mode: parallel Fourier noise with the transfer folded in, then the shared-transform
        Zeldovich schedule -- 2 x-passes (each one IS its FFT transpose), 3 y-passes, 3 z-passes, 3 barriers."""
        ops, comm, p = self.ops, self.comm, self.push
        delta, (ex0, ex1) = p["delta"], p["ex"]
        comm.stream_barrier()                      # nobody reads delta / the exchange buffers of the previous IC any more
        ops.noise_staged(delta, p["delta_peers"], seed, reverse, spectrum, grid_length, ex0)   # ex0 idle: staged draws
        comm.stream_barrier()                      # every rank's modes have landed
        self.gather_mirror_planes(delta)
        ops.x_inverse_mirror_push(delta, self.planes_all, p["ex_peers"][0], grid_length, 1, 0)   # i k_x G -> owners
        aux = self._aux_stream() if p.get("overlap") else None
        if aux is None:
            ops.x_inverse_mirror_push(delta, self.planes_all, p["ex_peers"][1], grid_length, 0, 0)   # G (mirrored) -> owners
            comm.stream_barrier()
            ops.y_pass(ex0, self.bufA, +1, True, False)
            emit(0, self.bufA)
        else:
            # the second push is NVLink-bound: it runs on a side stream under psi_x's y- and z-passes (HBM-bound)
            torch = ops.torch
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(main)
            aux.wait_event(ev)
            with torch.cuda.stream(aux):
                ops.x_inverse_mirror_push(delta, self.planes_all, p["ex_peers"][1], grid_length, 0, 0)
                ev2 = torch.cuda.Event()
                ev2.record(aux)
            comm.stream_barrier()                  # exchange buffer 0 is complete on every rank
            ops.y_pass(ex0, self.bufA, +1, True, False)
            emit(0, self.bufA)
            main.wait_event(ev2)
            comm.stream_barrier()                  # exchange buffer 1 is complete on every rank
        ops.y_pass_op(ex1, self.bufA, True, grid_length, 1)                                    # i k_y on load
        emit(1, self.bufA)
        ops.y_pass_op(ex1, self.bufA, True, grid_length, 2)                                    # i k_z on load
        emit(2, self.bufA)

    # ---- general transforms (Field::toReal / toFourier on a slab) ---------------------------------------------
    def gather_mirror_planes(self, slab):
        self.ops.extract_planes(slab, self.planes_local)
        self.comm.all_gather(self.planes_all, self.planes_local)

    def inverse_after_x(self, work):
        """bufA holds the x-transformed ky-slab: exchange, y-pass into `work` ([x_local][y][kz])."""
        self.comm.all_to_all(self.bufB, self.bufA)
        self.ops.y_pass(self.bufB, work, +1, True, False)

    def to_real(self, slab):
        """In: ky-slab (Fourier). Out: x-slab (real, padded rows) in the same buffer."""
        self.gather_mirror_planes(slab)
        self.ops.x_inverse_mirror(slab, self.planes_all, self.bufA, 1.0, 0, 0)
        self.inverse_after_x(slab)
        self.ops.z_c2r(slab)

    def to_fourier(self, slab):
        """In: x-slab (real). Out: ky-slab (Fourier) in the same buffer."""
        self.ops.z_r2c(slab)
        self.ops.y_pass(slab, self.bufA, -1, False, True)
        self.comm.all_to_all(slab, self.bufA)
        self.ops.x_pass(slab, -1, True)

    # ---- the IC path -------------------------------------------------------------------------------------------
    def draw_noise(self, delta, seed, mode, reverse):
        """RandomFieldGenerator::draw for this rank's slab. With per-shell streams and more than one rank, the chunk parse
        is split between the ranks and its 40-byte-per-chunk results are all-gathered (one small collective)."""
        ops, g = self.ops, self.geo
        if g.world > 1 and mode == L.NOISE_FOURIER_PARALLEL and hasattr(ops, "noise_pass_a"):
            nb = ops.noise_fn_bytes()
            if self.fn_local is None or self.fn_local.numel() != nb:
                self.fn_local, self.fn_all = ops.alloc(nb), ops.alloc(nb * g.world)
            ops.noise_pass_a(self.fn_local, seed)
            self.comm.all_gather(self.fn_all, self.fn_local)
            ops.noise_finish(delta, self.fn_all, seed, reverse)
        else:
            ops.noise_draw(delta, seed, mode, reverse)

    def draw_delta(self, delta, seed, mode, reverse, fix_power, grid_length, spectrum):
        self.draw_noise(delta, seed, mode, reverse)
        if mode == L.NOISE_REAL:
            self.to_fourier(delta)
            if self.geo.rank == 0:
                self.ops.zero_dc(delta)
        if fix_power:
            self.ops.unit_variance(delta)
        self.ops.apply_transfer(delta, grid_length, spectrum, 0.5, 1)  # leaves delta_k / k^2

    def zeldovich(self, delta, grid_length, emit):
        """delta: ky-slab holding delta_k / k^2. emit(axis, work) consumes the half-transformed x-slab
        ([x_local][y][kz], y and x already inverse-transformed) of each displacement component."""
        self.gather_mirror_planes(delta)
        if not self.overlap:
            for axis in range(3):
                self.ops.x_inverse_mirror(delta, self.planes_all, self.bufA, grid_length, 1, axis)
                self.inverse_after_x(self.bufA)  # B <- exchange(A); A <- y-pass(B)
                emit(axis, self.bufA)
            return
        A, B = (self.bufA, self.bufA2), (self.bufB, self.bufB2)
        self.ops.x_inverse_mirror(delta, self.planes_all, A[0], grid_length, 1, 0)
        work = self.comm.all_to_all_async(B[0], A[0])
        for axis in range(3):
            cur, nxt = axis & 1, (axis + 1) & 1
            if axis + 1 < 3:   # runs while the exchange of `axis` is in flight; A[nxt] was last read by emit(axis - 1)
                self.ops.x_inverse_mirror(delta, self.planes_all, A[nxt], grid_length, 1, axis + 1)
            if work is not None:
                work.wait()
            if axis + 1 < 3:   # B[nxt] was last read by the y-pass of axis - 1, already enqueued on this stream
                work = self.comm.all_to_all_async(B[nxt], A[nxt])
            self.ops.y_pass(B[cur], A[cur], +1, True, False)
            emit(axis, A[cur])


class EmitStager:
    """Bounded staging for the particles of one rank: the z-pass emits `planes` x-planes of one component at a time into
    one of two device halves; with `to_host` each half is copied to pinned host memory on a copy stream while the next
    emit runs (the host side is a two-slot ring: a writer thread is its consumer in an application). Keeps a 2048^3 run
    inside 180 GB at any world size -- whole-slab pos+vel arrays would need 24 bytes x n^3 / world per rank on top of
    the three slab buffers."""

    def __init__(self, ops, geo, out_dtype, stage_bytes=1 << 30):
        import torch
        self.torch, self.ops, self.geo = torch, ops, geo
        self.osz = np.dtype(out_dtype).itemsize
        n2 = geo.n * geo.n
        self.planes = int(max(1, min(geo.P, stage_bytes // (4 * n2 * self.osz))))
        half = self.planes * n2 * self.osz
        self.dev = [(ops.alloc(half), ops.alloc(half)) for _ in range(2)]
        self.host = None
        self.copy_stream = torch.cuda.Stream()
        self.ev_copy = [None, None]
        self.turn = 0
        self.d2h_bytes = 0

    def enable_host(self):
        if self.host is None:
            half = self.dev[0][0].numel()
            with L.numa_local(self.torch.cuda.current_device()):           # pinned pages on this GPU's NUMA node
                self.host = [(self.torch.empty(half, dtype=self.torch.uint8, pin_memory=True),
                              self.torch.empty(half, dtype=self.torch.uint8, pin_memory=True)) for _ in range(2)]

    def emit(self, work, axis, grid, vel_ratio, out_dtype, to_host):
        torch, g = self.torch, self.geo
        cur = torch.cuda.current_stream()
        for plane0 in range(0, g.P, self.planes):
            npl = min(self.planes, g.P - plane0)
            b = self.turn
            self.turn ^= 1
            if to_host and self.ev_copy[b] is not None:
                cur.wait_event(self.ev_copy[b])          # the half is free once its previous copy has finished
            dpos, dvel = self.dev[b]
            self.ops.z_emit(work, axis, float(grid.cellSize), float(grid.offsetLower[axis]),
                            float(grid.periodicDomainSize), vel_ratio, plane0, npl, dpos, dvel, out_dtype)
            if to_host:
                ev = torch.cuda.Event()
                ev.record(cur)
                self.copy_stream.wait_event(ev)
                nb = npl * g.n * g.n * self.osz
                with torch.cuda.stream(self.copy_stream):
                    self.host[b][0][:nb].copy_(dpos[:nb], non_blocking=True)
                    self.host[b][1][:nb].copy_(dvel[:nb], non_blocking=True)
                    self.ev_copy[b] = torch.cuda.Event()
                    self.ev_copy[b].record(self.copy_stream)
                self.d2h_bytes += 2 * nb

    def finish(self):
        self.copy_stream.synchronize()


def unigrid_ic_slab(geo, ops, comm, seed, mode, reverse, fix_power, grid, spectrum, vel_ratio, out_dtype, delta, pipe,
                    out_pos, out_vel, planes_per_emit=None, stager=None, to_host=False):
    """One IC for this rank's x-slab. Either out_pos/out_vel are [3][particles] (component-major) device tensors that
    receive every particle, or `stager` streams them through bounded staging (and to pinned host memory if to_host)."""
    g = geo
    osz = np.dtype(out_dtype).itemsize

    def emit(axis, work):
        if stager is not None:
            stager.emit(work, axis, grid, vel_ratio, out_dtype, to_host)
            return
        ops.z_emit(work, axis, float(grid.cellSize), float(grid.offsetLower[axis]), float(grid.periodicDomainSize),
                   vel_ratio, 0, g.P, out_pos[axis * g.particles * osz:], out_vel[axis * g.particles * osz:], out_dtype)

    if pipe.push is not None and mode == L.NOISE_FOURIER_PARALLEL and not fix_power and pipe.push["delta"] is delta:
        pipe.ic_push(seed, reverse, float(grid.thisGridSize), spectrum, emit)
    else:
        pipe.draw_delta(delta, seed, mode, reverse, fix_power, float(grid.thisGridSize), spectrum)
        pipe.zeldovich(delta, float(grid.thisGridSize), emit)
    if stager is not None and to_host:
        stager.finish()


# ---- bench entry (python bench.py under torchrun, N > 1) ------------------------------------------------------------
def bench_main(args, lib, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    import genetic_b200 as gb
    import bench as B

    n = args.n
    dtype = np.float32 if getattr(args, "dtype", "f64") == "f32" else np.float64   # f32: the reference's -f build (config 5)
    out_dtype = np.float32
    dname = "f32" if dtype == np.float32 else "f64"
    box = B.box_for(n)
    cos = gb.CosmologicalParameters(B.COSMO["Om"], B.COSMO["Ol"], B.COSMO["s8"], B.COSMO["z"], ns=B.COSMO["ns"],
                                    hubble=B.COSMO["hubble"])
    grid = gb.Grid(box, n, dtype=dtype)
    spectrum = gb.CAMB(cos, B.CAMB_FILE, "all", dtype)
    geo = SlabGeometry(n, world, rank, np.dtype(dtype).itemsize)
    ops = CudaSlabOps(geo, dtype)
    comm = TorchComm()
    # the overlapped exchange needs 5 slab-sized buffers (+ staging): use it when they fit
    free_b, _ = torch.cuda.mem_get_info()
    push = os.environ.get("GX_SLAB_PUSH", "1") != "0" and world <= 8
    if push:
        # peer-memory mode: delta, two exchange buffers (stored into by the peers) and one work buffer
        pipe = SlabPipeline(geo, ops, comm, overlap=False, lean=True)
        delta = ops.alloc_shared(geo.slab_bytes)
        push = pipe.enable_push(delta, overlap=os.environ.get("GX_SLAB_PUSH_OVERLAP", "0") == "1")
    if not push:
        overlap = os.environ.get("GX_SLAB_OVERLAP", "1") != "0" and free_b > 5 * geo.slab_bytes + (12 << 30)
        pipe = SlabPipeline(geo, ops, comm, overlap=overlap)
        delta = ops.alloc(geo.slab_bytes)
    osz = 4
    vel_ratio = cos.zeldovichVelocityToOffsetRatio(dtype)
    stager = EmitStager(ops, geo, out_dtype)   # particles leave through bounded staging, as in gx_unigrid_ic
    stager.enable_host()

    def device_step():
        unigrid_ic_slab(geo, ops, comm, B.SEED, L.NOISE_FOURIER_PARALLEL, 0, 0, grid, spectrum, vel_ratio, out_dtype,
                        delta, pipe, None, None, stager=stager, to_host=False)

    def e2e_step():
        sp = gb.CAMB(cos, B.CAMB_FILE, "all", dtype)
        unigrid_ic_slab(geo, ops, comm, B.SEED, L.NOISE_FOURIER_PARALLEL, 0, 0, grid, sp, vel_ratio, out_dtype, delta,
                        pipe, None, None, stager=stager, to_host=True)
        torch.cuda.synchronize()
        return sp

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = B.ClockSampler(local_rank)
    sampler.start()
    lib.gx_timing_enable(1)
    l0 = lib.gx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    ms_local = e0.elapsed_time(e1) / args.steps
    launches = lib.gx_launch_count() - l0
    buf = C.create_string_buffer(1 << 16)
    lib.gx_timing_report(buf, len(buf))
    lib.gx_timing_enable(0)
    clocks = sampler.stop()
    t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # (config 5 moves 1.65 TB to the host per IC: GX_BENCH_E2E_STEPS / GX_BENCH_E2E_WARMUP bound that leg)
    e2e_steps = int(os.environ.get("GX_BENCH_E2E_STEPS", args.steps))
    for _ in range(int(os.environ.get("GX_BENCH_E2E_WARMUP", min(args.warmup, 2)))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_local = (time.perf_counter() - t0) / max(e2e_steps, 1)
    t = torch.tensor([e2e_local], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    classes = {}
    for line in buf.value.decode().strip().split("\n"):
        if line:
            nm, cnt, tot = line.split()
            classes[nm] = (int(cnt), float(tot))
    n3 = n ** 3
    checksum = B.particle_checksum(lib, torch, device_step, dist)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(B.ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        def n_launch(k):
            return classes[k][0] / args.steps if k in classes else 0
        # class bytes per step on this rank; a class may be several launches per step
        n_gauss_rank = 8 * sum(3 * ks * ks for ks in range(1, n // 2)) // world
        class_bytes = {"fft_cols_x": n_launch("fft_cols_x") * 2 * geo.slab_bytes,
                       "fft_cols_y": n_launch("fft_cols_y") * 2 * geo.slab_bytes,
                       "fft_rows_c2r": 3 * (geo.slab_bytes + 2 * osz * geo.particles),
                       "rng_gen": 4 * n_gauss_rank, "rng_deliver": 4 * n_gauss_rank + geo.slab_bytes}
        cand = {k: v for k, v in classes.items() if k in class_bytes and class_bytes[k] > 0}
        dom = max(cand, key=lambda k: cand[k][1]) if cand else None
        roof = None
        if dom:
            cnt, tot = classes[dom]
            avg_ms = tot / cnt
            per_launch = class_bytes[dom] * args.steps / cnt
            achieved = per_launch / (avg_ms * 1e-3) / 1e9
            passes = {k: {"GB/s": class_bytes[k] / (classes[k][1] / args.steps * 1e-3) / 1e9,
                          "frac": class_bytes[k] / (classes[k][1] / args.steps * 1e-3) / 1e9 / peak}
                      for k in class_bytes if k in classes and classes[k][1] > 0}
            roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None,
                    "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "algorithmic_bytes_per_launch": per_launch, "avg_launch_ms": avg_ms,
                    "per_class_ms_per_step_rank0": {k: v[1] / args.steps for k, v in sorted(classes.items())},
                    "passes": passes, "share_of_step": tot / args.steps / ms}
        if pipe.push is not None:
            # peer-memory mode: the two inverse x-passes store (world-1)/world of their output into the peers' exchange buffers
            # (that IS the FFT transpose), the noise delivery the same share of the field; the x-pass is timed by CUDA events
            # on this rank, so bytes / time is the NVLink rate its stores sustained
            nx = classes["fft_cols_x"][0] / args.steps if "fft_cols_x" in classes else 0
            push_bytes = geo.slab_bytes * (world - 1) // world
            x_ms = classes["fft_cols_x"][1] / classes["fft_cols_x"][0] if nx else float("nan")
            gbs = push_bytes / (x_ms * 1e-3) / 1e9
            nvlink = {"mode": "peer stores fused into the inverse x-pass and the noise delivery (no all-to-all)",
                      "x_pass_push_bytes_per_gpu_per_launch": push_bytes, "x_pass_launches_per_step": nx,
                      "x_pass_ms_per_launch_rank0": x_ms, "GB/s_per_direction": gbs, "frac_of_900": gbs / 900.0,
                      "frac_of_measured_770_peer_copy": gbs / 770.0,
                      "noise_delivery_push_bytes_per_gpu_per_step": push_bytes,
                      "bytes_sent_per_gpu_per_step": int(push_bytes * (nx + 1))}
        else:
            a2a_bytes = 3 * geo.slab_bytes * (world - 1) // world
            nvlink = {"mode": "NCCL all_to_all_single per inverse transform",
                      "all_to_all_bytes_sent_per_gpu_per_step": a2a_bytes}
        line = {"metric": "IC particles/sec (Zeldovich %d^3)" % n, "value": n3 / (ms * 1e-3), "unit": "particles/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dname, "data": "synthetic",
                "config": {"workload": B.workload_name(n, dname), "output": "float32 pos+vel, component-major, x-slab per rank",
                           "parallelism": ("slab%d: ky-slabs in Fourier space, x-slabs in real space; peer-memory mode: noise "
                                           "delivery and inverse x-passes store straight into the owners' buffers over NVLink "
                                           "(CUDA IPC), 2 transposes per IC, 3 stream-ordered barriers" % world)
                           if pipe.push is not None else
                           "slab%d: ky-slabs in Fourier space, x-slabs in real space, one NCCL all-to-all "
                           "per inverse transform%s" % (world, ", overlapped with the neighbouring components' "
                                                               "passes" if pipe.overlap else ""),
                           "l2": "inputs larger than L2 (each slab %.2f GB)" % (geo.slab_bytes / 1e9)},
                "clocks": clocks,
                "e2e": {"value": n3 / e2e_s, "unit": "particles/s", "h2d_bytes_per_step": 3 * 8 * 213 * world,
                        "d2h_bytes_per_step": 6 * n3 * osz, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps},
                "gpu_launches": int(launches) * world,
                "nvlink": nvlink,
                "roofline": roof, "cpu_baseline": None, "checksum": checksum}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0
