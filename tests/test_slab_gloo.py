"""World-size-2 (and 4) run of the slab pipeline's host logic over gloo on CPU.

genetic_b200.slab.SlabPipeline sequences the per-rank passes and the two collectives; on a B200 node the passes are
CUDA kernels (CudaSlabOps). Here the same SlabPipeline / TorchComm objects run with a numpy stand-in for the passes
(built from the oracle), so the buffer layouts, the exchange-block addressing, the mirror-plane all-gather and the
ownership ranges are checked against the single-process oracle without a GPU. The stand-in lives in tests/ only.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
CAMB_FILE = os.path.join(GOLD, "camb_transfer_kmax40_z0.dat")


class NumpyOps:
    """numpy/oracle implementation of the CudaSlabOps interface (test double)."""

    def __init__(self, geo, box, cosmo):
        import torch
        from oracle import model as M
        self.torch, self.M, self.geo, self.box, self.cosmo = torch, M, geo, box, cosmo

    def alloc(self, nbytes):
        return self.torch.zeros(int(nbytes), dtype=self.torch.uint8)

    def _c(self, t, shape):
        n = int(np.prod(shape))
        return t.numpy().view(np.complex128)[:n].reshape(shape)

    def noise_draw(self, slab, seed, mode, reverse, offset=0):
        g, M = self.geo, self.M
        name = {0: "real", 1: "fourier_reverse" if reverse else "fourier_serial", 2: "fourier_parallel"}[mode]
        if mode == 0:
            from oracle import crng
            real = crng.draw_real(crng.Stream(seed), g.n, False).reshape(g.n, g.n, g.n)
            rows = slab.numpy().view(np.float64).reshape(g.P, g.n, g.n + 2)
            rows[:] = 0
            rows[:, :, :g.n] = real[g.x0:g.x0 + g.P]
            return 0
        W = M.draw_white_noise(g.n, seed, name)
        self._c(slab, (g.n, g.P, g.nc))[:] = W[:, g.iy0:g.iy0 + g.P, :]
        return 0

    def unit_variance(self, slab):
        S = self._c(slab, (self.geo.n, self.geo.P, self.geo.nc))
        a = np.abs(S)
        S[a > 0] /= a[a > 0]

    def apply_transfer(self, slab, grid_length, spectrum, power, divide_k2):
        g, M = self.geo, self.M
        S = self._c(slab, (g.n, g.P, g.nc))
        full = np.zeros((g.n, g.n, g.nc), dtype=np.complex128)
        full[:, g.iy0:g.iy0 + g.P] = S
        cov = spectrum.covariance(grid_length, g.n, "all")
        full = M.apply_transfer(full, cov, power)
        if divide_k2:
            kmin = 2.0 * np.pi / grid_length
            ks = M.signed_modes(g.n).astype(np.float64) * kmin
            kz = np.arange(g.nc) * kmin
            k2 = ks[:, None, None] ** 2 + ks[None, :, None] ** 2 + kz[None, None, :] ** 2
            ind = M.independent_mask(g.n)
            with np.errstate(divide="ignore", invalid="ignore"):
                q = full / k2
            q[k2 == 0] = 0
            full = np.where(ind, q, full)
        S[:] = full[:, g.iy0:g.iy0 + g.P]

    def extract_planes(self, slab, planes_local):
        g = self.geo
        S = self._c(slab, (g.n, g.P, g.nc))
        pl = self._c(planes_local, (g.n, g.P, 2))
        pl[:, :, 0] = S[:, :, 0]
        pl[:, :, 1] = S[:, :, g.n // 2]

    def x_inverse_mirror(self, slab, planes_all, out, grid_length, op, axis):
        g, M = self.geo, self.M
        n, P, h = g.n, g.P, g.n // 2
        S = self._c(slab, (n, P, g.nc)).copy()
        pa = self._c(planes_all, (g.world, n, P, 2))
        kmin = 2.0 * np.pi / grid_length
        ks = M.signed_modes(n).astype(np.float64) * kmin

        def operator(d, ix, iy, iz):
            if op == 0:
                return d
            k = (ks[ix], ks[iy], iz * kmin)
            k2 = k[0] * k[0] + k[1] * k[1] + k[2] * k[2]
            ia = (ix, iy, iz)[axis]
            if ia == h or k2 == 0:
                return 0j
            if op == 2:
                d = d / k2
            return complex(-d.imag * k[axis], d.real * k[axis])

        R = np.empty_like(S)
        for ix in range(n):
            for iyl in range(P):
                iy = g.iy0 + iyl
                for iz in range(g.nc):
                    if iz in (0, h):
                        dependent = ix > h or (ix in (0, h) and iy > h)
                        self_conj = ix in (0, h) and iy in (0, h)
                        if dependent:
                            mx, my = (n - ix) % n, (n - iy) % n
                            v = np.conj(operator(pa[my // P, mx, my % P, 0 if iz == 0 else 1], mx, my, iz))
                        else:
                            v = operator(S[ix, iyl, iz], ix, iy, iz)
                            if self_conj:
                                v = np.conj(v)
                        R[ix, iyl, iz] = v
                    else:
                        R[ix, iyl, iz] = operator(S[ix, iyl, iz], ix, iy, iz)
        self._c(out, (n, P, g.nc))[:] = np.fft.ifft(R, axis=0) * n

    def x_pass(self, slab, direction, scale):
        g = self.geo
        S = self._c(slab, (g.n, g.P, g.nc))
        R = np.fft.fft(S, axis=0) if direction < 0 else np.fft.ifft(S, axis=0) * g.n
        S[:] = R * (float(g.n) ** -1.5 if scale else 1.0)

    def _plain(self, t, exchange):
        g = self.geo
        if exchange:
            E = self._c(t, (g.world, g.P, g.P, g.nc))
            return E.transpose(1, 0, 2, 3).reshape(g.P, g.n, g.nc)
        return self._c(t, (g.P, g.n, g.nc))

    def y_pass(self, src, dst, direction, src_exchange, dst_exchange):
        g = self.geo
        X = self._plain(src, src_exchange)
        R = np.fft.fft(X, axis=1) if direction < 0 else np.fft.ifft(X, axis=1) * g.n
        if dst_exchange:
            self._c(dst, (g.world, g.P, g.P, g.nc))[:] = R.reshape(g.P, g.world, g.P, g.nc).transpose(1, 0, 2, 3)
        else:
            self._c(dst, (g.P, g.n, g.nc))[:] = R

    def z_r2c(self, slab):
        g = self.geo
        rows = slab.numpy().view(np.float64).reshape(g.P * g.n, g.n + 2)
        R = np.fft.rfft(rows[:, :g.n].copy(), axis=1)
        self._c(slab, (g.P * g.n, g.nc))[:] = R

    def _c2r(self, slab):
        g = self.geo
        return np.fft.irfft(self._c(slab, (g.P, g.n, g.nc)), n=g.n, axis=2) * g.n * float(g.n) ** -1.5

    def z_c2r(self, slab):
        g = self.geo
        r = self._c2r(slab)
        rows = slab.numpy().view(np.float64).reshape(g.P, g.n, g.n + 2)
        rows[:, :, :g.n] = r

    def z_emit(self, slab, axis, cell_size, offset_lower, domain_size, vel_ratio, plane0, nplanes, out_pos, out_vel,
               out_dtype):
        g = self.geo
        psi = self._c2r(slab)[plane0:plane0 + nplanes]
        idx = [np.arange(g.x0 + plane0, g.x0 + plane0 + nplanes), np.arange(g.n), np.arange(g.n)][axis]
        cen = (idx.astype(np.float64) * cell_size + offset_lower) + cell_size / 2
        shape = [1, 1, 1]
        shape[axis] = -1
        q = np.fmod(psi + cen.reshape(shape), domain_size)
        q = np.where(q < 0, q + domain_size, q)
        cnt = nplanes * g.n * g.n
        out_pos.numpy().view(out_dtype)[:cnt] = q.reshape(-1).astype(out_dtype)
        out_vel.numpy().view(out_dtype)[:cnt] = (psi * vel_ratio).reshape(-1).astype(out_dtype)

    def zero_dc(self, slab):
        slab.numpy().view(np.float64)[0] = 0.0


class PushNumpyOps(NumpyOps):
    """Stand-in for the peer-memory entry points of CudaSlabOps (gx_slab_noise_staged, gx_slab_fft_x_mirror_push,
    gx_slab_fft_y_op): processes on a CPU cannot store into each other's buffers, so the "push" is the local x-pass followed
    by the same block exchange through gloo; the sequencing under test (SlabPipeline.ic_push: transfer folded into the noise,
    ONE exchanged x-transform of G shared by psi_y and psi_z) is the product's."""

    def __init__(self, geo, box, cosmo, comm):
        super().__init__(geo, box, cosmo)
        self.comm = comm

    def alloc_shared(self, nbytes):
        return self.alloc(nbytes)

    def noise_staged(self, slab, peer_slabs, seed, reverse, spectrum, grid_length, scratch):
        self.noise_draw(slab, seed, 2, reverse)
        if spectrum is not None:
            self.apply_transfer(slab, grid_length, spectrum, 0.5, 1)

    def x_inverse_mirror_push(self, slab, planes_all, peer_exchange, grid_length, op, axis):
        send = self.alloc(self.geo.slab_bytes)
        self.x_inverse_mirror(slab, planes_all, send, grid_length, op, axis)
        self.comm.all_to_all(self.comm.registered[peer_exchange[self.geo.rank]], send)

    def y_pass_op(self, src, dst, src_exchange, grid_length, axis):
        g, M = self.geo, self.M
        X = self._plain(src, src_exchange).copy()
        kmin = 2.0 * np.pi / grid_length
        if axis == 1:
            k = M.signed_modes(g.n).astype(np.float64) * kmin
            k[g.n // 2] = 0.0                      # zero on the Nyquist plane of the component's axis
            X = 1j * k[None, :, None] * X
        else:
            k = np.arange(g.nc) * kmin
            k[g.n // 2] = 0.0
            X = 1j * k[None, None, :] * X
        self._c(dst, (g.P, g.n, g.nc))[:] = np.fft.ifft(X, axis=1) * g.n


def _push_worker(rank, world, port, n, outdir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from genetic_b200 import slab as SL
    from genetic_b200 import host as H
    from oracle import model as M
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        class Comm(SL.TorchComm):
            registered = {}

            def stream_barrier(self):
                self.barrier()

            def peer_pointers(self, buf):          # a handle per buffer; the stand-in ops resolve their own entry
                key = len(self.registered) + 1
                self.registered[key] = buf
                return [key] * self.world

        box, seed = 50.0, 8896131
        ocos = M.Cosmology(0.279, 0.721, 0.817, 99)
        camb = M.CAMB(CAMB_FILE, ocos)
        geo = SL.SlabGeometry(n, world, rank, 8)
        comm = Comm()
        ops = PushNumpyOps(geo, box, ocos, comm)
        pipe = SL.SlabPipeline(geo, ops, comm, lean=True)
        delta = ops.alloc_shared(geo.slab_bytes)
        assert pipe.enable_push(delta) and pipe.bufB is None
        grid = H.Grid(box, n)
        ratio = M.velocity_to_offset_ratio(ocos)
        out_pos = ops.alloc(3 * geo.particles * 8)
        out_vel = ops.alloc(3 * geo.particles * 8)
        for _ in range(2):
            SL.unigrid_ic_slab(geo, ops, comm, seed, 2, 0, 0, grid, camb, ratio, np.float64, delta, pipe, out_pos, out_vel)
        np.savez(os.path.join(outdir, "rank%d.npz" % rank), pos=out_pos.numpy().view(np.float64).reshape(3, -1),
                 vel=out_vel.numpy().view(np.float64).reshape(3, -1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 16), (4, 16)])
def test_peer_memory_schedule_matches_single_process_oracle(tmp_path, world, n):
    """SlabPipeline.ic_push over gloo: noise with the transfer folded in, two pushed x-passes, psi_y / psi_z from the one
    exchanged x-transform of G with i k_y / i k_z applied on the load side of their y-passes, three barriers."""
    import torch.multiprocessing as mp
    from oracle import model as M
    port = _free_port()
    mp.spawn(_push_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    box, seed = 50.0, 8896131
    ocos = M.Cosmology(0.279, 0.721, 0.817, 99)
    opos, ovel, _, _ = M.unigrid_ic(n, box, seed, "fourier_parallel", CAMB_FILE, ocos)
    pos = np.concatenate([np.load(tmp_path / ("rank%d.npz" % r))["pos"] for r in range(world)], axis=1).T
    vel = np.concatenate([np.load(tmp_path / ("rank%d.npz" % r))["vel"] for r in range(world)], axis=1).T
    assert np.abs(vel - ovel).max() <= 1e-11 * np.abs(ovel).max()
    d = np.abs(pos - opos)
    assert np.minimum(d, box - d).max() <= 1e-11 * box


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, mode, outdir, overlap=False):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from genetic_b200 import slab as SL
    from genetic_b200 import host as H
    from oracle import model as M
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        box, seed = 50.0, 8896131
        ocos = M.Cosmology(0.279, 0.721, 0.817, 99)
        camb = M.CAMB(CAMB_FILE, ocos)
        geo = SL.SlabGeometry(n, world, rank, 8)
        ops = NumpyOps(geo, box, ocos)
        comm = SL.TorchComm()
        pipe = SL.SlabPipeline(geo, ops, comm, overlap=overlap)
        assert pipe.overlap == (overlap and world > 1)
        delta = ops.alloc(geo.slab_bytes)
        grid = H.Grid(box, n)
        ratio = M.velocity_to_offset_ratio(ocos)
        out_pos = ops.alloc(3 * geo.particles * 8)
        out_vel = ops.alloc(3 * geo.particles * 8)
        SL.unigrid_ic_slab(geo, ops, comm, seed, mode, 0, 0, grid, camb, ratio, np.float64, delta, pipe, out_pos, out_vel)
        np.savez(os.path.join(outdir, "rank%d.npz" % rank), pos=out_pos.numpy().view(np.float64).reshape(3, -1),
                 vel=out_vel.numpy().view(np.float64).reshape(3, -1))
        # round trip of the general transforms: to_real then to_fourier returns the (mirrored) field
        W = M.draw_white_noise(n, 4242, "fourier_parallel")
        f = ops.alloc(geo.slab_bytes)
        ops._c(f, (n, geo.P, geo.nc))[:] = W[:, geo.iy0:geo.iy0 + geo.P]
        pipe.to_real(f)
        real = f.numpy().view(np.float64).reshape(geo.P, n, n + 2)[:, :, :n].copy()
        pipe.to_fourier(f)
        np.savez(os.path.join(outdir, "rt%d.npz" % rank), real=real, four=ops._c(f, (n, geo.P, geo.nc)).copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,mode,overlap", [(2, 16, 2, False), (4, 16, 2, False), (2, 8, 0, False), (2, 8, 1, False),
                                                  (2, 16, 2, True), (4, 16, 2, True), (8, 16, 2, True), (8, 16, 1, False)])
def test_slab_pipeline_matches_single_process_oracle(tmp_path, world, n, mode, overlap):
    """overlap=True: the double-buffered Zeldovich loop (exchange of one component in flight during the neighbouring
    components' passes) must give the same particles as the sequential one."""
    import torch.multiprocessing as mp
    from oracle import model as M
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, mode, str(tmp_path), overlap), nprocs=world, join=True)
    box, seed = 50.0, 8896131
    ocos = M.Cosmology(0.279, 0.721, 0.817, 99)
    name = {0: "real", 1: "fourier_serial", 2: "fourier_parallel"}[mode]
    opos, ovel, _, _ = M.unigrid_ic(n, box, seed, name, CAMB_FILE, ocos)
    pos = np.concatenate([np.load(tmp_path / ("rank%d.npz" % r))["pos"] for r in range(world)], axis=1).T
    vel = np.concatenate([np.load(tmp_path / ("rank%d.npz" % r))["vel"] for r in range(world)], axis=1).T
    assert np.abs(vel - ovel).max() <= 1e-11 * np.abs(ovel).max()
    d = np.abs(pos - opos)
    assert np.minimum(d, box - d).max() <= 1e-11 * box
    # general transforms
    W = M.draw_white_noise(n, 4242, "fourier_parallel")
    ref_real = M.to_real(W)
    real = np.concatenate([np.load(tmp_path / ("rt%d.npz" % r))["real"] for r in range(world)], axis=0)
    assert np.abs(real - ref_real).max() <= 1e-12 * np.abs(ref_real).max()
    four = np.concatenate([np.load(tmp_path / ("rt%d.npz" % r))["four"] for r in range(world)], axis=1)
    ref_four = M.to_fourier(ref_real)
    assert np.abs(four - ref_four).max() <= 1e-12 * np.abs(ref_four).max()


def test_slab_geometry_rejects_bad_sizes():
    from genetic_b200 import slab as SL
    from genetic_b200.lib import GxError
    with pytest.raises(GxError):
        SL.SlabGeometry(16, 3, 0)
    with pytest.raises(GxError):
        SL.SlabGeometry(24, 2, 0)
    g = SL.SlabGeometry(16, 4, 3)
    assert (g.P, g.iy0, g.x0, g.id0, g.particles) == (4, 12, 12, 12 * 256, 4 * 256)
    assert g.slab_bytes * 4 == 2 * 16 * 16 * 9 * 8
