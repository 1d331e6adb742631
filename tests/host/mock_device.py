"""TEST INFRASTRUCTURE -- a numpy stand-in for the C ABI, used ONLY by the CPU (`-m "not gpu"`) tests of the Python host
logic (genetic_b200/host.py, multilevel.py): the control flow of the reference's classes can then be checked against
the oracle's independently written flow without a GPU. Nothing in genetic_b200/ imports this; the product path has no
CPU fallback (genetic_b200/lib.py raises when the CUDA library is missing).

"Device pointers" are integer handles into a dict of byte buffers; fields use the device layout (padded half-complex).
Only the entry points the multi-level host layer uses are provided.
"""
import ctypes as C

import numpy as np

from oracle import model as M, multilevel as ML


def _addr(p):
    if isinstance(p, int):
        return p
    if p is None:
        return 0
    if hasattr(p, "value"):
        return p.value or 0
    if hasattr(p, "_obj"):
        return _addr(p._obj)
    return C.cast(p, C.c_void_p).value


class MockSpectrum:
    """Stands in for cosmology::CAMB on the device: covariance through the oracle's C spline."""

    def __init__(self, cosmology, filename, species="all", dtype=np.float64):
        self.camb = M.CAMB(filename, M.Cosmology(cosmology.OmegaM0, cosmology.OmegaLambda0, cosmology.sigma8,
                                                 cosmology.redshift, ns=cosmology.ns, hubble=cosmology.hubble))
        self.species = species
        self.handle = self
        self.amplitude = self.camb.amplitude

    def covariance(self, L, n):
        return self.camb.covariance(L, n, self.species)

    def getPowerSpectrumForGrid(self, grid):
        import genetic_b200 as gb
        return gb.Field(grid, self.covariance(float(grid.thisGridSize), grid.size).astype(np.complex128), fourier=True)


class MockLib:
    def __init__(self):
        self.mem = {}
        self.next = 0x10000
        self.cdll = self
        self.launches = 0

    # ---- memory ---------------------------------------------------------------------------------------------
    def gx_field_bytes(self, n, dt):
        return 2 * n * n * (n // 2 + 1) * (8 if dt == 1 else 4)

    def gx_malloc(self, pp, nbytes):
        h = self.next
        self.next += (nbytes + 0xfff) & ~0xfff
        self.mem[h] = np.zeros(nbytes, dtype=np.uint8)
        pp._obj.value = h
        return 0

    def gx_free(self, p):
        self.mem.pop(_addr(p), None)
        return 0

    def gx_memset_zero(self, p, nbytes, st):
        self.mem[_addr(p)][:nbytes] = 0
        return 0

    def gx_synchronize(self, st):
        return 0

    def _host(self, p, nbytes):
        return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(_addr(p)))

    def gx_memcpy_h2d(self, d, h, nbytes, st):
        self.mem[_addr(d)][:nbytes] = self._host(h, nbytes)
        return 0

    def gx_memcpy_d2d(self, d, s, nbytes, st):
        self.mem[_addr(d)][:nbytes] = self.mem[_addr(s)][:nbytes]
        return 0

    def gx_memcpy_d2h(self, h, d, nbytes, st):
        base, off = self._resolve(d)
        self._host(h, nbytes)[:] = self.mem[base][off:off + nbytes]
        return 0

    def _resolve(self, p):
        """An address inside a mock allocation -> (handle, byte offset)."""
        a = _addr(p)
        if a in self.mem:
            return a, 0
        base = max(h for h in list(self.mem) if h <= a)    # (a Field finaliser may free a buffer meanwhile)
        assert a - base < len(self.mem[base])
        return base, a - base

    # ---- particle evaluation (zeldovich.hpp:87-119, generator.hpp:49-53, grid.hpp:519-531,656-669) -------------------
    def gx_emit_cells(self, psi, n, dt, axis, dx, lower, domain, ratio, cells, count, out_pos, out_vel, odt, stride, st):
        assert dt == 1
        ids = self.mem[_addr(cells)][:8 * count].view(np.int64)
        return self._emit(psi, n, axis, dx, lower, domain, ratio, ids, out_pos, out_vel, odt, stride)

    def gx_emit_component(self, psi, n, dt, axis, dx, lower, domain, ratio, out_pos, out_vel, odt, stride, id0, count, st):
        assert dt == 1
        return self._emit(psi, n, axis, dx, lower, domain, ratio, np.arange(id0, id0 + count, dtype=np.int64), out_pos,
                          out_vel, odt, stride)

    def _emit(self, psi, n, axis, dx, lower, domain, ratio, ids, out_pos, out_vel, odt, stride):
        p = self.R(psi, n).reshape(-1)[ids]
        c = np.unravel_index(ids, (n, n, n))[axis].astype(np.float64)
        q = np.fmod(p + ((c * dx + lower) + dx / 2), domain)
        q = np.where(q < 0, q + domain, q)
        O = np.float64 if odt == 1 else np.float32
        for ptr, val in ((out_pos, q), (out_vel, p * ratio)):
            if _addr(ptr):
                base, off = self._resolve(ptr)
                self.mem[base][off:].view(O)[::stride][:len(ids)] = val.astype(O)
        return 0

    # views of a field buffer (f64 only)
    def F(self, p, n):
        return self.mem[_addr(p)].view(np.complex128).reshape(n, n, n // 2 + 1)

    def R(self, p, n):
        return self.mem[_addr(p)].view(np.float64).reshape(n, n, n + 2)[:, :, :n]

    def gx_field_upload(self, d, h, n, dt, layout, st):
        assert dt == 1
        if layout == 1:
            self.mem[_addr(d)][:] = self._host(h, self.gx_field_bytes(n, dt))
        else:
            self.R(d, n)[:] = self._host(h, 8 * n ** 3).view(np.float64).reshape(n, n, n)
        return 0

    def gx_field_download(self, h, d, n, dt, layout, st):
        if layout == 1:
            self._host(h, self.gx_field_bytes(n, dt))[:] = self.mem[_addr(d)]
        else:
            self._host(h, 8 * n ** 3).view(np.float64).reshape(n, n, n)[:] = self.R(d, n)
        return 0

    # ---- transforms and spectral operators ------------------------------------------------------------------------
    def gx_fft_forward(self, p, n, dt, st):
        self.F(p, n)[:] = M.to_fourier(self.R(p, n).copy())
        return 0

    def gx_fft_inverse(self, p, n, dt, st):
        self.R(p, n)[:] = M.to_real(self.F(p, n).copy())
        return 0

    def gx_noise_draw(self, p, n, dt, seed, mode, reverse, off, st):
        from oracle import crng
        o = off._obj if off is not None else None
        if o is None or o.value == 0:           # first level: new gsl stream; later levels continue it
            self._stream = crng.Stream(seed)
        name = {0: "real", 1: "fourier_reverse" if reverse else "fourier_serial", 2: "fourier_parallel"}[mode]
        # parallel mode: one stream per shell, seeded from the base seed on every level (randomfieldgenerator.hpp:228-229)
        self.F(p, n)[:] = M.draw_white_noise(n, seed, name, stream=None if mode == 2 else self._stream)
        if o is not None:
            o.value += 1                         # (the device counts raw outputs; any non-zero value continues)
        return 0

    def gx_apply_transfer(self, p, n, dt, L, spec, power, st):
        self.F(p, n)[:] = M.apply_transfer(self.F(p, n), spec.covariance(L, n), power)
        return 0

    def gx_apply_transfer_field(self, p, cov, n, dt, power, st):
        self.F(p, n)[:] = M.apply_transfer(self.F(p, n), self.F(cov, n).real.copy(), power)
        return 0

    def gx_covariance_field(self, p, n, dt, L, spec, st):
        self.F(p, n)[:] = spec.covariance(L, n)
        return 0

    def gx_apply_filter(self, p, n, dt, L, fac, nf, st):
        g = ML.Grid(L, n)
        factors = [(fac[i].kind, fac[i].kcut, fac[i].power) for i in range(nf)]
        self.F(p, n)[:] = ML.apply_filter(self.F(p, n), g, factors)
        return 0

    def gx_unit_variance(self, p, n, dt, st):
        F = self.F(p, n)
        a = np.abs(F)
        ind = M.independent_mask(n)
        F[ind & (a > 0)] /= a[ind & (a > 0)]
        return 0

    def gx_reverse_small_k(self, p, n, dt, L, kmax, st):
        k2 = ML.k_magnitude(ML.Grid(L, n)) ** 2
        F = self.F(p, n)
        sel = M.independent_mask(n) & (k2 < kmax * kmax) & (k2 != 0)
        F[sel] = -F[sel]
        return 0

    def gx_scale(self, p, n, dt, a, st):
        self.mem[_addr(p)].view(np.float64)[:] *= a
        return 0

    def gx_axpy(self, y, x, n, dt, a, st):
        self.mem[_addr(y)].view(np.float64)[:] += a * self.mem[_addr(x)].view(np.float64)
        return 0

    def gx_axpy_fourier(self, y, x, n, dt, a, st):
        ind = M.independent_mask(n)
        self.F(y, n)[ind] += a * self.F(x, n)[ind]
        return 0

    def gx_mul_mask(self, p, mask, n, dt, st):
        self.R(p, n)[:] *= self.mem[_addr(mask)][: n ** 3].reshape(n, n, n)
        return 0

    def gx_power_spectrum_bins(self, p, cov, n, dt, L, nbins, out, st):
        """cosmology::dumpPowerSpectrum's binning (parameters.hpp:112-170) on the host mirrors: every (ix, iy, iz) in
        [-n/2, n/2]^3 visited as the reference does (the +-Nyquist planes twice)."""
        import ctypes as C
        import math

        def mirrored(F):
            F = F.copy()
            a = np.arange(n)
            dep = (a[:, None] > n // 2) | (((a[:, None] == 0) | (a[:, None] == n // 2)) & (a[None, :] > n // 2))
            ia, ib = np.nonzero(dep)
            for kz in (0, n // 2):
                F[ia, ib, kz] = np.conj(F[(-ia) % n, (-ib) % n, kz])
            return F

        field = mirrored(self.F(p, n))
        P0 = mirrored(self.F(cov, n)) if cov else np.zeros_like(field)
        kmax, kmin = math.pi / L * n, 2.0 * math.pi / L
        dklog, kw = math.log10(kmax / kmin) / nbins, 2.0 * math.pi / L
        i1 = np.arange(-(n // 2), n // 2 + 1)
        ix, iy, iz = np.meshgrid(i1, i1, i1, indexing="ij")
        neg = iz < 0
        a = np.where(neg, (-ix) % n, ix % n)
        b = np.where(neg, (-iy) % n, iy % n)
        c = np.abs(iz)
        k = np.sqrt((ix * ix + iy * iy + iz * iz).astype(np.float64)) * kw
        ok = (k >= kmin) & (k < kmax)
        idx = (1.0 / dklog * np.log10(k[ok] / kmin)).astype(np.int64)
        bins = np.zeros((4, nbins))
        bins[0] = np.bincount(idx, minlength=nbins)
        bins[1] = np.bincount(idx, weights=k[ok], minlength=nbins)
        bins[2] = np.bincount(idx, weights=np.abs(field[a[ok], b[ok], c[ok]]) ** 2, minlength=nbins)
        bins[3] = np.bincount(idx, weights=P0[a[ok], b[ok], c[ok]].real, minlength=nbins)
        dst = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_double)), shape=(4 * nbins,))
        dst[:] = bins.reshape(-1)
        return 0

    def gx_dot_fourier(self, a, b, n, dt, out, st):
        out._obj.value = float(np.sum(M.mode_weights(n) * (np.conj(self.F(a, n)) * self.F(b, n)).real))
        return 0

    def gx_dot_real(self, a, b, n, dt, out, st):
        out._obj.value = float(np.sum(self.R(a, n) * self.R(b, n)))
        return 0

    def gx_set_mode(self, p, n, dt, kx, ky, kz, re, im):
        self.F(p, n)[kx % n, ky % n, kz] = complex(re, im)
        return 0

    def gx_zeldovich_fourier(self, d, psi, n, dt, L, axis, st):
        self.F(psi, n)[:] = M.zeldovich_offsets_fourier(self.F(d, n), L)[axis]
        return 0

    def gx_zeldovich_offset(self, d, psi, n, dt, L, axis, st):
        """The fused form: psi <- c2r(i k_axis delta / k^2), real space."""
        self.R(psi, n)[:] = M.to_real(M.zeldovich_offsets_fourier(self.F(d, n), L)[axis])
        return 0

    def gx_scale_inv_k2(self, p, n, dt, L, scale, potential_form, st):
        g = ML.Grid(L, n)
        k2 = ML.k_magnitude(g) ** 2
        ind = M.independent_mask(n)
        F = self.F(p, n)
        with np.errstate(divide="ignore", invalid="ignore"):
            if potential_form:
                ky0 = (np.arange(n) == 0)[None, :, None]
                F[:] = np.where(ind, np.where(ky0, 0.0, scale * F * (1.0 / k2)), F)
            else:
                F[:] = np.where(ind, np.where(k2 == 0, 0.0, scale * F / k2), F)
        return 0

    # ---- cross-grid ----------------------------------------------------------------------------------------------
    @staticmethod
    def _g(d):
        d = d._obj
        return ML.Grid(d.domain_size, d.n, d.cell_size, tuple(d.offset_lower))

    def gx_interp_add(self, dst, gd, src, gs, dt, st):
        gd, gs = self._g(gd), self._g(gs)
        self.R(dst, gd.n)[:] = ML.add_from_different_grid(gd, self.R(dst, gd.n).copy(), gs, self.R(src, gs.n).copy())
        return 0

    def gx_deinterp_add(self, dst, gc, src, gf, dt, scale, st):
        gc, gf = self._g(gc), self._g(gf)
        self.R(dst, gc.n)[:] += ML.deinterpolate_into(gc, gf, self.R(src, gf.n).copy(), scale)
        return 0

    def gx_zero_window(self, p, g, dt, w, inside, st):
        g, w = self._g(g), self._g(w)
        m = ML.window_mask(g, w)
        r = self.R(p, g.n)
        r[m if inside else ~m] = 0.0
        return 0

    def gx_masked_sum(self, p, mask, n, dt, out, st):
        m = self.mem[_addr(mask)][: n ** 3].reshape(n, n, n).astype(bool)
        out._obj.value = float(self.R(p, n)[m].sum())
        return 0

    def gx_masked_shift_scale(self, p, mask, n, dt, shift, scale, st):
        m = self.mem[_addr(mask)][: n ** 3].reshape(n, n, n).astype(bool)
        r = self.R(p, n)
        r[:] = np.where(m, (r - shift) * scale, 0.0)
        return 0


def install(monkeypatch):
    """Route genetic_b200's loader to the mock and replace the device spectrum class."""
    import genetic_b200 as gb
    from genetic_b200 import lib as L, host
    mock = MockLib()
    monkeypatch.setattr(L, "_lib", mock)
    monkeypatch.setattr(L, "load", lambda: mock)
    monkeypatch.setattr(host, "CAMB", MockSpectrum)
    monkeypatch.setattr(gb, "CAMB", MockSpectrum)
    return mock
