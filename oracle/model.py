"""TEST INFRASTRUCTURE -- numpy restatement of genetIC's unigrid field pipeline (not product code).

Each function cites the reference file:line it follows (paths relative to
/root/reference/genetIC/src).  Scalar cosmology / spline / per-mode transcendental work is done in
oracle/gx_oracle.c with the host libm (exactly what the reference calls); numpy is used for layout,
FFTs (pocketfft) and element-wise IEEE arithmetic.

Storage convention (tools/numerics/fourier.hpp:300-319, 471-476): a field is 2*N*N*(N/2+1) reals.
Fourier space: complex [kx][ky][kz>=0]. Real space: the first N^3 reals, C order [x][y][z].
"""
import ctypes
import struct

import numpy as np

from . import lib
from . import crng


# --------------------------------------------------------------------------------------------
# layout helpers
# --------------------------------------------------------------------------------------------
def signed_modes(n):
    """Storage index -> signed integer wavenumber; index N/2 is +Nyquist (fourier.hpp:300-319)."""
    i = np.arange(n)
    return np.where(i <= n // 2, i, i - n)


def independent_mask(n):
    """Storage slots reached by the half-space mode iteration (fourier.hpp:110-151)."""
    h = n // 2
    m = np.ones((n, n, h + 1), dtype=bool)
    ix = np.arange(n)[:, None]
    iy = np.arange(n)[None, :]
    dead = (ix > h) | (((ix == 0) | (ix == h)) & (iy > h))
    for iz in (0, h):
        m[:, :, iz] = ~dead
    return m


def mode_weights(n):
    """Weight with which each independent slot enters an accumulation (fourier.hpp:116-147):
    2 when the mode has a distinct conjugate partner that the iteration skips, 1 otherwise."""
    h = n // 2
    w = np.full((n, n, h + 1), 2.0)
    # modes reached directly with kx in [0,N/2]; weights by the reference's rule
    kx = signed_modes(n)[:, None, None]
    ky = signed_modes(n)[None, :, None]
    kz = np.arange(h + 1)[None, None, :]
    # The iteration visits (kx>=0, ky, kz signed). A slot holds either the visited mode itself or its conjugate;
    # weights depend on |kx|,|ky|,|kz| membership in {0, N/2} only, so evaluate on absolute values.
    ax, ay, az = np.abs(kx), np.abs(ky), np.abs(kz)
    selfx = (ax == 0) | (ax == h)
    selfy = (ay == 0) | (ay == h)
    selfz = (az == 0) | (az == h)
    wx = np.where(selfx, 1, 2)
    # cellWeightY doubles when ky_lower==0 (kx self-conj) and ky not self-conj
    wy = np.where(selfx & ~selfy, wx * 2, wx)
    # cellWeightZ doubles when kz_lower==0 (kx,ky self-conj) and kz not self-conj
    wz = np.where(selfx & selfy & ~selfz, wy * 2, wy)
    w = wz.astype(np.float64) * np.ones_like(w)
    w[~independent_mask(n)] = 0.0
    return w


def mirror(F):
    """ensureFourierModesAreMirrored (fourier.hpp:331-357), in place on complex [n,n,n/2+1].
    Sequential semantics matter only for the self-mapping entries (their imaginary part flips sign)."""
    n = F.shape[0]
    h = n // 2
    for iz in (0, h):
        P = F[:, :, iz]
        for kx in range(0, h + 1):
            for ky in range(h, -h, -1):
                P[(-kx) % n, (-ky) % n] = np.conj(P[kx % n, ky % n])
    return F


def mirror_fast(F):
    """Vectorised equivalent of mirror() (validated against it in tests)."""
    n = F.shape[0]
    h = n // 2
    for iz in (0, h):
        P = F[:, :, iz]
        src = P.copy()
        # rows 0<kx<h: dest (-kx,-ky) <- conj(src(kx,ky)) for all ky
        kx = np.arange(1, h)
        if len(kx):
            P[(-kx) % n, :] = np.conj(src[kx][:, (-np.arange(n)) % n])
        for kxs in (0, h):
            row = src[kxs].copy()
            # ky = h (self), h-1..1 -> dest -ky, 0 (self), -1..-(h-1) -> dest +ky (restores itself)
            new = row.copy()
            new[h] = np.conj(row[h])
            ky = np.arange(1, h)
            new[(-ky) % n] = np.conj(row[ky])
            new[0] = np.conj(row[0])
            P[kxs] = new
    return F


def to_fourier(real):
    """performTransform forward (fourier.hpp:479-518): r2c, divided by N^1.5."""
    n = real.shape[0]
    norm = float(n) ** 1.5
    return np.fft.rfftn(real) / norm


def to_real(F, do_mirror=True):
    """performTransform inverse: mirror, unnormalised c2r, divide by N^1.5 (fourier.hpp:497-513)."""
    n = F.shape[0]
    G = F.copy()
    if do_mirror:
        mirror_fast(G)
    norm = float(n) ** 1.5
    return np.fft.irfftn(G, s=(n, n, n), axes=(0, 1, 2)) * (float(n) ** 3) / norm


# --------------------------------------------------------------------------------------------
# cosmology (host scalars and per-mode covariance through the C oracle)
# --------------------------------------------------------------------------------------------
def read_camb(filename, is_float=False):
    """camb.hpp:224-268 + io/input.hpp:47-95. Returns k, T_dm, T_baryon, T_all (normalised by input[1])."""
    rows = []
    with open(filename) as f:
        lines = f.read().split("\n")
    first = True
    for line in lines:
        toks = line.split()
        if first:
            first = False
            try:
                vals = [float(t) for t in toks]
            except ValueError:
                continue  # header line ignored
            rows.extend(vals)
            continue
        rows.extend(float(t) for t in toks)
    second = lines[1].split()
    ncol = len(second)
    if ncol not in (7, 13):
        raise RuntimeError("CAMB transfer file doesn't have a sensible number of columns")
    a = np.array(rows, dtype=np.float64)
    a = a[: (len(a) // ncol) * ncol].reshape(-1, ncol)
    norm = a.reshape(-1)[1]
    keep = a[:, 0] > 0
    a = a[keep]
    if is_float:
        # kInterpolationPoints.push_back(CoordinateType(..)); CoordinateType(col)/CoordinateType(input[1]) (camb.hpp:254-261)
        f = np.float32
        nf = f(norm)
        return (a[:, 0].astype(f).astype(np.float64), (a[:, 1].astype(f) / nf).astype(np.float64),
                (a[:, 2].astype(f) / nf).astype(np.float64), (a[:, 6].astype(f) / nf).astype(np.float64))
    return a[:, 0].copy(), a[:, 1] / norm, a[:, 2] / norm, a[:, 6] / norm


SPECIES = {"dm": 0, "baryon": 1, "all": 2}


class Cosmology:
    def __init__(self, Om, Ol, s8, zin, ns=0.96, hubble=0.701, Ob=0.0):
        self.Om, self.Ol, self.s8, self.z, self.ns, self.hubble, self.Ob = Om, Ol, s8, zin, ns, hubble, Ob
        self.a = 1.0 / (zin + 1.0)  # ic.hpp:339-342


class CAMB:
    """cosmology::CAMB<T> (camb.hpp:173-332) for T=double (is_float=False) or T=float."""

    def __init__(self, filename, cosmo, is_float=False):
        L = lib()
        k, tdm, tb, tall = read_camb(filename, is_float)
        self.is_float = is_float
        self._k = np.ascontiguousarray(k)
        self._t = [np.ascontiguousarray(x) for x in (tdm, tb, tall)]
        L.orc_camb_create.restype = ctypes.c_void_p
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self.h = ctypes.c_void_p(L.orc_camb_create(p(self._k), p(self._t[0]), p(self._t[1]), p(self._t[2]), len(k)))
        self.cosmo = cosmo
        self.ns = cosmo.ns
        if is_float:
            L.orc_camb_amplitude_f.restype = ctypes.c_float
            c_f = ctypes.c_float
            self.amplitude = L.orc_camb_amplitude_f(self.h, c_f(cosmo.Om), c_f(cosmo.Ol), c_f(cosmo.s8), c_f(cosmo.ns),
                                                    c_f(np.float32(1.) / np.float32(np.float32(cosmo.z) + np.float32(1.))))
        else:
            L.orc_camb_amplitude_d.restype = ctypes.c_double
            c_d = ctypes.c_double
            self.amplitude = L.orc_camb_amplitude_d(self.h, c_d(cosmo.Om), c_d(cosmo.Ol), c_d(cosmo.s8), c_d(cosmo.ns),
                                                    c_d(cosmo.a))

    def transfer(self, k, species="all"):
        L = lib()
        L.orc_camb_transfer.restype = ctypes.c_double
        return L.orc_camb_transfer(self.h, SPECIES[species], ctypes.c_double(k))

    def spline_tables(self, species="all"):
        L = lib()
        m = L.orc_camb_spline_export(self.h, SPECIES[species], None, None, None)
        lx, ly, c = (np.empty(m) for _ in range(3))
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        L.orc_camb_spline_export(self.h, SPECIES[species], p(lx), p(ly), p(c))
        return lx, ly, c

    def covariance(self, box, n, species="all"):
        """getPowerSpectrumForGridUncached (camb.hpp:101-119): real array [n,n,n/2+1] (0 on redundant slots)."""
        L = lib()
        if self.is_float:
            out = np.empty((n, n, n // 2 + 1), dtype=np.float32)
            L.orc_covariance_field_f(self.h, SPECIES[species], ctypes.c_float(self.amplitude), ctypes.c_float(self.ns),
                                     ctypes.c_float(box), n, out.ctypes.data_as(ctypes.c_void_p))
        else:
            out = np.empty((n, n, n // 2 + 1), dtype=np.float64)
            L.orc_covariance_field_d(self.h, SPECIES[species], ctypes.c_double(self.amplitude),
                                     ctypes.c_double(self.ns), ctypes.c_double(box), n,
                                     out.ctypes.data_as(ctypes.c_void_p))
        return out

    def __del__(self):
        try:
            lib().orc_camb_destroy(self.h)
        except Exception:
            pass


def growth_factor(Om, Ol, a):
    L = lib()
    L.orc_growth_factor_d.restype = ctypes.c_double
    return L.orc_growth_factor_d(ctypes.c_double(Om), ctypes.c_double(Ol), ctypes.c_double(a))


def velocity_to_offset_ratio(cosmo, is_float=False):
    """parameters.hpp:83-106"""
    L = lib()
    if is_float:
        L.orc_velocity_to_offset_ratio_f.restype = ctypes.c_float
        a = np.float32(1.) / np.float32(np.float32(cosmo.z) + np.float32(1.))
        return L.orc_velocity_to_offset_ratio_f(ctypes.c_float(cosmo.Om), ctypes.c_float(cosmo.Ol), ctypes.c_float(a))
    L.orc_velocity_to_offset_ratio_d.restype = ctypes.c_double
    return L.orc_velocity_to_offset_ratio_d(ctypes.c_double(cosmo.Om), ctypes.c_double(cosmo.Ol),
                                            ctypes.c_double(cosmo.a))


def apply_transfer(F, cov, power=0.5):
    """Field::applyTransferFunction (field.hpp:394-412), independent slots only. Returns new array."""
    L = lib()
    n = F.shape[0]
    if F.dtype == np.complex64:
        G = np.ascontiguousarray(F).copy()
        L.orc_apply_transfer_f(G.ctypes.data_as(ctypes.c_void_p), np.ascontiguousarray(cov, dtype=np.float32).ctypes.data_as(ctypes.c_void_p),
                               n, ctypes.c_double(power))
    else:
        G = np.ascontiguousarray(F, dtype=np.complex128).copy()
        L.orc_apply_transfer_d(G.ctypes.data_as(ctypes.c_void_p), np.ascontiguousarray(cov, dtype=np.float64).ctypes.data_as(ctypes.c_void_p),
                               n, ctypes.c_double(power))
    return G


# --------------------------------------------------------------------------------------------
# white noise
# --------------------------------------------------------------------------------------------
def draw_white_noise(n, seed, mode, dtype=np.float64, stream=None):
    """RandomFieldGenerator::draw for one level (randomfieldgenerator.hpp:122-139).
    mode: 'real' | 'fourier_serial' | 'fourier_parallel' | 'fourier_reverse'. Returns Fourier-space complex array.
    `stream` continues a previous level's stream (serial modes draw all levels from one stream)."""
    is_float = np.dtype(dtype) == np.float32
    st = stream if stream is not None else crng.Stream(seed)
    if mode == "real":
        g = crng.draw_real(st, n, is_float)
        F = to_fourier(g.astype(dtype))
        F[0, 0, 0] = complex(0.0, F[0, 0, 0].imag)  # tools::set_zero(fieldData[0]) zeroes Re(DC) only (:195)
        return F
    parallel = mode == "fourier_parallel"
    reverse = mode == "fourier_reverse"
    return crng.draw_fourier(st, n, parallel, reverse, is_float)


# --------------------------------------------------------------------------------------------
# Zeldovich
# --------------------------------------------------------------------------------------------
def zeldovich_offsets_fourier(F, box):
    """calculateOffsetFields, Fourier-space gradient (zeldovich.hpp:152-181), via generateNewFourierFields
    (fourier.hpp:253-274): three new zero fields, independent slots filled. Returns (psi_x, psi_y, psi_z)."""
    n = F.shape[0]
    T = F.real.dtype.type
    kMin = T(2.0 * np.pi / box) if T is np.float64 else T(np.float64(2.0) * np.pi / np.float32(box))
    ks = signed_modes(n).astype(T) * kMin
    kx = ks[:, None, None]
    ky = ks[None, :, None]
    kz = (np.arange(n // 2 + 1).astype(T) * kMin)[None, None, :]
    kfft = kx * kx + ky * ky + kz * kz
    nyq = T(n // 2) * kMin
    with np.errstate(divide="ignore", invalid="ignore"):
        rx_re = -F.imag / kfft
        rx_im = F.real / kfft
    ind = independent_mask(n)
    out = []
    for kk in (kx, ky, kz):
        kkb = np.broadcast_to(kk, kfft.shape)
        re = rx_re * kkb
        im = rx_im * kkb
        zero = (np.abs(kkb) == nyq) | (kfft == 0)
        # NB the reference tests `kx == nyquist` on iteration coordinates, which reach +N/2 only;
        # the storage slot with index N/2 is that mode (or its conjugate), hence |k| == nyq here.
        psi = (re + 1j * im).astype(F.dtype)
        psi[zero] = 0
        psi[~ind] = 0
        out.append(psi)
    return out


def particles_unigrid(psi_real, box, cosmo, eps_norm=0.01075, is_float=False):
    """ZeldovichParticleEvaluator::getParticle (zeldovich.hpp:87-119, generator.hpp:49-53, grid.hpp:519-531,
    656-669). psi_real: three [n,n,n] arrays. Returns pos[n^3,3], vel[n^3,3], mass, eps in the field type."""
    T = np.float32 if is_float else np.float64
    n = psi_real[0].shape[0]
    dx = T(box) / T(n)
    ratio = T(velocity_to_offset_ratio(cosmo, is_float))
    idx = np.arange(n, dtype=np.int64)
    cx = (idx.astype(T) * dx + T(0)) + dx / T(2)
    cent = [np.broadcast_to(cx[:, None, None], (n, n, n)), np.broadcast_to(cx[None, :, None], (n, n, n)),
            np.broadcast_to(cx[None, None, :], (n, n, n))]
    pos = np.empty((n ** 3, 3), dtype=T)
    vel = np.empty((n ** 3, 3), dtype=T)
    for j in range(3):
        p = psi_real[j].astype(T)
        vel[:, j] = (p * ratio).reshape(-1)
        q = (p + cent[j]).reshape(-1)
        q = np.fmod(q, T(box))
        q = np.where(q < 0, q + T(box), q)
        pos[:, j] = q
    boxMass = T(T(27.744948) * T(cosmo.Om) * T(np.float64(T(box)) ** 3.0)) if is_float else 27.744948 * cosmo.Om * box ** 3.0
    cellMassFrac = T(np.float64(dx / T(box)) ** 3.0)
    mass = T(boxMass) * cellMassFrac
    eps = dx * T(1.0) * T(eps_norm)
    return pos, vel, mass, eps


# --------------------------------------------------------------------------------------------
# tipsy golden files (io/tipsy.hpp:20-27,62-71,207-271)
# --------------------------------------------------------------------------------------------
def read_tipsy(filename):
    with open(filename, "rb") as f:
        raw = f.read()
    a, n, ndim, ngas, ndark, nstar = struct.unpack("<diiiii", raw[:28])
    off = 32  # header struct is padded to 32 bytes
    assert ngas == 0 and nstar == 0
    d = np.frombuffer(raw, dtype=np.float32, count=ndark * 9, offset=off).reshape(ndark, 9)
    return {"a": a, "n": n, "mass": d[:, 0], "pos": d[:, 1:4], "vel": d[:, 4:7], "eps": d[:, 7]}


def to_tipsy_units(pos, vel, mass, eps, box, cosmo, npart):
    """TipsyOutput::operator() unit conversions (tipsy.hpp:207-249), float32 results like the file."""
    tot_mass = np.float64(0)
    m = np.float64(mass)
    # tot_mass accumulates sequentially in FloatType (double build: double)
    tot_mass = np.cumsum(np.full(npart, m))[-1]
    mass_factor = cosmo.Om / tot_mass
    pos_factor = 1.0 / box
    dKpcUnit = box * 1000 / cosmo.hubble
    dMsolUnit = 1e10 / cosmo.hubble / mass_factor
    dKmsUnit = np.sqrt(4.3022682e-6 * dMsolUnit / dKpcUnit)
    vel_factor = cosmo.a ** -0.5 / dKmsUnit
    out = {
        "pos": (pos.astype(np.float64) * pos_factor - 0.5).astype(np.float32),
        "vel": (vel.astype(np.float64) * vel_factor).astype(np.float32),
        "mass": np.float32(m * mass_factor),
        "eps": np.float32(np.float64(eps) * pos_factor),
    }
    return out


# --------------------------------------------------------------------------------------------
# whole unigrid chain
# --------------------------------------------------------------------------------------------
def unigrid_ic(n, box, seed, mode, camb_file, cosmo, dtype=np.float64, return_intermediate=False):
    """ICGenerator::done on a unigrid (ic.hpp:1633-1648, 1275-1331)."""
    is_float = np.dtype(dtype) == np.float32
    ctype = np.complex64 if is_float else np.complex128
    W = draw_white_noise(n, seed, mode, dtype).astype(ctype)
    camb = CAMB(camb_file, cosmo, is_float)
    cov = camb.covariance(box, n, "all")
    D = apply_transfer(W, cov, 0.5)
    psis_k = zeldovich_offsets_fourier(D, dtype(box) if is_float else box)
    psis = [to_real(p).astype(dtype) for p in psis_k]
    pos, vel, mass, eps = particles_unigrid(psis, box, cosmo, is_float=is_float)
    if return_intermediate:
        return dict(white=W, cov=cov, delta_k=D, psi_k=psis_k, psi=psis, pos=pos, vel=vel, mass=mass, eps=eps,
                    camb=camb)
    return pos, vel, mass, eps


def chi2_unigrid(W):
    """MultiLevelField::getChi2 on one level (multilevelfield.hpp:461-487): filters are identity, so
    chi^2 = sum_modes w |W_k|^2."""
    n = W.shape[0]
    return float(np.sum(mode_weights(n) * (W.real ** 2 + W.imag ** 2)))
