"""Python host layer mirroring the reference's class surface for the hot path.

Every class forwards to the C ABI (include/genetic_b200.h); nothing here computes field data on the CPU.
Names and argument meaning follow the reference (paths relative to genetIC/src/):
  grids::Grid                         simulation/grid/grid.hpp:84-130
  fields::Field                       simulation/field/field.hpp:181-892
  fields::RandomFieldGenerator        simulation/field/randomfieldgenerator.hpp:25-263
  cosmology::CAMB / PowerLaw          cosmology/camb.hpp:150-332
  cosmology::CosmologicalParameters   cosmology/parameters.hpp:20-33
  particle::ZeldovichParticleGenerator simulation/particles/zeldovich.hpp:132-304
  filters::LowPassFermiFilter etc.    simulation/filters/filter.hpp:209-317
"""
import ctypes as C
import math

import numpy as np

from . import lib as L


def _dt(dtype):
    d = np.dtype(dtype)
    if d == np.float64:
        return L.F64
    if d == np.float32:
        return L.F32
    raise L.GxError("dtype must be float32 or float64")


class CosmologicalParameters:
    """cosmology::CosmologicalParameters (parameters.hpp:20-33) with ICGenerator's defaults (ic.hpp:201-207)."""

    def __init__(self, OmegaM0=0.3, OmegaLambda0=None, sigma8=None, redshift=0.0, ns=0.96, hubble=0.701,
                 OmegaBaryons0=0.0, TCMB=2.725):
        self.OmegaM0, self.OmegaLambda0, self.sigma8 = OmegaM0, OmegaLambda0, sigma8
        self.ns, self.hubble, self.OmegaBaryons0, self.TCMB = ns, hubble, OmegaBaryons0, TCMB
        self.setRedshift(redshift)

    def setRedshift(self, z):  # ICGenerator::setZ0, ic.hpp:339-342
        self.redshift = z
        self.scalefactor = 1.0 / (z + 1.0)

    def zeldovichVelocityToOffsetRatio(self, dtype=np.float64):
        """parameters.hpp:83-106"""
        T = np.dtype(dtype).type
        Om, Ol, a = T(self.OmegaM0), T(self.OmegaLambda0), T(T(1.) / T(T(self.redshift) + T(1.)))
        if T is np.float64:
            return float(1.0 * 100. * math.sqrt(Om / (a * a * a) + Ol) * math.sqrt(a))
        return float(T(T(T(1.0) * 100.) * np.sqrt(T(Om / (a * a * a) + Ol), dtype=T) * np.sqrt(a, dtype=T)))


class Grid:
    """grids::Grid<T> metadata (grid.hpp:84-130): host only."""

    def __init__(self, simsize, n, dx=None, x0=0.0, y0=0.0, z0=0.0, dtype=np.float64):
        T = np.dtype(dtype).type
        self.dtype = np.dtype(dtype)
        self.periodicDomainSize = T(simsize)
        self.cellSize = T(dx) if dx is not None else T(T(simsize) / T(n))
        self.thisGridSize = T(self.cellSize * T(n))
        self.offsetLower = (T(x0), T(y0), T(z0))
        self.size, self.size2, self.size3 = n, n * n, n * n * n
        self.kMin = T(2. * math.pi / float(self.thisGridSize))  # setKmin, grid.hpp:128-130

        self.simEquivalentSize = int(round(float(simsize) / float(self.cellSize)))  # grid.hpp:111
        self.flags = np.zeros(0, dtype=np.int64)  # sorted flagged cell indices (Grid::flagCells / getFlaggedCells)
        self._mask = None
        self.cellMassFrac = T(math.pow(float(T(self.cellSize / T(simsize))), 3.0))   # grid.hpp:112
        self.cellSofteningScale = T(1.0)
        self.base = self                          # the stored grid a virtual grid is a view of

    # virtual views used by the particle mappers (virtualgrid.hpp): same cells at another resolution / mass
    def makeSupersampled(self, ratio):
        return VirtualGrid(self, "super", int(ratio))

    def makeSubsampled(self, ratio):
        return VirtualGrid(self, "sub", int(ratio))

    def makeScaledMassVersion(self, massScale):
        return VirtualGrid(self, "mass", massScale)

    def getFourierKmin(self):
        return self.kMin

    def desc(self):
        """The const metadata a cross-grid kernel needs (gx_grid_desc)."""
        d = L.GridDesc()
        d.n, d.cell_size, d.domain_size = self.size, float(self.cellSize), float(self.periodicDomainSize)
        for a in range(3):
            d.offset_lower[a] = float(self.offsetLower[a])
        return d

    def flagCells(self, indices, append=False):
        """Sets the flagged cells; append=True is the reference's Grid::flagCells (grid.hpp:269-275: union with the flags
        already there), the default replaces them (unflagAllCells + flagCells)."""
        new = np.asarray(indices, dtype=np.int64).reshape(-1)
        self.flags = np.unique(np.concatenate([self.flags, new]) if append else new)
        self._mask = None

    def unflagAllCells(self):
        self.flagCells([])

    def getFlaggedCells(self):
        return self.flags

    def flagMask(self, complement=False):
        """Device byte mask of the flagged cells (n^3 bytes, 1 = flagged), cached until the flags change."""
        if self._mask is None:
            self._mask = {}
        if complement not in self._mask:
            lib = L.load()
            m = np.full(self.size3, 1 if complement else 0, dtype=np.uint8)
            m[self.flags] = 0 if complement else 1
            p = C.c_void_p()
            lib.gx_malloc(C.byref(p), m.nbytes)
            lib.gx_memcpy_h2d(p, m.ctypes.data_as(C.c_void_p), m.nbytes, None)
            lib.gx_synchronize(None)
            self._mask[complement] = p
        return self._mask[complement]

    def coversFullSimulation(self):
        return self.periodicDomainSize == self.thisGridSize


class VirtualGrid(Grid):
    """grids::SuperSampleGrid / SubSampleGrid / MassScaledGrid (virtualgrid.hpp:125-182, 520-616, 657-676): metadata of a
    view of `underlying`; flags are proxied to the stored grid (scaled up or down), fields are evaluated by tricubic
    interpolation (super) or sub-cell averaging (sub) of the stored grid's field."""

    def __init__(self, underlying, kind, arg):
        T = underlying.dtype.type
        u = underlying
        if kind == "super":
            n, dx = u.size * arg, T(u.cellSize / T(arg))
            frac, soft = T(u.cellMassFrac / T(arg * arg * arg)), T(arg)      # (sic: the factor is passed as softening scale)
        elif kind == "sub":
            if u.size % arg:
                raise L.GxError("subsampling factor does not divide the grid size")
            n, dx = u.size // arg, T(u.cellSize * T(arg))
            frac, soft = T(u.cellMassFrac * T(math.pow(float(arg), 3.0))), u.cellSofteningScale
        elif kind == "mass":
            n, dx = u.size, u.cellSize
            frac, soft = T(T(arg) * u.cellMassFrac), u.cellSofteningScale
        else:
            raise ValueError(kind)
        Grid.__init__(self, u.periodicDomainSize, n, dx=dx, x0=u.offsetLower[0], y0=u.offsetLower[1], z0=u.offsetLower[2],
                      dtype=u.dtype)
        self.underlying, self.kind, self.arg, self.base = u, kind, arg, u.base
        self.cellMassFrac, self.cellSofteningScale = frac, soft

    # resolution of this view relative to the stored grid, as (up, down) integer factors
    def resolution(self):
        up, down = self.underlying.resolution() if isinstance(self.underlying, VirtualGrid) else (1, 1)
        if self.kind == "super":
            up *= self.arg
        elif self.kind == "sub":
            down *= self.arg
        return up, down

    def makeSupersampled(self, ratio):
        if self.kind == "super":
            return self.underlying.makeSupersampled(ratio * self.arg)
        if self.kind == "sub":
            if ratio > self.arg:
                return self.underlying.makeSupersampled(_int_ratio(ratio, self.arg))
            if ratio == self.arg:
                return self.underlying
            return self.underlying.makeSubsampled(_int_ratio(self.arg, ratio))
        return VirtualGrid(self, "super", int(ratio))

    def makeSubsampled(self, ratio):
        if self.kind == "sub":
            return self.underlying.makeSubsampled(ratio * self.arg)
        if self.kind == "super":
            if ratio > self.arg:
                return self.underlying.makeSubsampled(_int_ratio(ratio, self.arg))
            if ratio == self.arg:
                return self.underlying
            return self.underlying.makeSupersampled(_int_ratio(self.arg, ratio))
        return VirtualGrid(self, "sub", int(ratio))

    def _scale(self, cells, src_n, dst_n):
        return scale_cells(cells, src_n, dst_n)

    def getFlaggedCells(self):
        u = self.underlying.getFlaggedCells()
        return u if self.kind == "mass" else self._scale(u, self.underlying.size, self.size)

    def flagCells(self, indices, append=False):
        ids = np.asarray(indices, dtype=np.int64).reshape(-1)
        self.underlying.flagCells(ids if self.kind == "mass" else self._scale(ids, self.size, self.underlying.size), append)

    def unflagAllCells(self):
        self.underlying.unflagAllCells()


class ResolutionMatchingGrid(Grid):
    """grids::ResolutionMatchingGrid (virtualgrid.hpp:183-360), the output grid of a zoom level when stray particles are
    allowed: the whole extent of the low-resolution grid at the cell size of the high-resolution one. Inside the
    high-resolution window a cell is a cell of `hi`; outside it is a cell of `lo` super-sampled (tricubic interpolation of
    the low-resolution fields). Flags live on `hi` and, for the interpolated part, in an independent list."""

    def __init__(self, hi, lo):
        T = hi.dtype.type
        r = float(lo.cellSize) / float(hi.cellSize)
        if abs(r - round(r)) > 1e-8 or round(r) < 1:
            raise L.GxError("ResolutionMatchingGrid: the cell size ratio is not a positive integer")
        self.ratio = int(round(r))
        Grid.__init__(self, lo.periodicDomainSize, lo.size * self.ratio, dx=hi.cellSize, x0=lo.offsetLower[0],
                      y0=lo.offsetLower[1], z0=lo.offsetLower[2], dtype=hi.dtype)
        self.cellMassFrac, self.cellSofteningScale = hi.cellMassFrac, hi.cellSofteningScale
        self.hi, self.lo, self.base = hi, lo, hi.base
        self.loInterp = lo.makeSupersampled(self.ratio)          # withIndependentFlags(): its flags are self.interpFlags
        self.interpFlags = np.zeros(0, dtype=np.int64)
        self.corner = [int(round(float(T(hi.offsetLower[a] - lo.offsetLower[a])) / float(self.cellSize))) for a in range(3)]
        W = self.simEquivalentSize
        self.win_lo = [c % W for c in self.corner]
        self.win_hi = [((c + hi.size - 1) % W) + 1 for c in self.corner]

    def inHiResWindow(self, coords):
        ok = np.ones(len(coords[0]), dtype=bool)
        for a in range(3):
            lo, hi, c = self.win_lo[a], self.win_hi[a], coords[a]
            ok &= ((c < hi) & (c >= lo)) if hi >= lo else ((c < hi) | (c >= lo))
        return ok

    def split(self, cells):
        """cells of this grid -> (mask of those in the high-resolution window, their cell ids on `hi`)."""
        n = self.size
        c = np.unravel_index(np.asarray(cells, dtype=np.int64), (n, n, n))
        m = self.inHiResWindow(c)
        W, h = self.simEquivalentSize, self.hi.size
        ch = [((c[a][m] - self.corner[a]) % W) for a in range(3)]
        return m, (ch[0] * h + ch[1]) * h + ch[2]

    def getFlaggedCells(self):
        h, n, W = self.hi.size, self.size, self.simEquivalentSize
        hf = np.asarray(self.hi.getFlaggedCells(), dtype=np.int64)
        c = np.unravel_index(hf, (h, h, h))
        up = [(c[a] + self.corner[a]) % W for a in range(3)]
        mapped = (up[0] * n + up[1]) * n + up[2]
        m, _ = self.split(self.interpFlags)
        return np.unique(np.concatenate([mapped, self.interpFlags[~m]]))

    def flagCells(self, indices, append=False):
        ids = np.asarray(indices, dtype=np.int64).reshape(-1)
        m, hi_cells = self.split(ids)
        self.hi.flagCells(hi_cells, append)
        self.interpFlags = np.unique(np.concatenate([self.interpFlags, ids[~m]]) if append else ids[~m])

    def unflagAllCells(self):
        self.hi.unflagAllCells()
        self.interpFlags = np.zeros(0, dtype=np.int64)

    def makeSupersampled(self, ratio):
        return ResolutionMatchingGrid(self.hi.makeSupersampled(ratio), self.lo.makeSupersampled(ratio))

    def makeSubsampled(self, ratio):
        return ResolutionMatchingGrid(self.hi.makeSubsampled(ratio), self.lo.makeSubsampled(ratio))


def scale_cells(cells, src_n, dst_n):
    """Grid::upscaleCellFlagVector / downscaleCellFlagVector (grid.hpp:384-436): cell ids of a grid of side src_n as ids
    of the same volume at side dst_n (all sub-cells when refining, the unique parents when coarsening)."""
    cells = np.asarray(cells, dtype=np.int64)

    c = np.stack(np.unravel_index(cells, (src_n,) * 3), axis=0)
    if dst_n >= src_n:
        f = dst_n // src_n
        o = np.arange(f, dtype=np.int64)
        x = c[0][:, None, None, None] * f + o[None, :, None, None]
        y = c[1][:, None, None, None] * f + o[None, None, :, None]
        z = c[2][:, None, None, None] * f + o[None, None, None, :]
        return ((x * dst_n + y) * dst_n + z).reshape(-1)
    f = src_n // dst_n
    return np.unique(((c[0] // f) * dst_n + (c[1] // f)) * dst_n + (c[2] // f))


def _int_ratio(a, b):
    if a % b:
        raise L.GxError("The ratio of sub/supersampling factors is not an integer")
    return a // b


class Field:
    """fields::Field<T> with device-resident storage (padded half-complex layout in HBM)."""

    def __init__(self, grid, data=None, fourier=True, stream=None):
        self.lib = L.load()
        self.grid = grid
        self.n = grid.size
        self.dtype = grid.dtype
        self.dt = _dt(self.dtype)
        self.fourier = bool(fourier)
        self.stream = stream
        self.nbytes = self.lib.gx_field_bytes(self.n, self.dt)
        p = C.c_void_p()
        self.lib.gx_malloc(C.byref(p), self.nbytes)
        self.ptr = p
        if data is None:
            self.lib.gx_memset_zero(self.ptr, self.nbytes, self.stream)
        else:
            self.setDataVector(data)

    def __del__(self):
        try:
            if getattr(self, "ptr", None) is not None and self.ptr.value:
                self.lib.cdll.gx_free(self.ptr)
                self.ptr = None
        except Exception:
            pass

    # ---- host mirror -------------------------------------------------------------------------------
    def setDataVector(self, data):
        """Upload host data in the reference's layout for the field's current space."""
        a = np.ascontiguousarray(data)
        n = self.n
        if self.fourier:
            if np.iscomplexobj(a):
                a = np.ascontiguousarray(a.astype(np.complex64 if self.dt == L.F32 else np.complex128))
            else:
                a = np.ascontiguousarray(a, dtype=self.dtype)
            assert a.nbytes == self.nbytes, "Fourier-space data must have 2*n*n*(n/2+1) reals"
            self.lib.gx_field_upload(self.ptr, a.ctypes.data_as(C.c_void_p), n, self.dt, L.LAYOUT_FOURIER, self.stream)
        else:
            a = np.ascontiguousarray(a, dtype=self.dtype).reshape(-1)
            assert a.size == n ** 3
            self.lib.gx_field_upload(self.ptr, a.ctypes.data_as(C.c_void_p), n, self.dt, L.LAYOUT_REAL, self.stream)
        self.lib.gx_synchronize(self.stream)

    def getDataVector(self):
        """Host copy in the reference's layout: complex [n,n,n/2+1] in Fourier space, real [n,n,n] otherwise."""
        n = self.n
        if self.fourier:
            out = np.empty((n, n, n // 2 + 1), dtype=np.complex64 if self.dt == L.F32 else np.complex128)
            self.lib.gx_field_download(out.ctypes.data_as(C.c_void_p), self.ptr, n, self.dt, L.LAYOUT_FOURIER, self.stream)
        else:
            out = np.empty((n, n, n), dtype=self.dtype)
            self.lib.gx_field_download(out.ctypes.data_as(C.c_void_p), self.ptr, n, self.dt, L.LAYOUT_REAL, self.stream)
        self.lib.gx_synchronize(self.stream)
        return out

    def copy(self):
        f = Field(self.grid, fourier=self.fourier, stream=self.stream)
        self.lib.gx_memcpy_d2d(f.ptr, self.ptr, self.nbytes, self.stream)
        return f

    def getGrid(self):
        return self.grid

    # ---- Fourier support (field.hpp:635-738) ---------------------------------------------------------
    def isFourier(self):
        return self.fourier

    def setFourier(self, fourier):
        self.fourier = bool(fourier)

    def toFourier(self):
        if self.fourier:
            return
        self.lib.gx_fft_forward(self.ptr, self.n, self.dt, self.stream)
        self.fourier = True

    def toReal(self):
        if not self.fourier:
            return
        self.lib.gx_fft_inverse(self.ptr, self.n, self.dt, self.stream)
        self.fourier = False

    def ensureFourierModesAreMirrored(self):
        assert self.fourier
        self.lib.gx_mirror(self.ptr, self.n, self.dt, self.stream)

    def getFourierCoefficient(self, kx, ky, kz):
        re, im = C.c_double(), C.c_double()
        self.lib.gx_synchronize(self.stream)
        self.lib.gx_get_mode(self.ptr, self.n, self.dt, kx, ky, kz, C.byref(re), C.byref(im))
        return complex(re.value, im.value)

    def setFourierCoefficient(self, kx, ky, kz, value):
        self.lib.gx_synchronize(self.stream)
        self.lib.gx_set_mode(self.ptr, self.n, self.dt, kx, ky, kz, float(value.real), float(value.imag))

    # ---- operators ------------------------------------------------------------------------------------
    def applyTransferFunction(self, spectrum_or_covariance, power):
        """Field::applyTransferFunction (field.hpp:394-412). Accepts a PowerSpectrum (covariance evaluated in
        the kernel, nothing materialised) or a covariance Field as the reference does."""
        assert self.fourier
        if isinstance(spectrum_or_covariance, Field):
            self.lib.gx_apply_transfer_field(self.ptr, spectrum_or_covariance.ptr, self.n, self.dt, float(power),
                                             self.stream)
        else:
            self.lib.gx_apply_transfer(self.ptr, self.n, self.dt, float(self.grid.thisGridSize),
                                       spectrum_or_covariance.handle, float(power), self.stream)

    def applyFilter(self, filt):
        """Field::applyFilter (field.hpp:741-746); converts to Fourier space first like forEachFourierCell does."""
        self.toFourier()
        fac = filt.factors()
        arr = (L.FilterFactor * len(fac))(*fac)
        self.lib.gx_apply_filter(self.ptr, self.n, self.dt, float(self.grid.thisGridSize), arr, len(fac), self.stream)

    def enforceUnitVariance(self):
        self.toFourier()
        self.lib.gx_unit_variance(self.ptr, self.n, self.dt, self.stream)

    def __imul__(self, value):
        self.lib.gx_scale(self.ptr, self.n, self.dt, float(value), self.stream)
        return self

    def addScaled(self, other, scale):
        """Field::addScaled over the whole storage (field.hpp:355-362)."""
        self.lib.gx_axpy(self.ptr, other.ptr, self.n, self.dt, float(scale), self.stream)

    def addScaledFourier(self, other, scale):
        """MultiLevelField::addScaled for one level (multilevelfield.hpp:194-212)."""
        self.toFourier()
        self.lib.gx_axpy_fourier(self.ptr, other.ptr, self.n, self.dt, float(scale), self.stream)

    def innerProduct(self, other):
        """Real space: Field::innerProduct (field.hpp:365-377). Fourier space: the weighted mode sum of
        MultiLevelField::innerProduct (multilevelfield.hpp:285-290)."""
        out = C.c_double()
        if self.fourier:
            assert other.fourier
            self.lib.gx_dot_fourier(self.ptr, other.ptr, self.n, self.dt, C.byref(out), self.stream)
        else:
            assert not other.fourier
            self.lib.gx_dot_real(self.ptr, other.ptr, self.n, self.dt, C.byref(out), self.stream)
        return out.value

    def norm(self):
        return math.sqrt(self.innerProduct(self))

    def __iadd__(self, other):
        """operator+= over the whole storage (tools/numerics/vectormath.hpp)."""
        self.lib.gx_axpy(self.ptr, other.ptr, self.n, self.dt, 1.0, self.stream)
        return self

    def __isub__(self, other):
        self.lib.gx_axpy(self.ptr, other.ptr, self.n, self.dt, -1.0, self.stream)
        return self

    def __neg__(self):
        f = self.copy()
        f *= -1.0
        return f

    def mulMask(self, mask_dev):
        """operator*=(Field<char>) (field.hpp:306-318) with the mask as n^3 device bytes."""
        assert not self.fourier
        self.lib.gx_mul_mask(self.ptr, mask_dev, self.n, self.dt, self.stream)

    # ---- windows and other grids (field.hpp:752-849) ---------------------------------------------------
    def setZeroInsideWindow(self, window_grid):
        self.toReal()
        self.lib.gx_zero_window(self.ptr, C.byref(self.grid.desc()), self.dt, C.byref(window_grid.desc()), 1, self.stream)

    def setZeroOutsideWindow(self, window_grid):
        self.toReal()
        self.lib.gx_zero_window(self.ptr, C.byref(self.grid.desc()), self.dt, C.byref(window_grid.desc()), 0, self.stream)

    def applyFilterInWindow(self, filt, window_grid, windowFirst):
        """Field::applyFilterInWindow (field.hpp:752-772); the window is window_grid.getWindow()."""
        inWindow = self.copy()
        if windowFirst:
            inWindow.setZeroOutsideWindow(window_grid)
            inWindow.toFourier()
            inWindow.applyFilter(filt)
        else:
            inWindow.toFourier()
            inWindow.applyFilter(filt)
            inWindow.setZeroOutsideWindow(window_grid)
        self.setZeroInsideWindow(window_grid)
        if inWindow.fourier:
            self.toFourier()
        else:
            self.toReal()
        self += inWindow

    def addFieldFromDifferentGrid(self, source):
        """Field::addFieldFromDifferentGrid (field.hpp:804-812,834-840): tricubic interpolation of a coarser source,
        sub-cell average of a finer one."""
        source.toReal()
        self.toReal()
        self.lib.gx_interp_add(self.ptr, C.byref(self.grid.desc()), source.ptr, C.byref(source.grid.desc()), self.dt,
                               self.stream)

    def addFieldFromDifferentGridWithFilter(self, source, filt):
        """FILTER_ON_COARSE_GRID branch (field.hpp:819-832), the reference's default build."""
        tmp = source.copy()
        tmp.applyFilter(filt)
        tmp.toReal()
        self.toReal()
        self.addFieldFromDifferentGrid(tmp)


# ---- filters -----------------------------------------------------------------------------------------
class _Filter:
    def factors(self):
        raise NotImplementedError

    def __mul__(self, other):
        if isinstance(other, (int, float)):
            return FilterProduct(self.factors() + [L.FilterFactor(L.FILTER_CONST, float(other), 0.0, 1)])
        return FilterProduct(self.factors() + other.factors())


class FilterProduct(_Filter):
    def __init__(self, factors):
        self._f = list(factors)

    def factors(self):
        return list(self._f)


class LowPassFermiFilter(_Filter):
    """filters::LowPassFermiFilter (filter.hpp:209-239): temperature = 0.1 * kcut."""

    def __init__(self, kcut, kind=L.FILTER_FERMI_LP):
        self.kcut, self.kind = float(kcut), kind

    def factors(self):
        return [L.FilterFactor(self.kind, self.kcut, self.kcut * 0.1, 1)]


def ComplementaryCovarianceFilter(kcut):
    return LowPassFermiFilter(kcut, L.FILTER_FERMI_HP_COV)


def ComplementaryFilter(kcut):
    return LowPassFermiFilter(kcut, L.FILTER_FERMI_HP)


# ---- power spectra -------------------------------------------------------------------------------------
def read_camb_table(filename):
    """CAMB::readLinesFromCambOutput (camb.hpp:224-268) with io::getBufferIgnoringColumnHeaders /
    getNumberOfColumns (io/input.hpp:47-95). Returns the raw numeric table [rows, ncol]."""
    with open(filename) as f:
        lines = f.read().split("\n")
    vals = []
    for i, line in enumerate(lines):
        toks = line.split()
        if i == 0:
            try:
                vals.extend(float(t) for t in toks)
            except ValueError:
                pass  # column headers on the first line are skipped
            continue
        try:
            vals.extend(float(t) for t in toks)
        except ValueError:
            raise L.GxError("Error reading file " + filename)
    ncol = len(lines[1].split())
    if ncol not in (7, 13):
        raise L.GxError("CAMB transfer file doesn't have a sensible number of columns")
    a = np.array(vals, dtype=np.float64)
    return a[: (a.size // ncol) * ncol].reshape(-1, ncol)


class _PowerSpectrum:
    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.cdll.gx_spectrum_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def __call__(self, k):
        return self.lib.gx_spectrum_eval(self.handle, float(k))

    def getPowerSpectrumForGrid(self, grid):
        """PowerSpectrum::getPowerSpectrumForGrid (camb.hpp:75-119): the covariance as a Fourier field."""
        f = Field(grid, fourier=True)
        self.lib.gx_covariance_field(f.ptr, grid.size, _dt(grid.dtype), float(grid.thisGridSize), self.handle, None)
        return f


class CAMB(_PowerSpectrum):
    """cosmology::CAMB<T> (camb.hpp:173-332). species: 'dm' (col 1), 'baryon' (col 2), 'all' (col 6)."""
    COLUMN = {"dm": 1, "baryon": 2, "all": 6}

    def __init__(self, cosmology, filename, species="all", dtype=np.float64):
        self.lib = L.load()
        tab = read_camb_table(filename)
        norm = tab.reshape(-1)[1]  # first CDM entry (camb.hpp:254)
        tab = tab[tab[:, 0] > 0]
        if np.dtype(dtype) == np.float32:
            f = np.float32
            k = tab[:, 0].astype(f).astype(np.float64)
            col = lambda j: (tab[:, j].astype(f) / f(norm)).astype(np.float64)
        else:
            k = tab[:, 0].copy()
            col = lambda j: tab[:, j] / norm
        T = np.ascontiguousarray(col(self.COLUMN[species]))
        Tn = np.ascontiguousarray(col(6))
        k = np.ascontiguousarray(k)
        h = C.c_void_p()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        a_scale = cosmology.scalefactor
        if np.dtype(dtype) == np.float32:
            a_scale = float(np.float32(1.) / np.float32(np.float32(cosmology.redshift) + np.float32(1.)))
        self.lib.gx_spectrum_create_camb(p(k), p(T), p(Tn), len(k), cosmology.OmegaM0, cosmology.OmegaLambda0,
                                         cosmology.sigma8, cosmology.ns, a_scale, _dt(dtype), C.byref(h))
        self.handle = h
        self.amplitude = self.lib.gx_spectrum_amplitude(h)


class PowerLawPowerSpectrum(_PowerSpectrum):
    """cosmology::PowerLawPowerSpectrum (camb.hpp:150-168)."""

    def __init__(self, cosmology, amplitude, dtype=np.float64):
        self.lib = L.load()
        h = C.c_void_p()
        self.lib.gx_spectrum_create_powerlaw(float(amplitude), float(cosmology.ns), _dt(dtype), C.byref(h))
        self.handle = h
        self.amplitude = amplitude


# ---- white noise ---------------------------------------------------------------------------------------
class RandomFieldGenerator:
    """fields::RandomFieldGenerator (randomfieldgenerator.hpp:25-263) for a list of per-level Fields."""

    def __init__(self, fields, seed=0):
        self.fields = fields if isinstance(fields, (list, tuple)) else [fields]
        self.baseSeed = seed
        self.drawInFourierSpace = False
        self.reverseRandomDrawOrder = False
        self.parallel = False
        self.seeded = False

    def setDrawInFourierSpace(self, v):
        self.drawInFourierSpace = bool(v)

    def setParallel(self, v):
        self.parallel = bool(v)

    def setReverseRandomDrawOrder(self, v):
        self.reverseRandomDrawOrder = bool(v)

    def seed(self, seed):
        if self.seeded:
            raise L.GxError("The random number generator has already been seeded")
        self.baseSeed = int(seed)
        self.seeded = True

    def draw(self):
        if not self.seeded:
            raise L.GxError("The random number generator has not been seeded")
        offset = C.c_ulonglong(0)  # one gsl stream across all levels in the serial modes (:125-138)
        for f in self.fields:
            if self.drawInFourierSpace:
                mode = L.NOISE_FOURIER_PARALLEL if self.parallel else L.NOISE_FOURIER_SERIAL
            else:
                mode = L.NOISE_REAL
            f.lib.gx_noise_draw(f.ptr, f.n, f.dt, self.baseSeed, mode, int(self.reverseRandomDrawOrder),
                                C.byref(offset), f.stream)
            f.fourier = True


# ---- Zeldovich -------------------------------------------------------------------------------------------
class ZeldovichParticleGenerator:
    """particle::ZeldovichParticleGenerator (zeldovich.hpp:132-304) on one level."""

    def __init__(self, linearOverdensityField, fused=True):
        self.delta = linearOverdensityField
        self.grid = linearOverdensityField.grid
        self.fused = fused
        self.pOff = None
        self.recalculate()

    def recalculate(self):
        d = self.delta
        d.toFourier()
        out = []
        for axis in range(3):
            psi = Field(self.grid, fourier=True, stream=d.stream)
            if self.fused:
                d.lib.gx_zeldovich_offset(d.ptr, psi.ptr, d.n, d.dt, float(self.grid.thisGridSize), axis, d.stream)
            else:
                d.lib.gx_zeldovich_fourier(d.ptr, psi.ptr, d.n, d.dt, float(self.grid.thisGridSize), axis, d.stream)
                d.lib.gx_fft_inverse(psi.ptr, d.n, d.dt, d.stream)
            psi.fourier = False
            out.append(psi)
        self.pOff = out

    def getGeneratedFields(self):
        return self.pOff


def emit_particles(psis, grid, cosmology, out_dtype=np.float32):
    """ZeldovichParticleEvaluator::getParticle for every cell of a unigrid, in cell-index (= particle id) order.
    Returns host arrays pos[n^3,3], vel[n^3,3] of out_dtype."""
    lib = L.load()
    n = grid.size
    n3 = n ** 3
    odt = _dt(out_dtype)
    esz = np.dtype(out_dtype).itemsize
    dpos, dvel = C.c_void_p(), C.c_void_p()
    lib.gx_malloc(C.byref(dpos), n3 * 3 * esz)
    lib.gx_malloc(C.byref(dvel), n3 * 3 * esz)
    ratio = cosmology.zeldovichVelocityToOffsetRatio(grid.dtype)
    try:
        for axis in range(3):
            lib.gx_emit_component(psis[axis].ptr, n, psis[axis].dt, axis, float(grid.cellSize),
                                  float(grid.offsetLower[axis]), float(grid.periodicDomainSize), ratio,
                                  C.c_void_p(dpos.value + axis * esz), C.c_void_p(dvel.value + axis * esz), odt, 3, 0,
                                  n3, None)
        pos = np.empty((n3, 3), dtype=out_dtype)
        vel = np.empty((n3, 3), dtype=out_dtype)
        lib.gx_memcpy_d2h(pos.ctypes.data_as(C.c_void_p), dpos, pos.nbytes, None)
        lib.gx_memcpy_d2h(vel.ctypes.data_as(C.c_void_p), dvel, vel.nbytes, None)
        lib.gx_synchronize(None)
    finally:
        lib.gx_free(dpos)
        lib.gx_free(dvel)
    return pos, vel


def generate_unigrid_ic(n, box, seed, cosmology, camb_file, mode="fourier_parallel", dtype=np.float64,
                        out_dtype=np.float32, fix_power=False):
    """ICGenerator::done on a unigrid (ic.hpp:1633-1648 -> 1275-1331): draw, [fix_power], apply P(k)^1/2,
    Zeldovich, particle evaluation. Returns (pos, vel, mass, eps) with pos/vel as host arrays."""
    grid = Grid(box, n, dtype=dtype)
    field = Field(grid, fourier=True)
    rfg = RandomFieldGenerator(field)
    rfg.seed(seed)
    rfg.setDrawInFourierSpace(mode != "real")
    rfg.setParallel(mode == "fourier_parallel")
    rfg.setReverseRandomDrawOrder(mode == "fourier_reverse")
    rfg.draw()
    if fix_power:
        field.enforceUnitVariance()
    spectrum = CAMB(cosmology, camb_file, "all", dtype)
    field.applyTransferFunction(spectrum, 0.5)
    gen = ZeldovichParticleGenerator(field)
    pos, vel = emit_particles(gen.getGeneratedFields(), grid, cosmology, out_dtype)
    T = np.dtype(dtype).type
    boxMass = T(27.744948) * T(cosmology.OmegaM0) * T(float(grid.periodicDomainSize) ** 3.0)  # zeldovich.hpp:61
    mass = T(boxMass) * T(float(grid.cellSize / grid.periodicDomainSize) ** 3.0)
    eps = grid.cellSize * T(1.0) * T(0.01075)
    return pos, vel, mass, eps
