"""TEST INFRASTRUCTURE -- numpy restatement of genetIC's multi-level (zoom) field machinery, modifications and
splicing (not product code; see oracle/model.py for the single-level pieces it builds on).

Paths are relative to /root/reference/genetIC/src. Compile-time options follow the reference's default Makefile
(genetIC/Makefile:23): CUBIC_INTERPOLATION, FILTER_ON_COARSE_GRID, ZELDOVICH_GRADIENT_FOURIER_SPACE on;
VELOCITY_MODIFICATION_GRADIENT_FOURIER_SPACE off.

Fields are complex [n,n,n/2+1] arrays in Fourier space or real [n,n,n] arrays, as in oracle/model.py. Operations
that the reference performs only on the independent Fourier slots (forEachFourierCell*, fourier.hpp:110-151) leave
the redundant slots of the kz = 0 / N/2 planes untouched here as well; every transform to real space mirrors first.
"""
import math

import numpy as np

from . import model as M


# --------------------------------------------------------------------------------------------------------------
# grids (simulation/grid/grid.hpp:84-130) and integer windows (simulation/window.hpp)
# --------------------------------------------------------------------------------------------------------------
class Grid:
    def __init__(self, simsize, n, cell=None, offset=(0.0, 0.0, 0.0)):
        self.simsize = float(simsize)
        self.n = int(n)
        self.cell = float(cell) if cell is not None else float(simsize) / n
        self.size = self.cell * n                      # thisGridSize
        self.offset = tuple(float(o) for o in offset)  # offsetLower
        self.kmin = 2.0 * math.pi / self.size          # setKmin, grid.hpp:128-130
        self.sim_equiv = int(round(self.simsize / self.cell))  # simEquivalentSize, grid.hpp:111
        self.flags = np.zeros(0, dtype=np.int64)       # sorted flagged cell indices

    def covers_full(self):
        return self.simsize == self.size

    def centroids_1d(self, axis):
        """getCentroidFromCoordinate (grid.hpp:662-669): coord*cell + offsetLower + cell/2, in that order."""
        return (np.arange(self.n, dtype=np.float64) * self.cell + self.offset[axis]) + self.cell / 2

    def wrapped_offset(self, a, b):
        """Grid::getWrappedOffset (b - a wrapped into (-L/2, L/2])."""
        d = np.asarray(b, dtype=np.float64) - np.asarray(a, dtype=np.float64)
        L = self.simsize
        d = np.where(d < -L / 2, d + L, d)
        d = np.where(d > L / 2, d - L, d)
        return d

    def get_wrapped_offset(self, x0, x1):
        """Grid::getWrappedOffset(x0, x1) = x0 - x1 wrapped (grid.hpp:445-454)."""
        d = np.asarray(x0, dtype=np.float64) - np.asarray(x1, dtype=np.float64)
        L = self.simsize
        d = np.where(d > L / 2, d - L, d)
        return np.where(d < -L / 2, d + L, d)

    def index_from_point(self, p):
        """getIndexFromPoint (grid.hpp:672-677)."""
        c = []
        for a in range(3):
            x = math.fmod(p[a] - self.offset[a], self.simsize)
            if x < 0:
                x += self.simsize
            c.append(int(math.floor(x / self.cell)))
        return (c[0] * self.n + c[1]) * self.n + c[2]


def select(grids, base, centre, inclusion):
    """ICGenerator::select (ic.hpp:1458-1484): flags, on every level, the cells whose centroid offset from
    `centre` (wrapped with the base grid's period) satisfies inclusion(dx,dy,dz)."""
    for g in grids:
        d = [base.wrapped_offset(g.centroids_1d(a), centre[a]) for a in range(3)]
        m = inclusion(d[0][:, None, None], d[1][None, :, None], d[2][None, None, :])
        g.flags = np.flatnonzero(np.broadcast_to(m, (g.n, g.n, g.n)).reshape(-1)).astype(np.int64)


def select_sphere(grids, centre, radius):
    r2 = np.float32(radius) * np.float32(radius)  # selectSphere(float radius): r2 formed in float, ic.hpp:1490-1496
    select(grids, grids[0], centre, lambda x, y, z: x * x + y * y + z * z < float(r2))


def select_cube(grids, centre, side):
    h = float(np.float32(side) / np.float32(2))   # selectCube(float side), ic.hpp:1519-1524
    select(grids, grids[0], centre, lambda x, y, z: (np.abs(x) < h) & (np.abs(y) < h) & (np.abs(z) < h))


def select_ellipse(grids, centre, a, b, c):
    a2, b2, c2 = (float(np.float32(v) * np.float32(v)) for v in (a, b, c))   # selectEllipse(float a, b, c), ic.hpp:1503-1513
    select(grids, grids[0], centre, lambda x, y, z: x * x / a2 + y * y / b2 + z * z / c2 < 1.)


def select_nearest(grids, centre):
    """ICGenerator::selectNearest (ic.hpp:1447-1455): one cell on the deepest level."""
    for g in grids:
        g.flags = np.zeros(0, dtype=np.int64)
    g = grids[-1]
    g.flags = np.array([g.index_from_point(centre)], dtype=np.int64)


class IntWindow:
    """Window<int> (window.hpp): periodic cuboid [lci, uce) per axis."""

    def __init__(self, wrap, point):
        self.wrap = wrap
        self.lci = list(point)
        self.uce = [p + 1 for p in point]

    def _w(self, v):
        if v < 0:
            v += self.wrap
        if v >= self.wrap:
            v -= self.wrap
        return v

    @staticmethod
    def _within(lo, hi, v):
        return (lo <= v < hi) if hi >= lo else (v < hi or v >= lo)

    def contains_axis(self, v, a):
        return self._within(self.lci[a], self.uce[a], v)

    def expand_to_include(self, p):
        for a in range(3):
            lo, hi, v = self.lci[a], self.uce[a], p[a]
            if not self._within(lo, hi, v):
                size_lower = self._w(hi - v - 1) + 1
                size_upper = self._w(v + 1 - lo - 1) + 1
                can_lower = self._within(v, hi, lo)
                can_upper = self._within(lo, v + 1, hi - 1)
                if can_lower and (size_lower < size_upper or not can_upper):
                    self.lci[a] = v
                else:
                    self.uce[a] = v + 1

    def sizes(self):
        return [self._w(self.uce[a] - self.lci[a] - 1) + 1 for a in range(3)]

    def wrap_corners(self):
        self.lci = [self._w(v) for v in self.lci]
        self.uce = [self._w(v - 1) + 1 for v in self.uce]

    def shift(self, d):
        self.lci = [self.lci[a] + d[a] for a in range(3)]
        self.uce = [self.uce[a] + d[a] for a in range(3)]
        self.wrap_corners()

    def centre(self):
        return [self._w(self.lci[a] + self._w(self.uce[a] - 1 - self.lci[a]) // 2) for a in range(3)]

    def expand_symmetrically(self, new):
        """expandSymmetricallyToSize (window.hpp), including its comparison of the y corner against oldLCI.z."""
        old = list(self.lci)
        if new == self.wrap:
            self.lci, self.uce = [0, 0, 0], [new, new, new]
            return
        c = self.centre()
        self.lci = [self._w(c[a] - new // 2) for a in range(3)]
        if new % 2 == 0:
            self.uce = [self._w(c[a] + new // 2 - 1) + 1 for a in range(3)]
        else:
            self.uce = [self._w(c[a] + new // 2) + 1 for a in range(3)]
        if self.lci[0] > old[0]:
            self.lci[0] -= 1
            self.uce[0] -= 1
        if self.lci[1] > old[2]:
            self.lci[1] -= 1
            self.uce[1] -= 1
        if self.lci[2] > old[2]:
            self.lci[2] -= 1
            self.uce[2] -= 1

    def wrapped_offset(self, a, b):
        d = b - a
        # integer arithmetic as in Window<int>::getWrappedOffset (-wrapLength / 2 truncates towards zero)
        if d < -(self.wrap // 2):
            d += self.wrap
        if d > self.wrap // 2:
            d -= self.wrap
        return d


def zoom_corner(parent, zoomfac):
    """ICGenerator::initZoomGrid (ic.hpp:398-479) for a parent whose output grid is the grid itself (no stray
    particles): lower-left corner, in parent cells, of the zoom window that holds parent.flags."""
    n = parent.n
    coords = np.stack(np.unravel_index(parent.flags, (n, n, n)), axis=1)
    wrap = parent.sim_equiv
    w = IntWindow(wrap, [int(v) for v in coords[0]])
    for c in coords:
        w.expand_to_include([int(v) for v in c])
    n_user = n // zoomfac
    n_req = max(w.sizes())
    if n_req > n_user:
        raise RuntimeError("Zoom particles do not fit in specified sub-box. Decrease zoom, or choose different particles")
    if n_req < n_user:
        w.expand_symmetrically(n_user)
    # window of the level above in cells of that level
    lo = [int(round(parent.offset[a] / parent.cell)) for a in range(3)]
    w.shift(lo)
    # moveInto(actualWindowAbove)
    hi = [lo[a] + n for a in range(3)]
    above = IntWindow(wrap, lo)
    above.uce = hi
    above.wrap_corners()
    move = [0, 0, 0]
    for a in range(3):
        if not above.contains_axis(w.lci[a], a):
            move[a] = w.wrapped_offset(w.lci[a], above.lci[a])
        if not above.contains_axis(w.uce[a] - 1, a):
            move[a] = w.wrapped_offset(w.uce[a], above.uce[a])
    w.shift(move)
    w.shift([-v for v in above.lci])
    return tuple(w.lci)


def add_zoom_level(grids, zoomfac, n, corner=None):
    """initZoomGridWithLowLeftCornerAt (ic.hpp:502-598): appends the zoom grid and transfers the flags."""
    parent = grids[-1]
    if corner is None:
        corner = zoom_corner(parent, zoomfac)
    off = tuple(parent.offset[a] + corner[a] * parent.cell for a in range(3))
    g = Grid(parent.simsize, n, parent.size / zoomfac / n, off)
    # flags of the parent seen through SuperSample + Section (grid.hpp:195-228, virtualgrid.hpp:161-166,489-498)
    ratio = int(round(parent.cell / g.cell))
    pm = np.zeros(parent.n ** 3, dtype=bool)
    pm[parent.flags] = True
    pm = pm.reshape(parent.n, parent.n, parent.n)
    idx = [((np.arange(n) // ratio) + corner[a]) % parent.sim_equiv for a in range(3)]
    inside = [(i < parent.n) for i in idx]
    sub = pm[np.ix_(idx[0] % parent.n, idx[1] % parent.n, idx[2] % parent.n)]
    sub &= inside[0][:, None, None] & inside[1][None, :, None] & inside[2][None, None, :]
    g.flags = np.flatnonzero(sub.reshape(-1)).astype(np.int64)
    grids.append(g)
    return g


def level_weights(grids):
    """MultiLevelGrid::addLevel (multilevelgrid.hpp:112-118)."""
    return [1.0 if i == 0 else (g.cell / grids[0].cell) ** 3.0 for i, g in enumerate(grids)]


def num_dof(grids):
    """getNumDof (multilevelgrid.hpp:160-172)."""
    dof = 0
    for i, g in enumerate(grids):
        dof += g.n ** 3
        if i > 0:
            f = int(round(grids[i - 1].cell / g.cell)) ** 3
            dof -= g.n ** 3 // f
    return dof


# --------------------------------------------------------------------------------------------------------------
# filters (simulation/filters/filter.hpp:209-317, filterfamily.hpp:86-143)
# --------------------------------------------------------------------------------------------------------------
# a filter is a list of factors (kind, kcut, power); kinds as in include/genetic_b200.h
CONST, FERMI_LP, FERMI_HP_COV, FERMI_HP = 0, 1, 2, 3
FRACTIONAL_K_SPLIT = 0.3
FRACTIONAL_FILTER_TEMPERATURE = 0.1


def fermi(k, kcut):
    with np.errstate(over="ignore"):  # exp overflows to inf far above the cut, giving 0 as in the reference
        return 1.0 / (1.0 + np.exp((k - kcut) / (kcut * FRACTIONAL_FILTER_TEMPERATURE)))


def eval_filter(factors, k):
    out = np.ones_like(k)
    for kind, kcut, power in factors:
        if kind == CONST:
            f = np.full_like(k, kcut)
        elif kind == FERMI_LP:
            f = fermi(k, kcut)
        elif kind == FERMI_HP_COV:
            lp = fermi(k, kcut)
            f = np.sqrt(1.0 - lp * lp)
        else:
            f = 1.0 - fermi(k, kcut)
        for _ in range(power):
            out = out * f
    return out


class FilterFamily:
    """filters::FilterFamily (filterfamily.hpp:86-143) before the levels are combined."""

    def __init__(self, grids):
        self.filters, self.lp, self.hp = [[]], [[]], [[]]
        for g in grids[:-1]:
            kcut = FRACTIONAL_K_SPLIT * (g.n * g.kmin)
            filt = self.filters.pop()
            self.lp.pop()
            self.filters.append(filt + [(FERMI_LP, kcut, 1)])
            self.lp.append([(FERMI_LP, kcut, 1)])
            self.filters.append(filt + [(FERMI_HP_COV, kcut, 1)])
            self.hp.append([(FERMI_HP_COV, kcut, 1)])
            self.lp.append([])


def k_magnitude(grid):
    """|k| per storage slot as Field::applyFilter sees it (field.hpp:741-746): sqrt(double(kx^2+ky^2+kz^2)) with
    k = integer wavenumber * kMin. The conjugate slot has the same magnitude."""
    n = grid.n
    ks = M.signed_modes(n).astype(np.float64) * grid.kmin
    kz = np.arange(n // 2 + 1, dtype=np.float64) * grid.kmin
    return np.sqrt(ks[:, None, None] ** 2 + ks[None, :, None] ** 2 + kz[None, None, :] ** 2)


def apply_filter(F, grid, factors):
    """Field::applyFilter: independent slots only."""
    if not factors:
        return F.copy()
    ind = M.independent_mask(grid.n)
    f = eval_filter(factors, k_magnitude(grid))
    return np.where(ind, F * f, F)


def mul(f1, f2):
    return list(f1) + list(f2)


# --------------------------------------------------------------------------------------------------------------
# moving fields between grids (field.hpp:483-508,598-617,804-849; evaluator.hpp; tricubic.hpp)
# --------------------------------------------------------------------------------------------------------------
def catmull_rom_weights(t):
    """Per-axis weights of LocalUnitTricubicApproximation (tricubic.hpp:54-77; the 64 coefficients of :80-430 are
    the tensor product of these; :450-549 lists the same products for the transpose)."""
    t = np.asarray(t, dtype=np.float64)
    return np.stack([-(t - 1) ** 2 * t / 2, (2 - 5 * t ** 2 + 3 * t ** 3) / 2, -t * (-1 - 4 * t + 3 * t ** 2) / 2,
                     (t - 1) * t ** 2 / 2], axis=-1)


def _stencil_1d(dst, src, axis):
    """For each dst cell along `axis`: the four src indices and weights used by evaluateInterpolated at the dst
    centroid (field.hpp:483-508: subtract offsetLower, wrapPoint, floor(x/cell - 0.5); makeTricubicInterpolator
    :598-617 wraps the stencil if the source spans the box, clamps otherwise)."""
    loc = dst.centroids_1d(axis) - src.offset[axis]
    loc = np.fmod(loc, src.simsize)
    loc = np.where(loc < 0, loc + src.simsize, loc)
    u = loc / src.cell - 0.5
    p0 = np.floor(u).astype(np.int64)
    t = u - p0
    idx = p0[:, None] + np.arange(-1, 3)[None, :]
    if src.n == src.sim_equiv:
        idx = np.where(idx > src.sim_equiv - 1, idx - src.sim_equiv, idx)
        idx = np.where(idx < 0, idx + src.sim_equiv, idx)
    else:
        idx = np.clip(idx, 0, src.n - 1)
    return idx, catmull_rom_weights(t)


def _contained_1d(dst, src, axis, ratio_fine):
    """Section/SuperSample containment (virtualgrid.hpp:199-206,472-480): dst cell -> cell of the supersampled source,
    wrapped with the simulation period, must fall inside the source grid."""
    rel = int(round((dst.offset[axis] - src.offset[axis]) / dst.cell))
    c = np.arange(dst.n) + rel
    S = src.sim_equiv * ratio_fine
    c = np.where(c > S - 1, c - S, c)
    c = np.where(c < 0, c + S, c)
    return (c // ratio_fine < src.n) & (c >= 0)


def interpolate_from(dst, src_grid, src_real):
    """Values a coarser source contributes to every cell of dst (SuperSampleEvaluator[SectionEvaluator], 0 where
    the cell is not contained). Separable evaluation of the tricubic."""
    ratio = int(round(src_grid.cell / dst.cell))
    st = [_stencil_1d(dst, src_grid, a) for a in range(3)]
    v = src_real
    # contract axis by axis: out[a,b,c] = sum_ijk wx[a,i] wy[b,j] wz[c,k] v[ix[a,i], iy[b,j], iz[c,k]]
    v = np.einsum("ai,aijk->ajk", st[0][1], v[st[0][0]])               # [A, ny, nz]
    v = np.einsum("bj,abjk->abk", st[1][1], v[:, st[1][0]])             # [A, B, nz]
    v = np.einsum("ck,abck->abc", st[2][1], v[:, :, st[2][0]])          # [A, B, C]
    cont = [_contained_1d(dst, src_grid, a, ratio) for a in range(3)]
    mask = cont[0][:, None, None] & cont[1][None, :, None] & cont[2][None, None, :]
    return np.where(mask, v, 0.0)


def subsample_from(dst, src_grid, src_real):
    """Values a finer source contributes to dst (SubSampleEvaluator[SectionEvaluator], evaluator.hpp:168-205;
    SubSampleGrid::forEachSubcell / containsCellWithCoordinate, virtualgrid.hpp:570-640): the mean over those of
    the ratio^3 subcells of a coarse cell that the fine grid holds (a zoom grid of n cells at zoom factor z covers
    n/z coarse cells, not necessarily an integer), for coarse cells whose first or last subcell it holds; 0
    elsewhere."""
    r = int(round(dst.cell / src_grid.cell))
    S = dst.sim_equiv * r
    A, lowin, highin = [], [], []
    for a in range(3):
        rel = int(round((dst.offset[a] - src_grid.offset[a]) / src_grid.cell))  # Section cellOffset (fine cells)
        f = np.arange(dst.n * r) + rel
        f = np.where(f > S - 1, f - S, f)
        f = np.where(f < 0, f + S, f)
        ok = f < src_grid.n
        m = np.zeros((dst.n, src_grid.n))
        c = np.arange(dst.n * r) // r
        m[c[ok], f[ok]] = 1.0
        A.append(m)
        okr = ok.reshape(dst.n, r)
        lowin.append(okr[:, 0])
        highin.append(okr[:, -1])
    s = np.einsum("ai,bj,ck,ijk->abc", A[0], A[1], A[2], src_real, optimize=True)
    cnt = A[0].sum(1)[:, None, None] * A[1].sum(1)[None, :, None] * A[2].sum(1)[None, None, :]
    cont = (lowin[0][:, None, None] & lowin[1][None, :, None] & lowin[2][None, None, :]) | \
           (highin[0][:, None, None] & highin[1][None, :, None] & highin[2][None, None, :])
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(cont & (cnt > 0), s / cnt, 0.0)


def add_from_different_grid(dst_grid, dst_real, src_grid, src_real):
    """Field::addFieldFromDifferentGrid (field.hpp:804-812)."""
    if src_grid.cell > dst_grid.cell:
        return dst_real + interpolate_from(dst_grid, src_grid, src_real)
    if src_grid.cell < dst_grid.cell:
        return dst_real + subsample_from(dst_grid, src_grid, src_real)
    raise NotImplementedError("same-resolution transfer between different grids")


def add_from_different_grid_with_filter(dst_grid, dst, dst_fourier, src_grid, src_F, factors):
    """addFieldFromDifferentGridWithFilter, FILTER_ON_COARSE_GRID branch (field.hpp:819-832): filter a copy of the
    source on its own grid, both to real space, interpolate/average. src_F is Fourier space. Returns dst in real
    space."""
    tmp = M.to_real(apply_filter(src_F, src_grid, factors))
    d = M.to_real(dst) if dst_fourier else dst
    return add_from_different_grid(dst_grid, d, src_grid, tmp)


def deinterpolate_into(lo_grid, hi_grid, hi_real, scale):
    """The loop of generateMultilevelCovectorFromHiresCovector (multilevelgrid.hpp:283-297) with Field::deInterpolate
    (field.hpp:420-480): every fine cell adds value*scale times the tricubic-transpose weights to 4^3 parent cells.
    (The reference wraps the stencil with `> size`, leaving index == size unwrapped; not reproduced -- zoom regions
    that touch the upper periodic boundary are outside the domain this restates.)"""
    st = [_stencil_1d_wrap(hi_grid, lo_grid, a) for a in range(3)]
    v = hi_real * scale
    n = lo_grid.n
    # transpose of the separable interpolation: scatter-add axis by axis
    out = np.zeros((n, hi_grid.n, hi_grid.n))
    for i in range(4):
        np.add.at(out, st[0][0][:, i], v * st[0][1][:, i][:, None, None])
    v = out
    out = np.zeros((n, n, hi_grid.n))
    for j in range(4):
        np.add.at(out, (slice(None), st[1][0][:, j]), v * st[1][1][:, j][None, :, None])
    v = out
    out = np.zeros((n, n, n))
    for k in range(4):
        np.add.at(out, (slice(None), slice(None), st[2][0][:, k]), v * st[2][1][:, k][None, None, :])
    return out


def _stencil_1d_wrap(dst, src, axis):
    """deInterpolate's own index logic (field.hpp:452-462): always wraps with the grid size."""
    loc = dst.centroids_1d(axis) - src.offset[axis]
    loc = np.fmod(loc, src.simsize)
    loc = np.where(loc < 0, loc + src.simsize, loc)
    u = loc / src.cell - 0.5
    p0 = np.floor(u).astype(np.int64)
    t = u - p0
    idx = p0[:, None] + np.arange(-1, 3)[None, :]
    idx = np.where(idx < 0, idx + src.n, idx)
    idx = np.where(idx >= src.n, idx - src.n, idx)
    return idx, catmull_rom_weights(t)


# --------------------------------------------------------------------------------------------------------------
# windows of a grid in physical coordinates (Grid::getWindow grid.hpp:502-504; Field::setZero*Window :774-795)
# --------------------------------------------------------------------------------------------------------------
def window_mask(grid, win_grid):
    """True where the centroid of a `grid` cell lies in win_grid.getWindow() (Window<T>::contains)."""
    L = grid.simsize
    m = []
    for a in range(3):
        lo = win_grid.offset[a]
        hi = win_grid.offset[a] + win_grid.size
        # Window ctor wrapCorners (window.hpp): lower wrapped, upper = nextafter(wrap(prev(upper)))
        lo = lo + L if lo < 0 else lo
        lo = lo - L if lo >= L else lo
        h = np.nextafter(hi, -np.inf)
        h = h + L if h < 0 else h
        h = h - L if h >= L else h
        hi = np.nextafter(h, np.inf)
        c = grid.centroids_1d(a)
        m.append((c < hi) & (c >= lo) if hi >= lo else (c < hi) | (c >= lo))
    return m[0][:, None, None] & m[1][None, :, None] & m[2][None, None, :]


def apply_filter_in_window(F, grid, factors, wmask, window_first):
    """Field::applyFilterInWindow (field.hpp:752-772). F in Fourier space; returns Fourier space if window_first
    else real space -- mirrored by the second return value (is_fourier)."""
    real = M.to_real(F)
    if window_first:
        inw = np.where(wmask, real, 0.0)
        inw = apply_filter(M.to_fourier(inw), grid, factors)
        this = M.to_fourier(np.where(wmask, 0.0, real))
        return this + inw, True
    inw = M.to_real(apply_filter(F, grid, factors))
    inw = np.where(wmask, inw, 0.0)
    this = np.where(wmask, 0.0, real)
    return this + inw, False


# --------------------------------------------------------------------------------------------------------------
# multi-level fields (simulation/field/multilevelfield.hpp)
# --------------------------------------------------------------------------------------------------------------
def apply_metric(grids, fields, to_covector):
    """MultiLevelField::applyMetric (multilevelfield.hpp:346-391). fields: Fourier-space arrays per level."""
    fam = FilterFamily(grids)
    w = level_weights(grids)
    nl = len(grids)
    out = []
    for level in range(nl):
        f = fam.filters[level]
        if to_covector and level < nl - 1:
            wm = window_mask(grids[level], grids[level + 1])
            r, is_f = apply_filter_in_window(fields[level], grids[level], f, wm, True)
            r, is_f = apply_filter_in_window(r, grids[level], f, wm, False)
        else:
            r, is_f = apply_filter(fields[level], grids[level], mul(f, f)), True
        for s in range(nl):
            if s == level:
                continue
            ratio = w[level] / w[s]
            flt = mul(mul(fam.filters[level], fam.filters[s]), [(CONST, math.sqrt(ratio), 1)])
            r = add_from_different_grid_with_filter(grids[level], r, is_f, grids[s], fields[s], flt)
            is_f = False
        out.append(r if is_f else M.to_fourier(r))
    return out


def inner_product(grids, a, b):
    """MultiLevelField::innerProduct (multilevelfield.hpp:248-294): per level sum over independent modes of
    w Re(conj(a) b) (weights fourier.hpp:116-147)."""
    tot = 0.0
    for g, x, y in zip(grids, a, b):
        tot += float(np.sum(M.mode_weights(g.n) * (np.conj(x) * y).real))
    return tot


def chi2(grids, fields):
    """MultiLevelField::getChi2 (multilevelfield.hpp:461-487)."""
    return inner_product(grids, apply_metric(grids, fields, True), fields)


def add_scaled(a, b, s):
    """MultiLevelField::addScaled (:194-212): independent slots."""
    out = []
    for x, y in zip(a, b):
        ind = M.independent_mask(x.shape[0])
        out.append(np.where(ind, x + s * y, x))
    return out


def draw_levels(grids, seed, mode):
    """RandomFieldGenerator::draw (randomfieldgenerator.hpp:122-139): one stream across the levels in the serial
    modes; the parallel Fourier mode reseeds per shell so every level of equal n gets the same noise."""
    from . import crng
    st = crng.Stream(seed)
    return [M.draw_white_noise(g.n, seed, mode, stream=None if mode == "fourier_parallel" else st) for g in grids]


def combine_levels(grids, delta):
    """initialiseParticleGeneratorBasedOnTemplate (multilevelgenerator.hpp:79-128) for the overdensity and the
    three Zeldovich offset fields. delta: per-level Fourier-space overdensity (transfer applied). Returns per
    level (delta_real, [psi_x, psi_y, psi_z] real)."""
    fam = FilterFamily(grids)
    nl = len(grids)
    d = [x.copy() for x in delta]
    psi = [M.zeldovich_offsets_fourier(d[l], grids[l].size) for l in range(nl)]
    d_is_f = [True] * nl
    p_is_f = [True] * nl
    for l in range(1, nl):
        d[l] = apply_filter(d[l], grids[l], fam.hp[l])
        psi[l] = [apply_filter(p, grids[l], fam.hp[l]) for p in psi[l]]
        src = d[l - 1] if d_is_f[l - 1] else M.to_fourier(d[l - 1])
        d[l] = add_from_different_grid_with_filter(grids[l], d[l], True, grids[l - 1], src, fam.lp[l - 1])
        d_is_f[l] = False
        newp = []
        for a in range(3):
            s = psi[l - 1][a] if p_is_f[l - 1] else M.to_fourier(psi[l - 1][a])
            newp.append(add_from_different_grid_with_filter(grids[l], psi[l][a], True, grids[l - 1], s, fam.lp[l - 1]))
        psi[l] = newp
        p_is_f[l] = False
    out = []
    for l in range(nl):
        dr = M.to_real(d[l]) if d_is_f[l] else d[l]
        pr = [M.to_real(p) for p in psi[l]] if p_is_f[l] else psi[l]
        out.append((dr, pr))
    return out


# --------------------------------------------------------------------------------------------------------------
# modifications (simulation/modifications/*.hpp)
# --------------------------------------------------------------------------------------------------------------
def multilevel_covector_from_hires(grids, hires_real):
    """generateMultilevelCovectorFromHiresCovector (multilevelgrid.hpp:245-321). Returns per-level real-space fields
    (the function ends with toReal())."""
    nl = len(grids)
    w = level_weights(grids)
    fields = [None] * nl
    fields[-1] = hires_real.copy()
    for l in range(nl - 2, -1, -1):
        r = int(round(grids[l].cell / grids[l + 1].cell))
        fields[l] = deinterpolate_into(grids[l], grids[l + 1], fields[l + 1], 1.0 / float(r ** 3))
    fam = FilterFamily(grids)
    out = []
    for l in range(nl):
        f = fields[l] * (w[l] / w[-1])
        F = apply_filter(M.to_fourier(f), grids[l], fam.filters[l])
        out.append(M.to_real(F))
    return out


def localisation_covector(grid):
    """LinearModification::calculateLocalisationCovector (linearmodification.hpp:78-90), real space."""
    f = np.zeros(grid.n ** 3)
    np.add.at(f, grid.flags, 1.0 / len(grid.flags))
    return f.reshape((grid.n,) * 3)


def velocity_localisation_covector(grid, direction):
    """VelocityModification::calculateLocalisationCovector, real-space 4th-order stencil branch
    (linearmodification.hpp:186-222); steps wrap with the simulation period (getIndexFromIndexAndStep)."""
    n = grid.n
    w = 1.0 / len(grid.flags)
    a, b = -w / 12. / grid.cell, w * 2. / 3. / grid.cell
    f = np.zeros((n, n, n))
    c = np.stack(np.unravel_index(grid.flags, (n, n, n)), axis=0)

    def step(cc, s):
        cc = cc.copy()
        cc[direction] += s
        cc[direction] = np.where(cc[direction] > grid.sim_equiv - 1, cc[direction] - grid.sim_equiv, cc[direction])
        cc[direction] = np.where(cc[direction] < 0, cc[direction] + grid.sim_equiv, cc[direction])
        return cc
    m1, p1 = step(c, -1), step(c, 1)
    m2, p2 = step(m1, -1), step(p1, 1)
    np.add.at(f, tuple(m2), a)
    np.add.at(f, tuple(m1), b)
    np.add.at(f, tuple(p1), -b)
    np.add.at(f, tuple(p2), -a)
    return f


def centre_wrapped(grid):
    """Grid::getCentreWrapped of the flagged cells (grid.hpp:466-494): mean offset from the first flagged cell."""
    n = grid.n
    c = np.stack(np.unravel_index(grid.flags, (n, n, n)), axis=0)
    out = []
    for a in range(3):
        pos = (c[a].astype(np.float64) * grid.cell + grid.offset[a]) + grid.cell / 2
        run = 0.0
        for v in pos[1:]:
            run += float(grid.get_wrapped_offset(v, pos[0]))
        run /= len(pos)
        x = math.fmod(run + pos[0], grid.simsize)
        out.append(x + grid.simsize if x < 0 else x)
    return out


def angular_momentum_localisation_covector(grid, direction):
    """AngMomentumModification::calculateLocalisationCovector (linearmodification.hpp:331-384)."""
    n = grid.n
    qc = centre_wrapped(grid)
    norm = len(grid.flags) * grid.cell
    a, b = 1. / 12. / norm, -2. / 3. / norm
    d1, d2 = (direction + 1) % 3, (direction + 2) % 3
    f = np.zeros((n, n, n))
    c = np.stack(np.unravel_index(grid.flags, (n, n, n)), axis=0)
    q = [(c[ax].astype(np.float64) * grid.cell + grid.offset[ax]) + grid.cell / 2 for ax in range(3)]
    dq = [grid.get_wrapped_offset(qc[ax], q[ax]) for ax in range(3)]   # qcenter - q
    coeff = {direction: np.zeros(len(grid.flags)), d2: dq[d1], d1: -dq[d2]}
    S = grid.sim_equiv

    def step(cc, ax, s):
        cc = cc.copy()
        v = cc[ax] + s
        v = np.where(v > S - 1, v - S, v)
        cc[ax] = np.where(v < 0, v + S, v)
        return cc
    for ax in range(3):
        if ax == direction:
            continue
        m1, p1 = step(c, ax, -1), step(c, ax, 1)
        m2, p2 = step(m1, ax, -1), step(p1, ax, 1)
        np.add.at(f, tuple(m2), coeff[ax] * a)
        np.add.at(f, tuple(m1), coeff[ax] * b)
        np.add.at(f, tuple(p1), -coeff[ax] * b)
        np.add.at(f, tuple(p2), -coeff[ax] * a)
    return f


def angular_momentum_covector(grids, covs, direction, vel_ratio):
    """AngMomentumModification (linearmodification.hpp:288-411): localisation, times C_dm^(1/2), times -scale / k^2."""
    cov = overdensity_covector(grids, covs, angular_momentum_localisation_covector(grids[-1], direction))
    return [_scale_inv_k2(g, F, -vel_ratio) for g, F in zip(grids, cov)]


def potential_covector(grids, covs, Om, a):
    """PotentialModification (linearmodification.hpp:128-146) with cosmology::densityToPotential (parameters.hpp:208-225),
    including its test of the *second wavenumber component* (named k2 there) against zero: the whole ky = 0 plane is
    zeroed, not just k = 0."""
    prefac = 3. / 2. * Om / a * 100. * 100. / 299792. / 299792.
    out = []
    for g, F in zip(grids, overdensity_covector(grids, covs)):
        k2 = k_magnitude(g) ** 2
        ky0 = (np.arange(g.n) == 0)[None, :, None]
        ind = M.independent_mask(g.n)
        with np.errstate(divide="ignore", invalid="ignore"):
            v = np.where(ky0, 0.0, -prefac * F * (1.0 / k2))
        out.append(np.where(ind, v, F))
    return out


def _scale_inv_k2(g, F, scale):
    k2 = k_magnitude(g) ** 2
    ind = M.independent_mask(g.n)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = np.where(k2 == 0, 0.0, scale * F / k2)
    return np.where(ind, v, F)


def overdensity_covector(grids, covs, hires_real=None):
    """LinearModification::getCovector for OverdensityModification (linearmodification.hpp:52-76,118-126):
    localisation -> multilevel covector -> times C_dm^(1/2) per level. covs: per-level covariance arrays.
    Returns Fourier-space fields per level."""
    loc = localisation_covector(grids[-1]) if hires_real is None else hires_real
    # calculateLocalisationCovector ends with toFourier; generateMultilevel... returns to real space
    loc = M.to_real(M.to_fourier(loc))
    ml = multilevel_covector_from_hires(grids, loc)
    return [M.apply_transfer(M.to_fourier(f), c, 0.5) for f, c in zip(ml, covs)]


def velocity_covector(grids, covs, direction, vel_ratio):
    """VelocityModification (linearmodification.hpp:173-285), real-space derivative: -scale * delta / k^2."""
    cov = overdensity_covector(grids, covs, velocity_localisation_covector(grids[-1], direction))
    out = []
    for g, F in zip(grids, cov):
        k2 = k_magnitude(g) ** 2
        ind = M.independent_mask(g.n)
        with np.errstate(divide="ignore", invalid="ignore"):
            v = np.where(k2 == 0, 0.0, -vel_ratio * F / k2)
        out.append(np.where(ind, v, F))
    return out


def orthonormalise(grids, alphas, targets):
    """ModificationManager::orthonormaliseModifications (modificationmanager.hpp:283-304). alphas: covectors."""
    alphas = [list(a) for a in alphas]
    targets = list(targets)
    for i in range(len(alphas)):
        for j in range(i):
            res = inner_product(grids, alphas[i], apply_metric(grids, alphas[j], False))
            alphas[i] = add_scaled(alphas[i], alphas[j], -res)
            targets[i] -= res * targets[j]
        norm = math.sqrt(inner_product(grids, alphas[i], apply_metric(grids, alphas[i], False)))
        alphas[i] = [x / norm for x in alphas[i]]
        targets[i] /= norm
    return alphas, targets


def apply_linear_modifications(grids, field, alphas, targets):
    """applyLinearModif (modificationmanager.hpp:199-217) after orthonormalisation. Returns the modified field and the
    orthonormalised family as VECTORS: the reference converts each covector in place (alpha_i.convertToVector()), and
    it is these converted objects that performIterations later receives."""
    alphas, targets = orthonormalise(grids, alphas, targets)
    existing = [inner_product(grids, a, field) for a in alphas]
    vecs = []
    for a, t, e in zip(alphas, targets, existing):
        v = apply_metric(grids, a, False)
        vecs.append(v)
        field = add_scaled(field, v, t - e)
    return field, vecs


def variance_push(grids, covs, field, scale):
    """QuadraticModification::pushMultiLevelFieldThroughMatrix with FilteredVarianceModification
    (quadraticmodification.hpp:52-79,152-221). field: white-noise vector per level (Fourier). Returns covector."""
    kcut = 2 * math.pi / scale
    hp = [(FERMI_HP, kcut, 1)]
    res = []
    for g, F, c in zip(grids, field, covs):
        n = g.n
        x = M.to_real(M.apply_transfer(F, c, 0.5))
        m = np.zeros(n ** 3, dtype=bool)
        m[g.flags] = True
        m = m.reshape(n, n, n)
        x = np.where(m, x, 0.0)
        x = M.to_real(apply_filter(M.to_fourier(x), g, hp))
        x = np.where(m, x, 0.0)
        cnt = len(g.flags)
        if cnt:
            s = float(np.sum(x.reshape(-1)[g.flags]))
            x = np.where(m, (x - s / cnt) / cnt, 0.0)
        x = M.to_real(apply_filter(M.to_fourier(x), g, hp))
        x = np.where(m, x, 0.0)
        res.append(M.to_fourier(x))
    res = apply_metric(grids, res, True)
    return [M.apply_transfer(F, c, 0.5) for F, c in zip(res, covs)]


def variance_value(grids, covs, field, scale):
    """QuadraticModification::calculateCurrentValue (quadraticmodification.hpp:43-49)."""
    return inner_product(grids, variance_push(grids, covs, field, scale), field)


def variance_iterations(grids, covs, field, alpha_vecs, target, scale, n_steps):
    """ModificationManager::performIterations (modificationmanager.hpp:249-272). alpha_vecs: the orthonormal family as
    left by applyLinearModif (vectors); addToOrthonormalFamily (:306-319) projects the pushed covector against them."""
    start = variance_value(grids, covs, field, scale)
    if n_steps in (0, 1) or start == target:       # tools::linspace (util_functions.hpp:78-92)
        targets = [target]
    else:
        step = (target - start) / n_steps
        targets = [start + c * step for c in range(n_steps + 1)]
    for i in range(n_steps):
        cur = variance_value(grids, covs, field, scale)
        pushed = variance_push(grids, covs, field, scale)
        norm = math.sqrt(inner_product(grids, pushed, apply_metric(grids, pushed, False)))
        for a in alpha_vecs:
            r = inner_product(grids, pushed, a)
            pushed = add_scaled(pushed, a, -r)
        nrm = math.sqrt(inner_product(grids, pushed, apply_metric(grids, pushed, False)))
        pushed = [x / nrm for x in pushed]
        mult = 0.5 * (targets[i + 1] - cur) / norm
        field = add_scaled(field, apply_metric(grids, pushed, False), mult)
    return field


def apply_variance_modification(grids, covs, field, alpha_vecs, target, scale, init_steps=10, precision=0.001):
    """applyLinQuadModif for one quadratic modification (modificationmanager.hpp:219-247)."""
    test = variance_iterations(grids, covs, [f.copy() for f in field], alpha_vecs, target, scale, init_steps)
    achieved = abs(variance_value(grids, covs, test, scale) - target)
    n_steps = init_steps * int(math.ceil(math.sqrt(achieved / (target * precision))))
    return variance_iterations(grids, covs, field, alpha_vecs, target, scale, n_steps if n_steps > init_steps else init_steps)


# --------------------------------------------------------------------------------------------------------------
# splicing (simulation/modifications/splice.hpp:38-118, tools/numerics/cg.hpp:14-58)
# --------------------------------------------------------------------------------------------------------------
def _padded_dot(a, b):
    """Field::innerProduct in real space (field.hpp:365-377) runs over the whole padded storage; the fields here
    are unpadded so the quirk (stale padding entering the dot product) is dropped -- DESIGN.md."""
    return float(np.sum(a * b))


def conjugate_gradient(Q, b, rtol=1e-6, atol=1e-12, log=None):
    residual = b.copy()
    direction = -residual
    x = np.zeros_like(b)
    scale = math.sqrt(_padded_dot(b, b))
    if scale == 0.0:
        return x
    dim = b.size
    for i in range(dim + 1):
        Qd = Q(direction)
        alpha = -_padded_dot(residual, direction) / _padded_dot(direction, Qd)
        x = x + alpha * direction
        residual = Q(x) - b
        norm = math.sqrt(_padded_dot(residual, residual))
        if norm < rtol * scale or norm < atol:
            break
        if log is not None:
            log.append(norm)
        beta = _padded_dot(residual, Qd) / _padded_dot(direction, Qd)
        direction = direction * beta - residual
    return x


def splice_one_level(grid, a_F, b_F, cov):
    """modifications::spliceOneLevel(a = new field, b = original field, cov). mask = flagged cells."""
    n = grid.n
    mask = np.zeros(n ** 3)
    mask[grid.flags] = 1.0
    mask = mask.reshape(n, n, n)
    comp = 1.0 - mask
    pre = cov.copy()
    pre[0, 0, 0] = 1.0

    def T(xr, p):
        return M.to_real(M.apply_transfer(M.to_fourier(xr), pre, p))

    delta_diff = M.to_real(M.apply_transfer(b_F - a_F, pre, 0.5))
    z = delta_diff * mask
    z = T(z, -1.0) * comp
    z = T(z, 0.5)

    def X(v):
        v = T(v, 0.5) * comp
        v = T(v, -1.0) * comp
        return T(v, 0.5)

    alpha = conjugate_gradient(X, z)
    alpha = T(alpha, 0.5)
    b_delta = M.to_real(M.apply_transfer(b_F, pre, 0.5))
    alpha = alpha * comp + b_delta
    alpha = alpha - delta_diff * mask
    return M.apply_transfer(M.to_fourier(alpha), pre, -0.5)
