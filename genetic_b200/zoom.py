"""Host logic around the zoom levels: cell selection and zoom-window placement (pure index arithmetic, no field data).

Reference (paths relative to genetIC/src/):
  ICGenerator::select / selectSphere / selectCube / selectNearest     ic.hpp:1447-1524
  ICGenerator::initZoomGrid                                            ic.hpp:398-479
  ICGenerator::initZoomGridWithLowLeftCornerAt (flag transfer)         ic.hpp:502-598
  Window<int>                                                          simulation/window.hpp
"""
import math

import numpy as np

from . import lib as L


def _centroids(grid, axis):
    """Grid::getCentroidFromCoordinate (grid.hpp:662-669): coord*cellSize, + offsetLower, + cellSize/2, in the grid's type."""
    T = grid.dtype.type
    c = np.arange(grid.size).astype(grid.dtype)
    return (c * T(grid.cellSize) + T(grid.offsetLower[axis])) + T(grid.cellSize) / T(2)


def _wrapped_offset(base, a, b):
    """Grid::getWrappedOffset with the base grid's period (ICGenerator::get_wrapped_delta, ic.hpp:1368-1370)."""
    d = np.asarray(b) - np.asarray(a)
    Lbox = base.periodicDomainSize
    d = np.where(d < -Lbox / 2, d + Lbox, d)
    return np.where(d > Lbox / 2, d - Lbox, d)


def select(context, centre, inclusion):
    """ICGenerator::select (ic.hpp:1458-1484): on every level flag the cells whose centroid offset from `centre`
    satisfies inclusion(dx, dy, dz)."""
    base = context.getGridForLevel(0)
    outputGrids = [context.getOutputGridForLevel(l) for l in range(context.getNumLevels())]
    for g in outputGrids:
        g.unflagAllCells()
    for g in outputGrids:
        d = [_wrapped_offset(base, _centroids(g, a), g.dtype.type(centre[a])) for a in range(3)]
        m = inclusion(d[0][:, None, None], d[1][None, :, None], d[2][None, None, :])
        g.flagCells(np.flatnonzero(np.broadcast_to(m, (g.size,) * 3).reshape(-1)))


def selectSphere(context, centre, radius):
    r2 = float(np.float32(radius) * np.float32(radius))   # selectSphere(float radius), ic.hpp:1490-1496
    select(context, centre, lambda x, y, z: x * x + y * y + z * z < r2)


def selectCube(context, centre, side):
    h = float(np.float32(side) / np.float32(2))           # selectCube(float side), ic.hpp:1519-1524
    select(context, centre, lambda x, y, z: (np.abs(x) < h) & (np.abs(y) < h) & (np.abs(z) < h))


def selectEllipse(context, centre, a, b, c):
    a2, b2, c2 = (float(np.float32(v) * np.float32(v)) for v in (a, b, c))   # selectEllipse(float a, b, c), ic.hpp:1503-1513
    select(context, centre, lambda x, y, z: x * x / a2 + y * y / b2 + z * z / c2 < 1.)


def getWrappedOffset(grid, x0, x1):
    """Grid::getWrappedOffset(x0, x1) = x0 - x1 wrapped into [-L/2, L/2] (grid.hpp:445-454)."""
    d = np.asarray(x0, dtype=np.float64) - np.asarray(x1, dtype=np.float64)
    Lbox = float(grid.periodicDomainSize)
    d = np.where(d > Lbox / 2, d - Lbox, d)
    return np.where(d < -Lbox / 2, d + Lbox, d)


def getCentreWrapped(grid, flags):
    """Grid::getCentreWrapped (grid.hpp:466-494): the mean of the flagged cells' wrapped offsets from the first one."""
    if len(flags) == 0:
        raise L.GxError("Cannot calculate the center of an empty region")
    n = grid.size
    c = np.stack(np.unravel_index(np.asarray(flags), (n, n, n)), axis=0)
    out = []
    for a in range(3):
        pos = (c[a].astype(np.float64) * float(grid.cellSize) + float(grid.offsetLower[a])) + float(grid.cellSize) / 2
        run = 0.0
        for v in pos[1:]:
            run += float(getWrappedOffset(grid, v, pos[0]))
        run /= len(pos)
        x = math.fmod(run + pos[0], float(grid.periodicDomainSize))
        out.append(x + float(grid.periodicDomainSize) if x < 0 else x)
    return out


def selectNearest(context, centre):
    """ICGenerator::selectNearest (ic.hpp:1447-1455): the cell of the deepest level that holds `centre`
    (Grid::getIndexFromPoint, grid.hpp:672-677)."""
    for l in range(context.getNumLevels()):
        context.getOutputGridForLevel(l).unflagAllCells()
        context.getGridForLevel(l).unflagAllCells()
    g = context.grids[-1]
    c = []
    for a in range(3):
        x = math.fmod(float(centre[a]) - float(g.offsetLower[a]), float(g.periodicDomainSize))
        if x < 0:
            x += float(g.periodicDomainSize)
        c.append(int(math.floor(x / float(g.cellSize))))
    g.flagCells([(c[0] * g.size + c[1]) * g.size + c[2]])


class _Window:
    """Window<int> (window.hpp): a periodic cuboid [lower, upper) per axis."""

    def __init__(self, wrap, point):
        self.wrap, self.lo, self.hi = wrap, list(point), [p + 1 for p in point]

    def _w(self, v):
        if v < 0:
            v += self.wrap
        if v >= self.wrap:
            v -= self.wrap
        return v

    @staticmethod
    def _within(lo, hi, v):
        return (lo <= v < hi) if hi >= lo else (v < hi or v >= lo)

    def _expandAxis(self, a, v):
        lo, hi = self.lo[a], self.hi[a]
        if self._within(lo, hi, v):
            return
        sizeLower = self._w(hi - v - 1) + 1
        sizeUpper = self._w(v - lo) + 1
        canLower = self._within(v, hi, lo)
        canUpper = self._within(lo, v + 1, hi - 1)
        if canLower and (sizeLower < sizeUpper or not canUpper):
            self.lo[a] = v
        else:
            self.hi[a] = v + 1

    def expandToInclude(self, p):
        for a in range(3):
            self._expandAxis(a, p[a])

    def expandToIncludeAll(self, coords):
        """expandToInclude over the rows of an (m, 3) array, in order. The axes are independent, and while the window is a
        proper arc of the period it only grows, so a value that has been seen changes nothing when it comes again: per axis
        only the first occurrence of each distinct value is visited. If the arc ever closes on itself (flags spanning the
        whole period: [lo, hi) with hi == lo reads as empty) that argument fails and the rows are replayed one by one."""
        coords = np.asarray(coords)
        lo0, hi0 = list(self.lo), list(self.hi)
        for a in range(3):
            vals, first = np.unique(coords[:, a], return_index=True)
            for v in vals[np.argsort(first, kind="stable")]:
                self._expandAxis(a, int(v))
                if (self.hi[a] - self.lo[a]) % self.wrap == 0:
                    self.lo, self.hi = lo0, hi0
                    for c in coords:
                        self.expandToInclude([int(x) for x in c])
                    return

    def sizes(self):
        return [self._w(self.hi[a] - self.lo[a] - 1) + 1 for a in range(3)]

    def wrapCorners(self):
        self.lo = [self._w(v) for v in self.lo]
        self.hi = [self._w(v - 1) + 1 for v in self.hi]

    def shift(self, d):
        self.lo = [self.lo[a] + d[a] for a in range(3)]
        self.hi = [self.hi[a] + d[a] for a in range(3)]
        self.wrapCorners()

    def expandSymmetricallyToSize(self, new):
        old = list(self.lo)
        if new == self.wrap:
            self.lo, self.hi = [0, 0, 0], [new] * 3
            return
        c = [self._w(self.lo[a] + self._w(self.hi[a] - 1 - self.lo[a]) // 2) for a in range(3)]
        self.lo = [self._w(c[a] - new // 2) for a in range(3)]
        up = new // 2 - 1 if new % 2 == 0 else new // 2
        self.hi = [self._w(c[a] + up) + 1 for a in range(3)]
        # (the reference compares the y corner against the old *z* corner, window.hpp; kept)
        for a, ref in ((0, old[0]), (1, old[2]), (2, old[2])):
            if self.lo[a] > ref:
                self.lo[a] -= 1
                self.hi[a] -= 1

    def wrappedOffset(self, a, b):
        d = b - a
        if d < -(self.wrap // 2):
            d += self.wrap
        if d > self.wrap // 2:
            d -= self.wrap
        return d


def zoomCorner(grid, zoomfac, allowStrays=False, out=None):
    """ICGenerator::initZoomGrid (ic.hpp:398-479): the lower-left corner, in cells of `grid` (the grid the fields of the level
    above live on), of the zoom window of grid.size / zoomfac cells that holds the flagged cells. `out` is the output grid of
    the level above when it differs from `grid` (stray particles: a ResolutionMatchingGrid spanning the level above that
    one): the flags are its cells, and the window found on it is moved into `grid`."""
    n = grid.size
    out = grid if out is None else out
    flags = out.getFlaggedCells()
    if len(flags) == 0:
        raise L.GxError("Cannot initialise zoom without marking particles to be replaced")
    no = out.size
    coords = np.stack(np.unravel_index(flags, (no, no, no)), axis=1)
    wrap = grid.simEquivalentSize
    w = _Window(wrap, [int(v) for v in coords[0]])
    w.expandToIncludeAll(coords)
    n_user = n // zoomfac
    n_required = max(w.sizes())
    if n_required > n_user and not allowStrays:
        raise L.GxError("Zoom particles do not fit in specified sub-box. Decrease zoom, or choose different particles")
    if n_required < n_user:
        w.expandSymmetricallyToSize(n_user)
    lo_out = [int(round(float(out.offsetLower[a]) / float(grid.cellSize))) for a in range(3)]
    w.shift(lo_out)                         # relative to the origin of the simulation (outputWindowAbove's corner)
    lo = [int(round(float(grid.offsetLower[a]) / float(grid.cellSize))) for a in range(3)]
    above = _Window(wrap, lo)
    above.hi = [lo[a] + n for a in range(3)]
    above.wrapCorners()
    move = [0, 0, 0]
    for a in range(3):
        if not above._within(above.lo[a], above.hi[a], w.lo[a]):
            move[a] = w.wrappedOffset(w.lo[a], above.lo[a])
        if not above._within(above.lo[a], above.hi[a], w.hi[a] - 1):
            move[a] = w.wrappedOffset(w.hi[a], above.hi[a])
    w.shift(move)
    w.shift([-v for v in above.lo])
    return tuple(w.lo)


def flagsNearZoomBorder(grid, corner, zoomfac, borderSafety=3, out=None):
    """ICGenerator::initZoomGrid (ic.hpp:461-474): True when a flagged cell lies within `borderSafety` cells of the edge
    of the zoom window placed at `corner` (Window<int>::containsWithBorderSafety, window.hpp:296-302)."""
    n = grid.size
    w = _Window(grid.simEquivalentSize, list(corner))
    w.hi = [corner[a] + n // zoomfac for a in range(3)]
    w.wrapCorners()
    out = grid if out is None else out
    coords = list(np.unravel_index(out.getFlaggedCells(), (out.size,) * 3))
    for a in range(3):                       # offsetBetweenOutputAndActualWindow (ic.hpp:463)
        coords[a] = coords[a] - (int(round(float(grid.offsetLower[a]) / float(grid.cellSize)))
                                 - int(round(float(out.offsetLower[a]) / float(grid.cellSize))))
    ok = np.ones(len(coords[0]), dtype=bool)
    for a in range(3):
        lo, hi = w.lo[a] + borderSafety, w.hi[a] - borderSafety
        ok &= ((coords[a] < hi) & (coords[a] >= lo)) if hi >= lo else ((coords[a] < hi) | (coords[a] >= lo))
    return not ok.all()


def _trimToZoomWindow(up, flags, corner, nCoarse, out=None):
    """The trimming of initZoomGridWithLowLeftCornerAt (ic.hpp:531-554): flagged cells of the level above whose centroid
    lies outside the new window [centroid(corner) - eps, centroid(corner + nCoarse - 1) + eps] (Window<T>, periodic) are
    dropped from the list of cells to replace."""
    T = up.dtype.type
    n = up.size
    hi_cell = [corner[a] + nCoarse for a in range(3)]
    if up.coversFullSimulation():
        hi_cell = [(h - 1) % n + 1 for h in hi_cell]
    elif max(hi_cell) > n:
        raise L.GxError("The newly initialised zoom level is not fully contained within the previous level")
    eps = T(up.cellSize * T(1e-6))
    wrapLength = up.periodicDomainSize

    def wrap(v):
        if v < 0:
            v = T(v + wrapLength)
        if v >= wrapLength:
            v = T(v - wrapLength)
        return v

    def centroid(c, a):
        return T(T(T(T(c) * up.cellSize) + up.offsetLower[a]) + up.cellSize / T(2))

    keep = np.ones(len(flags), dtype=bool)
    out = up if out is None else out         # the flags are cells of the output grid above (ic.hpp:545)
    coords = np.unravel_index(np.asarray(flags, dtype=np.int64), (out.size,) * 3)
    for a in range(3):
        lo = wrap(T(centroid(corner[a], a) - eps))
        hi = T(centroid(hi_cell[a] - 1, a) + eps)
        hi = np.nextafter(wrap(np.nextafter(hi, T(-np.inf))), T(np.inf))
        cen = (coords[a].astype(up.dtype) * out.cellSize + out.offsetLower[a]) + out.cellSize / T(2)
        keep &= ((cen < hi) & (cen >= lo)) if hi >= lo else ((cen < hi) | (cen >= lo))
    return np.asarray(flags, dtype=np.int64)[keep]


def initZoomGrid(context, zoomfac, n, corner=None):
    """initZoomGrid / initZoomGridWithLowLeftCornerAt: appends the zoom level, records the cells of the level above that
    the particle mapper will replace (context.zoomParticleArray, trimmed to the zoom window: ic.hpp:531-571) and carries
    the flags of the level above onto the new grid (the SuperSample + Section view of the parent, grid.hpp:195-228,
    virtualgrid.hpp:161-166,489-498)."""
    up = context.grids[-1]
    out = context.getOutputGridForLevel(len(context.grids) - 1)   # differs from `up` below a zoom opened with strays_on
    if len(out.getFlaggedCells()) == 0:
        raise L.GxError("Cannot initialise zoom without marking particles to be replaced")
    strays = getattr(context, "allowStrays", False)
    if corner is None:
        corner = zoomCorner(up, zoomfac, strays, out)
    if min(corner) < 0:
        raise L.GxError("The newly initialised zoom level is not fully contained within the previous level")
    if not hasattr(context, "zoomParticleArray"):
        context.zoomParticleArray = []
    del context.zoomParticleArray[len(context.grids) - 1:]
    trimmed = _trimToZoomWindow(up, out.getFlaggedCells(), corner, up.size // zoomfac, out)
    context.zoomMissedParticles = len(out.getFlaggedCells()) - len(trimmed)
    # with strays the cells outside the window stay in the list: their particles are interpolated from the level above
    context.zoomParticleArray.append(np.asarray(out.getFlaggedCells(), dtype=np.int64).copy() if strays else trimmed)
    g = context.addZoomLevel(zoomfac, n, corner)
    ratio = int(round(float(up.cellSize) / float(g.cellSize)))
    pm = np.zeros(up.size3, dtype=bool)
    pm[up.getFlaggedCells()] = True
    pm = pm.reshape((up.size,) * 3)
    idx = [((np.arange(n) // ratio) + corner[a]) % up.simEquivalentSize for a in range(3)]
    inside = [i < up.size for i in idx]
    sub = pm[np.ix_(idx[0] % up.size, idx[1] % up.size, idx[2] % up.size)].copy()
    sub &= inside[0][:, None, None] & inside[1][None, :, None] & inside[2][None, None, :]
    g.flagCells(np.flatnonzero(sub.reshape(-1)))
    return g


def proxyFlags(source, target):
    """Grid::makeProxyGridToMatch(target)->getFlaggedCells (grid.hpp:195-228; SuperSampleGrid / SectionOfGrid /
    SubSampleGrid flag proxies, virtualgrid.hpp): the flagged cells of `source` expressed as cells of `target` -- refined
    first if the target is finer, shifted into the target's frame (periodic in units of the simulation box, cells outside
    the target dropped), coarsened last if the target is coarser."""
    flags = np.asarray(source.getFlaggedCells(), dtype=np.int64)
    n, dx = source.size, float(source.cellSize)
    c = np.stack(np.unravel_index(flags, (n, n, n)), axis=0) if len(flags) else np.zeros((3, 0), dtype=np.int64)
    if float(target.cellSize) < dx:
        r = _ratio(dx, float(target.cellSize))
        o = np.arange(r, dtype=np.int64)
        c = np.stack([(c[0][:, None, None, None] * r + o[None, :, None, None]) + 0 * o[None, None, :, None] + 0 * o[None, None, None, :],
                      (c[1][:, None, None, None] * r + o[None, None, :, None]) + 0 * o[None, :, None, None] + 0 * o[None, None, None, :],
                      (c[2][:, None, None, None] * r + o[None, None, None, :]) + 0 * o[None, :, None, None] + 0 * o[None, None, :, None]],
                     axis=0).reshape(3, -1)
        dx /= r
        n *= r
    sub = _ratio(float(target.cellSize), dx)
    pre = target.size * sub
    off = [_ratio_signed(float(target.offsetLower[a]) - float(source.offsetLower[a]), dx) for a in range(3)]
    if any(off) or pre != n:
        wrap = int(round(float(source.periodicDomainSize) / dx))
        c = np.stack([(c[a] - off[a]) % wrap for a in range(3)], axis=0)
        keep = np.all((c >= 0) & (c < pre), axis=0)
        c = c[:, keep]
    c = c // sub
    m = target.size
    return np.unique((c[0] * m + c[1]) * m + c[2])


def _ratio(a, b):
    r = a / b
    if abs(r - round(r)) > 1e-8 * max(1.0, abs(r)) or round(r) < 1:
        raise L.GxError("The ratio of cell sizes is not a positive integer")
    return int(round(r))


def _ratio_signed(a, b):
    r = a / b
    if abs(r - round(r)) > 1e-6 * max(1.0, abs(r)):
        raise L.GxError("Grid offsets do not differ by a whole number of cells")
    return int(round(r))
