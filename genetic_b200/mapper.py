"""Particle mappers: which cell of which level each output particle comes from, in the reference's output order.

Reference (paths relative to genetIC/src/simulation/particles/mapper/):
  ParticleMapper                         mapper.hpp
  OneLevelParticleMapper                 onelevelmapper.hpp:38-41        particle id = cell index
  TwoLevelParticleMapper                 twolevelmapper.hpp:24-899       level-1 particles minus the zoomed ones, then for
                                                                         every zoomed cell (ascending cell id) the block of
                                                                         fine cells inside it (Grid::insertCubeIdsIntoVector,
                                                                         grid.hpp:704-731: x outermost, z innermost)
  ICGenerator::updateParticleMapper      ic.hpp:1129-1221                left-expanding tree over the levels

The reference walks these with iterators that dereference one particle at a time (mapperiterator.hpp:151-156). Here a
mapper is resolved once into its *layout* -- two arrays (level, cell index) in output order, pure index arithmetic on the
host -- and the particles of each level are then evaluated on the device at that level's cell list in one launch per
component (gx_emit_cells). Not covered: gas splitting (AddGasMapper), sub/supersampling, intermediate-resolution padding,
stray particles, the grafic mapper.
"""
import ctypes as C
import math

import numpy as np

from . import lib as L
from .host import _dt


class OneLevelParticleMapper:
    """onelevelmapper.hpp: every cell of one grid, in cell-index order."""

    def __init__(self, grid, level, gadgetParticleType=1):
        self.grid, self.level, self.gadgetParticleType = grid, level, gadgetParticleType

    def size(self):
        return self.grid.size3

    def getFinestGrid(self):
        return self.grid

    def getCoarsestGrid(self):
        return self.grid

    def getFinestLevel(self):
        return self.level

    def grids(self):
        return {self.level: self.grid}

    def gadgetTypes(self):
        return {self.level: self.gadgetParticleType}

    def layout(self):
        return (np.full(self.grid.size3, self.level, dtype=np.int32), np.arange(self.grid.size3, dtype=np.int64))

    def unflagAllParticles(self):
        self.grid.unflagAllCells()

    def flagParticles(self, orderedParticleIndices):
        """onelevelmapper.hpp:98-100"""
        ids = np.asarray(orderedParticleIndices, dtype=np.int64)
        if len(ids) and (ids.min() < 0 or ids.max() >= self.grid.size3):
            raise L.GxError("Particle ID out of range")
        self.grid.flagCells(ids, append=True)

    def getFlaggedParticles(self):
        """onelevelmapper.hpp:103-106"""
        return np.sort(self.grid.getFlaggedCells())


def _wrap(x, period):
    """Grid::wrapIndividualCoordinate (grid.hpp:527-531)."""
    x = np.fmod(x, period)
    return np.where(x < 0, x + period, x)


def cube_ids(fine, x0c, dxc):
    """Grid::insertCubeIdsIntoVector (grid.hpp:704-731) for arrays of cube centres x0c[3][m] of side dxc: the ids of the
    cells of `fine` whose centres lie in each cube, [m][k] with x outermost and z innermost."""
    T = fine.dtype.type
    cs = fine.cellSize
    lo, hi = [], []
    for a in range(3):
        x = _wrap(np.asarray(x0c[a], dtype=fine.dtype) - fine.offsetLower[a], fine.periodicDomainSize)
        lo.append(np.floor((x - T(dxc) / T(2) + cs / T(2)) / cs).astype(np.int64))
        hi.append(np.floor((x + T(dxc) / T(2) - cs / T(2)) / cs).astype(np.int64))
    bad = np.zeros(len(lo[0]), dtype=bool)
    for a in range(3):
        bad |= (lo[a] < 0) | (hi[a] >= fine.size)
    if bad.any():
        raise L.GxError("Requested zoom region falls outside high resolution grid")
    w = [hi[a] - lo[a] + 1 for a in range(3)]
    r = int(w[0][0]) if len(w[0]) else 0
    if any((wa != r).any() for wa in w):
        raise L.GxError("Zoomed cells do not hold a fixed number of fine cells")
    o = np.arange(r, dtype=np.int64)
    n = fine.size
    ids = ((lo[0][:, None, None, None] + o[None, :, None, None]) * n
           + (lo[1][:, None, None, None] + o[None, None, :, None])) * n + (lo[2][:, None, None, None] + o[None, None, None, :])
    return ids.reshape(len(lo[0]), r ** 3)


class TwoLevelParticleMapper:
    """twolevelmapper.hpp: pLevel1 (any mapper) with the cells `level1CellsToReplace` of its finest grid replaced by the
    cells of pLevel2 (a OneLevelParticleMapper) that lie inside them."""

    def __init__(self, pLevel1, pLevel2, level1CellsToReplace, skipLevel1=False):
        if not isinstance(pLevel2, OneLevelParticleMapper):
            raise L.GxError("TwoLevelParticleMapper: level 2 must be a OneLevelParticleMapper")
        self.pLevel1, self.pLevel2, self.skipLevel1 = pLevel1, pLevel2, skipLevel1
        self.pGrid1, self.pGrid2 = pLevel1.getFinestGrid(), pLevel2.getCoarsestGrid()
        self.level1CellsToReplace = np.unique(np.asarray(level1CellsToReplace, dtype=np.int64))
        if len(self.level1CellsToReplace) == 0:
            raise L.GxError("TwoLevelParticleMapper: no cells to replace")
        lev, cell = pLevel1.layout()
        on_finest = np.flatnonzero(lev == pLevel1.getFinestLevel())
        hit = np.isin(cell[on_finest], self.level1CellsToReplace)
        self.level1ParticlesToReplace = on_finest[hit]           # ascending particle ids of level 1
        if len(self.level1ParticlesToReplace) != len(self.level1CellsToReplace):
            raise L.GxError("Consistency error when relating cells to particles")
        ratio = float(self.pGrid1.cellSize) / float(self.pGrid2.cellSize)
        if abs(ratio - round(ratio)) > 1e-8 or round(ratio) < 1:       # tools::getRatioAndAssertPositiveInteger
            raise L.GxError("TwoLevelParticleMapper: the cell size ratio is not a positive integer")
        self.n_hr_per_lr = int(round(ratio)) ** 3
        self.firstLevel2Particle = 0 if skipLevel1 else pLevel1.size() - len(self.level1ParticlesToReplace)
        self.totalParticles = self.firstLevel2Particle + self.n_hr_per_lr * len(self.level1ParticlesToReplace)
        # calculateHiresParticleList (twolevelmapper.hpp:866-893): blocks in ascending coarse cell id, not sorted further
        g1 = self.pGrid1
        T = g1.dtype.type
        c = np.unravel_index(self.level1CellsToReplace, (g1.size,) * 3)
        cen = [(ci.astype(g1.dtype) * g1.cellSize + g1.offsetLower[a]) + g1.cellSize / T(2) for a, ci in enumerate(c)]
        ids = cube_ids(self.pGrid2, cen, g1.cellSize)
        if ids.shape[1] != self.n_hr_per_lr:
            raise L.GxError("Zoomed cells do not hold n_hr_per_lr fine cells")
        self.zoomParticleArrayHiresUnsorted = ids.reshape(-1)
        keep = np.ones(len(lev), dtype=bool)
        keep[self.level1ParticlesToReplace] = False
        if skipLevel1:
            keep[:] = False
        self._layout = (np.concatenate([lev[keep], np.full(len(self.zoomParticleArrayHiresUnsorted), pLevel2.level, np.int32)]),
                        np.concatenate([cell[keep], self.zoomParticleArrayHiresUnsorted]))

    def size(self):
        return self.totalParticles

    def getFinestGrid(self):
        return self.pLevel2.getFinestGrid()

    def getCoarsestGrid(self):
        return self.pLevel1.getCoarsestGrid()

    def getFinestLevel(self):
        return self.pLevel2.level

    def grids(self):
        g = dict(self.pLevel1.grids())
        g.update(self.pLevel2.grids())
        return g

    def gadgetTypes(self):
        g = dict(self.pLevel1.gadgetTypes())
        g.update(self.pLevel2.gadgetTypes())
        return g

    def layout(self):
        return self._layout

    def unflagAllParticles(self):
        self.pLevel1.unflagAllParticles()
        self.pLevel2.unflagAllParticles()

    def flagParticles(self, orderedParticleIndices):
        """twolevelmapper.hpp:358-452: particle ids of this mapper -> flags on the grids. A low-resolution id flags its
        level-1 particle; a high-resolution id flags its fine cell AND the level-1 particle it replaced."""
        ids = np.asarray(orderedParticleIndices, dtype=np.int64)
        if len(ids) > 1 and (np.diff(ids) < 0).any():
            raise L.GxError("The particle list must be in ascending order")
        nZoom = len(self.level1ParticlesToReplace)
        if self.pLevel1.size() < nZoom:
            raise L.GxError("Zoom particle list is longer than the grid it refers to")
        firstHires = self.pLevel1.size() - nZoom
        kept = np.ones(self.pLevel1.size(), dtype=bool)
        kept[self.level1ParticlesToReplace] = False
        kept = np.flatnonzero(kept)                               # level-1 particle id of every unzoomed particle
        lr, hr = ids[ids < firstHires], ids[ids >= firstHires]
        lr_index = (hr - firstHires) // self.n_hr_per_lr
        if len(lr_index) and lr_index.max() >= nZoom:
            raise L.GxError("Particle ID out of range")
        level1 = np.sort(np.concatenate([kept[lr], self.level1ParticlesToReplace[lr_index]]))
        level2 = np.sort(self.zoomParticleArrayHiresUnsorted[hr - firstHires])
        self.pLevel1.flagParticles(level1)
        self.pLevel2.flagParticles(level2)

    def getFlaggedParticles(self):
        """twolevelmapper.hpp:455-494: flags on the grids -> ascending particle ids of this mapper (flagged level-1
        particles that were not zoomed, flagged fine cells that the zoom region holds)."""
        g1 = self.pLevel1.getFlaggedParticles()
        g1 = g1[~np.isin(g1, self.level1ParticlesToReplace)]
        out = [g1 - np.searchsorted(self.level1ParticlesToReplace, g1)]
        g2 = self.pLevel2.getFlaggedParticles()
        sortIndex = np.argsort(self.zoomParticleArrayHiresUnsorted, kind="stable")
        srt = self.zoomParticleArrayHiresUnsorted[sortIndex]
        pos = np.searchsorted(srt, g2)
        ok = (pos < len(srt))
        ok[ok] = srt[pos[ok]] == g2[ok]
        out.append(sortIndex[pos[ok]] + self.firstLevel2Particle)
        return np.sort(np.concatenate(out))


def updateParticleMapper(context, gadgetTypesForLevels=None):
    """ICGenerator::updateParticleMapper (ic.hpp:1129-1176) for dark matter without sub/supersampling: the one-level
    mapper of the base grid, wrapped once per zoom level with that level's zoomParticleArray (stored on the context by
    zoom.initZoomGrid)."""
    nl = context.getNumLevels()
    types = list(gadgetTypesForLevels) if gadgetTypesForLevels is not None else [1] * nl
    m = OneLevelParticleMapper(context.getGridForLevel(0), 0, types[0])
    for level in range(1, nl):
        fine = OneLevelParticleMapper(context.getGridForLevel(level), level, types[level])
        m = TwoLevelParticleMapper(m, fine, context.zoomParticleArray[level - 1])
    return m


def particle_mass(grid, cosmology):
    """ZeldovichParticleEvaluator::getMass (zeldovich.hpp:55-62,108-110): boxMass * cellMassFrac, cellMassFrac =
    pow(dx / simsize, 3) (grid.hpp:112), in the grid's type."""
    T = grid.dtype.type
    boxMass = T(T(27.744948) * T(cosmology.OmegaM0)) * T(math.pow(float(grid.periodicDomainSize), 3.0))
    return T(boxMass * T(math.pow(float(T(grid.cellSize / grid.periodicDomainSize)), 3.0)))


def particle_eps(grid, epsNorm=0.01075):
    """getEps (zeldovich.hpp:113-115): cellSize * cellSofteningScale(=1) * epsNorm."""
    T = grid.dtype.type
    return T(T(grid.cellSize * T(1.0)) * T(epsNorm))


class MappedParticles:
    """All particles of a mapper, evaluated level by level on the device and placed in mapper order on the host:
    pos[N,3], vel[N,3] (out_dtype), mass[N], eps[N] (grid type), level[N], cell[N]."""

    def __init__(self, mapper, particleGenerator, cosmology, out_dtype=np.float32, epsNorm=0.01075):
        lib = L.load()
        lev, cell = mapper.layout()
        N = len(lev)
        self.level, self.cell = lev, cell
        self.pos = np.empty((N, 3), dtype=out_dtype)
        self.vel = np.empty((N, 3), dtype=out_dtype)
        grids = mapper.grids()
        T = next(iter(grids.values())).dtype.type
        self.mass = np.empty(N, dtype=T)
        self.eps = np.empty(N, dtype=T)
        odt, esz = _dt(out_dtype), np.dtype(out_dtype).itemsize
        for level, grid in grids.items():
            where = np.flatnonzero(lev == level)
            self.mass[where] = particle_mass(grid, cosmology)
            self.eps[where] = particle_eps(grid, epsNorm)
            m = len(where)
            if m == 0:
                continue
            cells = np.ascontiguousarray(cell[where])
            psis = particleGenerator.getGeneratorForLevel(level).getGeneratedFields()
            ratio = cosmology.zeldovichVelocityToOffsetRatio(grid.dtype)
            dcells, dpos, dvel = C.c_void_p(), C.c_void_p(), C.c_void_p()
            lib.gx_malloc(C.byref(dcells), cells.nbytes)
            lib.gx_malloc(C.byref(dpos), m * 3 * esz)
            lib.gx_malloc(C.byref(dvel), m * 3 * esz)
            try:
                lib.gx_memcpy_h2d(dcells, cells.ctypes.data_as(C.c_void_p), cells.nbytes, None)
                for axis in range(3):
                    if psis[axis].isFourier():
                        psis[axis].toReal()
                    lib.gx_emit_cells(psis[axis].ptr, grid.size, psis[axis].dt, axis, float(grid.cellSize),
                                      float(grid.offsetLower[axis]), float(grid.periodicDomainSize), ratio, dcells, m,
                                      C.c_void_p(dpos.value + axis * esz), C.c_void_p(dvel.value + axis * esz), odt, 3,
                                      None)
                p = np.empty((m, 3), dtype=out_dtype)
                v = np.empty((m, 3), dtype=out_dtype)
                lib.gx_memcpy_d2h(p.ctypes.data_as(C.c_void_p), dpos, p.nbytes, None)
                lib.gx_memcpy_d2h(v.ctypes.data_as(C.c_void_p), dvel, v.nbytes, None)
                lib.gx_synchronize(None)
            finally:
                lib.gx_free(dcells)
                lib.gx_free(dpos)
                lib.gx_free(dvel)
            # the mapper's levels are contiguous runs in practice; a general index assignment covers every case
            if where[-1] - where[0] + 1 == m:
                self.pos[where[0]:where[0] + m] = p
                self.vel[where[0]:where[0] + m] = v
            else:
                self.pos[where] = p
                self.vel[where] = v
        self.gadgetType = np.empty(N, dtype=np.int32)
        for level, t in mapper.gadgetTypes().items():
            self.gadgetType[lev == level] = t
