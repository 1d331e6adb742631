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
component (gx_emit_cells). Gas: AddGasMapper over splitMass (gasmapper.hpp). Not covered: sub/supersampling, intermediate-resolution padding,
stray particles, the grafic mapper.
"""
import ctypes as C
import math

import numpy as np

from . import lib as L
from .host import _dt


class OneLevelParticleMapper:
    """onelevelmapper.hpp: every cell of one grid, in cell-index order."""

    def __init__(self, grid, level, gadgetParticleType=1, massScale=None):
        # massScale: the MassScaledGrid view of `grid` (virtualgrid.hpp:657-676) that splitMass creates for gas / dark matter
        self.grid, self.level, self.gadgetParticleType, self.massScale = grid, level, gadgetParticleType, massScale

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

    def layout(self):
        return (np.full(self.grid.size3, self.level, dtype=np.int32), np.arange(self.grid.size3, dtype=np.int64))

    def attrs(self):
        """Per particle: the mass scale of its (virtual) grid and its gadget type."""
        T = self.grid.dtype.type
        return (np.full(self.grid.size3, T(1) if self.massScale is None else T(self.massScale), dtype=self.grid.dtype),
                np.full(self.grid.size3, self.gadgetParticleType, dtype=np.int32))

    def size_gas(self):
        return 0

    def splitMass(self, massRatio, toLevels):
        """onelevelmapper.hpp:155-173: (gas mapper or None, dark-matter mapper)."""
        T = self.grid.dtype.type
        if self.level in toLevels:
            return (OneLevelParticleMapper(self.grid, self.level, 0, T(massRatio)),
                    OneLevelParticleMapper(self.grid, self.level, self.gadgetParticleType, T(T(1.0) - T(massRatio))))
        return None, OneLevelParticleMapper(self.grid, self.level, self.gadgetParticleType, self.massScale)

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
        self._keep = keep
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

    def layout(self):
        return self._layout

    def attrs(self):
        a1, a2 = self.pLevel1.attrs(), self.pLevel2.attrs()
        return tuple(np.concatenate([x[self._keep], y[self.zoomParticleArrayHiresUnsorted]]) for x, y in zip(a1, a2))

    def size_gas(self):
        return 0

    def splitMass(self, massRatio, toLevels):
        """twolevelmapper.hpp:519-553: gas exists only where level 2 (and possibly level 1) was asked to carry it; a level
        1 without gas is kept as the notional (skipped) low-resolution part of the gas mapper."""
        gas1, dm1 = self.pLevel1.splitMass(massRatio, toLevels)
        gas2, dm2 = self.pLevel2.splitMass(massRatio, toLevels)
        newskip = self.skipLevel1
        if gas1 is None:
            gas1, newskip = self.pLevel1, True
        gas = TwoLevelParticleMapper(gas1, gas2, self.level1CellsToReplace, newskip) if gas2 is not None else None
        return gas, TwoLevelParticleMapper(dm1, dm2, self.level1CellsToReplace, self.skipLevel1)

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


class AddGasMapper:
    """gasmapper.hpp: a gas mapper and a dark-matter mapper back to back (gas first for tipsy, last for gadget). Ids
    refer to the dark-matter part only (flagParticles / getFlaggedParticles, gasmapper.hpp:149-190)."""

    def __init__(self, pFirst, pSecond, gasFirst=True):
        self.firstMap, self.secondMap, self.gasFirst = pFirst, pSecond, gasFirst
        self.nFirst, self.nSecond = pFirst.size(), pSecond.size()
        gas, dm = (pFirst, pSecond) if gasFirst else (pSecond, pFirst)
        self.gasMap, self.dmMap = gas, dm
        if (gas.attrs()[1] != 0).any():
            raise L.GxError("AddGasMapper: the gas mapper holds particles of a gadget type other than 0")
        if (dm.attrs()[1] == 0).any():
            raise L.GxError("You have asked for some gadget DM particles to be output with type 0, but have also set "
                            "non-zero baryon density. Type 0 is reserved for gas particles in hydro simulations, and "
                            "genetIC particle types refer only to the dark matter part of the output.")

    def size(self):
        return self.nFirst + self.nSecond

    def size_gas(self):
        return self.nFirst if self.gasFirst else self.nSecond

    def size_dm(self):
        return self.nSecond if self.gasFirst else self.nFirst

    def getFinestGrid(self):
        return self.dmMap.getFinestGrid()

    def getCoarsestGrid(self):
        return self.dmMap.getCoarsestGrid()

    def grids(self):
        g = dict(self.firstMap.grids())
        g.update(self.secondMap.grids())
        return g

    def layout(self):
        return tuple(np.concatenate([a, b]) for a, b in zip(self.firstMap.layout(), self.secondMap.layout()))

    def attrs(self):
        return tuple(np.concatenate([a, b]) for a, b in zip(self.firstMap.attrs(), self.secondMap.attrs()))

    def isGas(self):
        g = np.zeros(self.size(), dtype=bool)
        if self.gasFirst:
            g[:self.nFirst] = True
        else:
            g[self.nFirst:] = True
        return g

    def unflagAllParticles(self):
        self.firstMap.unflagAllParticles()
        self.secondMap.unflagAllParticles()

    def flagParticles(self, ids):
        ids = np.asarray(ids, dtype=np.int64)
        if self.gasFirst:
            self.secondMap.flagParticles(ids[ids >= self.nFirst] - self.nFirst)
        else:
            self.firstMap.flagParticles(ids[ids < self.nFirst])

    def getFlaggedParticles(self):
        if self.gasFirst:
            return self.secondMap.getFlaggedParticles() + self.nFirst
        return self.firstMap.getFlaggedParticles()


def updateParticleMapper(context, gadgetTypesForLevels=None, cosmology=None, baryonsOnAllLevels=False, gasFirst=False):
    """ICGenerator::updateParticleMapper (ic.hpp:1129-1221) without sub/supersampling and padding: the one-level mapper of
    the base grid, wrapped once per zoom level with that level's zoomParticleArray (stored on the context by
    zoom.initZoomGrid); with OmegaBaryons0 > 0 the mass is split into a gas mapper (deepest level only unless
    baryonsOnAllLevels) and a dark-matter mapper, gas first for tipsy and last for gadget."""
    nl = context.getNumLevels()
    types = list(gadgetTypesForLevels) if gadgetTypesForLevels is not None else [1] * nl
    m = OneLevelParticleMapper(context.getGridForLevel(0), 0, types[0])
    for level in range(1, nl):
        fine = OneLevelParticleMapper(context.getGridForLevel(level), level, types[level])
        m = TwoLevelParticleMapper(m, fine, context.zoomParticleArray[level - 1])
    if cosmology is not None and cosmology.OmegaBaryons0 > 0:
        T = context.getGridForLevel(0).dtype.type
        gas, dm = m.splitMass(T(T(cosmology.OmegaBaryons0) / T(cosmology.OmegaM0)),
                              list(range(nl)) if baryonsOnAllLevels else [nl - 1])
        m = AddGasMapper(gas, dm, True) if gasFirst else AddGasMapper(dm, gas, False)
    return m


def particle_mass(grid, cosmology, massScale=None):
    """ZeldovichParticleEvaluator::getMass (zeldovich.hpp:55-62,108-110): boxMass * cellMassFrac, cellMassFrac =
    pow(dx / simsize, 3) (grid.hpp:112), times the mass scale of a MassScaledGrid (virtualgrid.hpp:672), in the grid's type.
    massScale may be an array (one entry per particle)."""
    T = grid.dtype.type
    boxMass = T(T(27.744948) * T(cosmology.OmegaM0)) * T(math.pow(float(grid.periodicDomainSize), 3.0))
    frac = T(math.pow(float(T(grid.cellSize / grid.periodicDomainSize)), 3.0))
    if massScale is None:
        return T(boxMass * frac)
    return boxMass * (np.asarray(massScale, dtype=grid.dtype) * frac)


def particle_eps(grid, epsNorm=0.01075):
    """getEps (zeldovich.hpp:113-115): cellSize * cellSofteningScale(=1) * epsNorm."""
    T = grid.dtype.type
    return T(T(grid.cellSize * T(1.0)) * T(epsNorm))


class MappedParticles:
    """All particles of a mapper, evaluated level by level on the device and placed in mapper order on the host:
    pos[N,3], vel[N,3] (out_dtype), mass[N], eps[N] (grid type), level[N], cell[N]."""

    def __init__(self, mapper, particleGenerator, cosmology, out_dtype=np.float32, epsNorm=0.01075, chunk=1 << 25):
        lib = L.load()
        lev, cell = mapper.layout()
        scale, self.gadgetType = mapper.attrs()
        N = len(lev)
        self.level, self.cell = lev, cell
        self.isGas = mapper.isGas() if hasattr(mapper, "isGas") else np.zeros(N, dtype=bool)
        self.pos = np.empty((N, 3), dtype=out_dtype)
        self.vel = np.empty((N, 3), dtype=out_dtype)
        grids = mapper.grids()
        T = next(iter(grids.values())).dtype.type
        self.mass = np.empty(N, dtype=T)
        self.eps = np.empty(N, dtype=T)
        odt, esz = _dt(out_dtype), np.dtype(out_dtype).itemsize
        for level, grid in grids.items():
            where = np.flatnonzero(lev == level)
            self.mass[where] = particle_mass(grid, cosmology, scale[where])
            self.eps[where] = particle_eps(grid, epsNorm)
            if len(where) == 0:
                continue
            psis = particleGenerator.getGeneratorForLevel(level).getGeneratedFields()
            for axis in range(3):
                if psis[axis].isFourier():
                    psis[axis].toReal()
            ratio = cosmology.zeldovichVelocityToOffsetRatio(grid.dtype)
            contiguous_out = where[-1] - where[0] + 1 == len(where)     # the mapper's levels are runs in practice
            for c0 in range(0, len(where), chunk):                       # bounded device staging
                w = where[c0:c0 + chunk]
                m = len(w)
                cells = np.ascontiguousarray(cell[w])
                is_range = int(cells[-1]) - int(cells[0]) + 1 == m and (m == 1 or bool(np.all(np.diff(cells) == 1)))
                dcells, dpos, dvel = C.c_void_p(), C.c_void_p(), C.c_void_p()
                lib.gx_malloc(C.byref(dpos), m * 3 * esz)
                lib.gx_malloc(C.byref(dvel), m * 3 * esz)
                try:
                    if not is_range:
                        lib.gx_malloc(C.byref(dcells), cells.nbytes)
                        lib.gx_memcpy_h2d(dcells, cells.ctypes.data_as(C.c_void_p), cells.nbytes, None)
                    for axis in range(3):
                        geo = (psis[axis].ptr, grid.size, psis[axis].dt, axis, float(grid.cellSize),
                               float(grid.offsetLower[axis]), float(grid.periodicDomainSize), ratio)
                        out = (C.c_void_p(dpos.value + axis * esz), C.c_void_p(dvel.value + axis * esz), odt, 3)
                        if is_range:     # a whole level, or the unzoomed stretch of one: no cell list needed
                            lib.gx_emit_component(*geo, *out, int(cells[0]), m, None)
                        else:
                            lib.gx_emit_cells(*geo, dcells, m, *out, None)
                    if contiguous_out:
                        p, v = self.pos[w[0]:w[0] + m], self.vel[w[0]:w[0] + m]
                    else:
                        p, v = np.empty((m, 3), dtype=out_dtype), np.empty((m, 3), dtype=out_dtype)
                    lib.gx_memcpy_d2h(p.ctypes.data_as(C.c_void_p), dpos, p.nbytes, None)
                    lib.gx_memcpy_d2h(v.ctypes.data_as(C.c_void_p), dvel, v.nbytes, None)
                    lib.gx_synchronize(None)
                    if not contiguous_out:
                        self.pos[w], self.vel[w] = p, v
                finally:
                    if dcells.value:
                        lib.gx_free(dcells)
                    lib.gx_free(dpos)
                    lib.gx_free(dvel)
