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
component (gx_emit_cells). Gas: AddGasMapper over splitMass (gasmapper.hpp). Sub/supersampling, intermediate-resolution
padding (autopad), the flagged particle type and stray particles (ResolutionMatchingGrid output grids, nested for a second
zoom level) are here as well; the grafic mapper lives in grafic.py. tools/fuzz_vs_reference.py compares the written files
with the reference binary's on random paramfiles.
"""
import ctypes as C
import itertools
import math

import numpy as np

from . import lib as L
from .host import _dt, Field, scale_cells


_layer_ids = itertools.count()


class OneLevelParticleMapper:
    """onelevelmapper.hpp: every cell of one grid, in cell-index order. `grid` may be a virtual view (host.VirtualGrid:
    super/subsampled or mass-scaled) of the stored grid of `level`; every instance is one *layer* of a layout."""

    def __init__(self, grid, level, gadgetParticleType=1):
        self.grid, self.level, self.gadgetParticleType = grid, level, gadgetParticleType
        self.layer = next(_layer_ids)

    def size(self):
        return self.grid.size3

    def getFinestGrid(self):
        return self.grid

    def getCoarsestGrid(self):
        return self.grid

    def getFinestLayer(self):
        return self.layer

    def layers(self):
        return {self.layer: self}

    def grids(self):
        return {self.layer: self.grid}

    def layout(self):
        return (np.full(self.grid.size3, self.layer, dtype=np.int32), np.arange(self.grid.size3, dtype=np.int64))

    def attrs(self):
        """Per particle: its gadget type."""
        return (np.full(self.grid.size3, self.gadgetParticleType, dtype=np.int32),)

    def size_gas(self):
        return 0

    def splitMass(self, massRatio, toGrids):
        """onelevelmapper.hpp:155-173: (gas mapper or None, dark-matter mapper) on mass-scaled views of the grid."""
        T = self.grid.dtype.type
        if any(self.grid.base is g for g in toGrids):
            return (OneLevelParticleMapper(self.grid.makeScaledMassVersion(T(massRatio)), self.level, 0),
                    OneLevelParticleMapper(self.grid.makeScaledMassVersion(T(T(1.0) - T(massRatio))), self.level,
                                           self.gadgetParticleType))
        return None, OneLevelParticleMapper(self.grid, self.level, self.gadgetParticleType)

    def insertIntermediateResolutionPadding(self, ratio, padCells):
        return self

    def getGadgetParticleTypeForFinestGrid(self):
        return self.gadgetParticleType

    def superOrSubSample(self, ratio, toGrids, super_):
        """onelevelmapper.hpp:179-195."""
        g = self.grid
        if any(g.base is t for t in toGrids):
            g = g.makeSupersampled(ratio) if super_ else g.makeSubsampled(ratio)
        return OneLevelParticleMapper(g, self.level, self.gadgetParticleType)

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
        on_finest = np.flatnonzero(lev == pLevel1.getFinestLayer())
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
        self._layout = (np.concatenate([lev[keep], np.full(len(self.zoomParticleArrayHiresUnsorted), pLevel2.layer, np.int32)]),
                        np.concatenate([cell[keep], self.zoomParticleArrayHiresUnsorted]))

    def size(self):
        return self.totalParticles

    def getFinestGrid(self):
        return self.pLevel2.getFinestGrid()

    def getCoarsestGrid(self):
        return self.pLevel1.getCoarsestGrid()

    def getFinestLayer(self):
        return self.pLevel2.layer

    def layers(self):
        g = dict(self.pLevel1.layers())
        g.update(self.pLevel2.layers())
        return g

    def grids(self):
        return {k: m.grid for k, m in self.layers().items()}

    def layout(self):
        return self._layout

    def attrs(self):
        a1, a2 = self.pLevel1.attrs(), self.pLevel2.attrs()
        return tuple(np.concatenate([x[self._keep], y[self.zoomParticleArrayHiresUnsorted]]) for x, y in zip(a1, a2))

    def size_gas(self):
        return 0

    def getGadgetParticleTypeForFinestGrid(self):
        return self.pLevel2.getGadgetParticleTypeForFinestGrid()

    def insertIntermediateResolutionPadding(self, ratio, padCells):
        """twolevelmapper.hpp:603-690 (`autopad`): where the zoom is more than `ratio` times finer than its surroundings, a
        shell of padCells cells around the zoomed cells is replaced by particles of intermediate resolution interpolated from
        the level above, recursively until the jump is at most `ratio`."""
        mapperLR = self.pLevel1.insertIntermediateResolutionPadding(ratio, padCells)
        cells = self.level1CellsToReplace
        finestLR = mapperLR.getFinestGrid()
        r = float(finestLR.cellSize) / float(self.pGrid2.cellSize)
        if abs(r - round(r)) > 1e-8:
            raise L.GxError("autopad: the cell size ratio is not an integer")
        newRatio = int(round(r))
        if newRatio <= ratio:
            return TwoLevelParticleMapper(mapperLR, self.pLevel2, cells, self.skipLevel1)
        if newRatio % ratio:
            raise L.GxError("autopad: the zoom factor is not a multiple of the padding ratio")
        supersampledCells = scale_cells(cells, finestLR.size, finestLR.size * ratio)
        expanded = _expandFlaggedRegion(cells, finestLR.size, padCells)
        if isinstance(mapperLR, TwoLevelParticleMapper):       # only cells the level-1 mapper can reach stay flagged
            expanded = np.intersect1d(expanded, mapperLR.zoomParticleArrayHiresUnsorted)
        # The reference builds a ResolutionMatchingGrid(zoom grid subsampled, finestLR) here (twolevelmapper.hpp:655-657), but
        # its compiled binary evaluates every padding particle from the level above (tricubic interpolation, mass of a
        # level-above cell / ratio^3, also where the zoom level's fields exist): measured against oracle/_ref/genetIC and
        # the goldens of test_10i / test_18, which is what a super-sampled view of finestLR gives.
        padGrid = finestLR.makeSupersampled(ratio)
        padMapper = OneLevelParticleMapper(padGrid, self.pLevel2.level, mapperLR.getGadgetParticleTypeForFinestGrid())
        lowAndPad = TwoLevelParticleMapper(mapperLR, padMapper, expanded, self.skipLevel1)
        everything = TwoLevelParticleMapper(lowAndPad, self.pLevel2, supersampledCells, self.skipLevel1)
        return everything.insertIntermediateResolutionPadding(ratio, padCells)

    def superOrSubSample(self, ratio, toGrids, super_):
        """twolevelmapper.hpp:568-591: both sub-mappers are resampled and the cells to replace are re-expressed on the new
        finest level-1 grid through the flags (virtual grids scale them), which leaves exactly those cells flagged."""
        subLR = self.pLevel1.superOrSubSample(ratio, toGrids, super_)
        subHR = self.pLevel2.superOrSubSample(ratio, toGrids, super_)
        subLR.unflagAllParticles()
        self.pLevel1.flagParticles(self.level1ParticlesToReplace)
        newCells = subLR.getFinestGrid().getFlaggedCells()
        try:
            return TwoLevelParticleMapper(subLR, subHR, newCells, self.skipLevel1)
        except L.GxError as e:
            if "outside high resolution grid" in str(e):
                raise L.GxError("Following sub/super-sampling, the zoom particles no longer fit fully on their grids. Try "
                                "rerunning with strays_on, or (better) increase the size of the zoom region.") from e
            raise

    def splitMass(self, massRatio, toGrids):
        """twolevelmapper.hpp:519-553: gas exists only where level 2 (and possibly level 1) was asked to carry it; a level
        1 without gas is kept as the notional (skipped) low-resolution part of the gas mapper."""
        gas1, dm1 = self.pLevel1.splitMass(massRatio, toGrids)
        gas2, dm2 = self.pLevel2.splitMass(massRatio, toGrids)
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
        if (gas.attrs()[0] != 0).any():
            raise L.GxError("AddGasMapper: the gas mapper holds particles of a gadget type other than 0")
        if (dm.attrs()[0] == 0).any():
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

    def layers(self):
        g = dict(self.firstMap.layers())
        g.update(self.secondMap.layers())
        return g

    def grids(self):
        return {k: m.grid for k, m in self.layers().items()}

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

    def insertIntermediateResolutionPadding(self, ratio, padCells):
        """gasmapper.hpp:299-310: padding goes around the dark-matter map only."""
        if self.gasFirst:
            return AddGasMapper(self.firstMap, self.secondMap.insertIntermediateResolutionPadding(ratio, padCells), True)
        return AddGasMapper(self.firstMap.insertIntermediateResolutionPadding(ratio, padCells), self.secondMap, False)


def _expandFlaggedRegion(flags, n, nCells):
    """Grid::expandFlaggedRegion (grid.hpp:277-306): the +-1 neighbours along each axis, nCells times (periodic)."""
    flags = np.asarray(flags, dtype=np.int64)
    for _ in range(nCells):
        for axis in range(3):
            c = np.stack(np.unravel_index(flags, (n, n, n)), axis=0)
            out = [flags]
            for step in (1, -1):
                d = c.copy()
                d[axis] = (d[axis] + step) % n
                out.append(np.ravel_multi_index(tuple(d), (n, n, n)))
            flags = np.unique(np.concatenate(out))
    return flags


def references(mapper, grid):
    """ParticleMapper::references: does the mapper use (a view of) this stored grid?"""
    if hasattr(mapper, "references"):
        return mapper.references(grid)
    return any(m.grid.base is grid for m in mapper.layers().values())


def extendParticleListToUnreferencedGrids(mapper, context):
    """mapper.hpp:183-196: every grid of `context` the mapper does not use receives the flags of the mapper's finest
    grid, re-expressed on it (zoom.proxyFlags)."""
    from . import zoom as Z
    for level in range(context.getNumLevels()):
        g = context.getGridForLevel(level)
        if not references(mapper, g):
            g.flagCells(Z.proxyFlags(mapper.getFinestGrid(), g), append=True)


def updateParticleMapper(context, gadgetTypesForLevels=None, cosmology=None, baryonsOnAllLevels=False, gasFirst=False,
                         supersample=1, subsample=1, supersampleGas=1, flaggedGadgetParticleType=None, flaggedCells=None,
                         autopad=0):
    """ICGenerator::updateParticleMapper (ic.hpp:1129-1236) without padding and strays: the one-level mapper of the base
    grid, wrapped once per zoom level with that level's zoomParticleArray (stored on the context by zoom.initZoomGrid);
    with OmegaBaryons0 > 0 the mass is split into a gas mapper (deepest level only unless baryonsOnAllLevels) and a
    dark-matter mapper, gas first for tipsy and last for gadget; `supersample` refines the deepest level, `subsample`
    coarsens the base level (superOrSubSample, ic.hpp:1223-1234)."""
    nl = context.getNumLevels()
    types = list(gadgetTypesForLevels) if gadgetTypesForLevels is not None else [1] * nl
    m = OneLevelParticleMapper(context.getOutputGridForLevel(0), 0, types[0])
    for level in range(1, nl):
        fine = OneLevelParticleMapper(context.getOutputGridForLevel(level), level, types[level])
        m = TwoLevelParticleMapper(m, fine, context.zoomParticleArray[level - 1])
    if flaggedGadgetParticleType is not None:       # ic.hpp:1176-1186: the flagged cells of the finest grid, retyped
        flagged = OneLevelParticleMapper(context.getOutputGridForLevel(nl - 1), nl - 1, flaggedGadgetParticleType)
        m = TwoLevelParticleMapper(m, flagged, flaggedCells)

    def resample(mp, sup, sub):
        if sup > 1:
            mp = mp.superOrSubSample(sup, [context.getGridForLevel(nl - 1)], True)
        if sub > 1:
            mp = mp.superOrSubSample(sub, [context.getGridForLevel(0)], False)
        return mp

    if cosmology is not None and cosmology.OmegaBaryons0 > 0:
        T = context.getGridForLevel(0).dtype.type
        gas, dm = m.splitMass(T(T(cosmology.OmegaBaryons0) / T(cosmology.OmegaM0)),
                              list(context.grids) if baryonsOnAllLevels else [context.getGridForLevel(nl - 1)])
        gas, dm = resample(gas, supersampleGas, 1), resample(dm, supersample, subsample)
        m = AddGasMapper(gas, dm, True) if gasFirst else AddGasMapper(dm, gas, False)
    else:
        m = resample(m, supersample, subsample)
    if autopad > 0:
        m = m.insertIntermediateResolutionPadding(2, autopad)      # ic.hpp:1216-1218
    return m


def particle_mass(grid, cosmology):
    """ZeldovichParticleEvaluator::getMass (zeldovich.hpp:55-62,108-110): boxMass * cellMassFrac; cellMassFrac =
    pow(dx / simsize, 3) for a stored grid (grid.hpp:112), rescaled by the virtual views (virtualgrid.hpp:145,544,672)."""
    T = grid.dtype.type
    boxMass = T(T(27.744948) * T(cosmology.OmegaM0)) * T(math.pow(float(grid.periodicDomainSize), 3.0))
    return T(boxMass * grid.cellMassFrac)


def particle_eps(grid, epsNorm=0.01075):
    """getEps (zeldovich.hpp:113-115): cellSize * cellSofteningScale * epsNorm."""
    T = grid.dtype.type
    return T(T(grid.cellSize * grid.cellSofteningScale) * T(epsNorm))


class MappedParticles:
    """All particles of a mapper, evaluated level by level on the device and placed in mapper order on the host:
    pos[N,3], vel[N,3] (out_dtype), mass[N], eps[N] (grid type), level[N], cell[N]."""

    def __init__(self, mapper, particleGenerator, cosmology, out_dtype=np.float32, epsNorm=0.01075, chunk=1 << 25):
        lib = L.load()
        lev, cell = mapper.layout()
        (self.gadgetType,) = mapper.attrs()
        N = len(lev)
        self.layer, self.cell = lev, cell
        self.level = np.empty(N, dtype=np.int32)             # the stored level each particle's fields come from
        self.isGas = mapper.isGas() if hasattr(mapper, "isGas") else np.zeros(N, dtype=bool)
        self.pos = np.empty((N, 3), dtype=out_dtype)
        self.vel = np.empty((N, 3), dtype=out_dtype)
        layers = mapper.layers()
        T = next(iter(layers.values())).grid.dtype.type
        self.mass = np.empty(N, dtype=T)
        self.eps = np.empty(N, dtype=T)
        odt, esz = _dt(out_dtype), np.dtype(out_dtype).itemsize
        self._lib, self._gen, self._cos, self._chunk = lib, particleGenerator, cosmology, chunk
        self._odt, self._esz, self._out_dtype, self._resampled = odt, esz, out_dtype, {}
        for layer, onelevel in layers.items():
            grid = onelevel.grid
            where = np.flatnonzero(lev == layer)
            self.mass[where] = particle_mass(grid, cosmology)
            self.eps[where] = particle_eps(grid, epsNorm)
            self.level[where] = onelevel.level
            if len(where):
                self._evaluate(grid, cell[where], where)
        del self._resampled

    def _evaluate(self, grid, cells, where):
        """Particles of `cells` of `grid` into rows `where` of pos / vel."""
        if len(cells) == 0:
            return
        inner = grid
        while getattr(inner, "kind", None) == "mass":      # MassScaledGrid: same cells, same fields (evaluator.hpp:311-315)
            inner = inner.underlying
        if hasattr(inner, "loInterp"):
            # ResolutionMatchingEvaluator (evaluator.hpp:218-256): cells in the high-resolution window come from the
            # zoom level's fields, stray cells from the level above interpolated at the fine cell centres
            m, hi_cells = inner.split(cells)
            self._evaluate(inner.hi, hi_cells, where[m])
            self._evaluate(inner.loInterp, cells[~m], where[~m])
            return
        lib, cosmology, chunk, odt, esz, out_dtype = self._lib, self._cos, self._chunk, self._odt, self._esz, self._out_dtype
        level = grid.base.levelIndex
        psis = self._gen.getGeneratorForLevel(level).getGeneratedFields()
        for axis in range(3):
            if psis[axis].isFourier():
                psis[axis].toReal()
        res = grid.resolution() if hasattr(grid, "resolution") else (1, 1)
        if res != (1, 1):
            # Super/SubSampleEvaluator (evaluator.hpp:96-205): the stored offsets interpolated tricubically at the
            # centres of a finer virtual grid, or averaged over the sub-cells of a coarser one -- the same cross-grid
            # kernels that combine the levels, writing a field on the virtual grid
            key = (level,) + res
            if key not in self._resampled:
                out = []
                for axis in range(3):
                    v = Field(grid, fourier=False, stream=psis[axis].stream)
                    v.addFieldFromDifferentGrid(psis[axis])
                    out.append(v)
                self._resampled[key] = out
            psis = self._resampled[key]
        ratio = cosmology.zeldovichVelocityToOffsetRatio(grid.dtype)
        contiguous_out = where[-1] - where[0] + 1 == len(where) and (len(where) == 1 or bool(np.all(np.diff(where) == 1)))
        for c0 in range(0, len(where), chunk):                       # bounded device staging
            w = where[c0:c0 + chunk]
            m = len(w)
            cc = np.ascontiguousarray(cells[c0:c0 + chunk])
            is_range = int(cc[-1]) - int(cc[0]) + 1 == m and (m == 1 or bool(np.all(np.diff(cc) == 1)))
            dcells, dpos, dvel = C.c_void_p(), C.c_void_p(), C.c_void_p()
            lib.gx_malloc(C.byref(dpos), m * 3 * esz)
            lib.gx_malloc(C.byref(dvel), m * 3 * esz)
            try:
                if not is_range:
                    lib.gx_malloc(C.byref(dcells), cc.nbytes)
                    lib.gx_memcpy_h2d(dcells, cc.ctypes.data_as(C.c_void_p), cc.nbytes, None)
                for axis in range(3):
                    geo = (psis[axis].ptr, grid.size, psis[axis].dt, axis, float(grid.cellSize),
                           float(grid.offsetLower[axis]), float(grid.periodicDomainSize), ratio)
                    out = (C.c_void_p(dpos.value + axis * esz), C.c_void_p(dvel.value + axis * esz), odt, 3)
                    if is_range:     # a whole level, or the unzoomed stretch of one: no cell list needed
                        lib.gx_emit_component(*geo, *out, int(cc[0]), m, None)
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

def field_on_mapper(mapper, fieldsOnLevels):
    """MapperIterator::getField (mapperiterator.hpp:188-197) for every particle: the value of a multi-level real-space
    field in the cell each particle comes from; virtual layers are resampled on the device like the particle offsets
    (makeEvaluator, evaluator.hpp:271-318). Returns a host array in mapper order (diagnostic output, `dump_tipsy`)."""
    lev, cell = mapper.layout()
    out = np.empty(len(lev), dtype=fieldsOnLevels[0].grid.dtype)
    for layer, onelevel in mapper.layers().items():
        w = np.flatnonzero(lev == layer)
        if len(w):
            out[w] = _grid_values(onelevel.grid, fieldsOnLevels, cell[w])
    return out


def _grid_values(grid, fieldsOnLevels, cells):
    while getattr(grid, "kind", None) == "mass":
        grid = grid.underlying
    if hasattr(grid, "loInterp"):                      # ResolutionMatchingEvaluator
        m, hi_cells = grid.split(cells)
        out = np.empty(len(cells), dtype=grid.dtype)
        out[m] = _grid_values(grid.hi, fieldsOnLevels, hi_cells)
        out[~m] = _grid_values(grid.loInterp, fieldsOnLevels, np.asarray(cells)[~m])
        return out
    src = fieldsOnLevels[grid.base.levelIndex]
    if src.isFourier():
        src.toReal()
    res = grid.resolution() if hasattr(grid, "resolution") else (1, 1)
    if res != (1, 1):
        v = Field(grid, fourier=False, stream=src.stream)
        v.addFieldFromDifferentGrid(src)
        src = v
    return src.getDataVector().reshape(-1)[cells]
