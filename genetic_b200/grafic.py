"""Grafic (RAMSES) output, its particle mapper and its refinement masks -- host side of `SURVEY.md` §8(f) N2 / N4.

The reference writes every level of a *centred intermediate* copy of the multi-level context as a directory of
Fortran-record files (io/grafic.hpp:36-270): the real grids plus virtual ones, so that the stack goes up in factors of two
(multilevelgrid.hpp:337-427: `subsample` adds coarser copies of the base grid, zoom factors > 2 add coarser copies of the zoom
grid, `supersample` adds finer copies of the finest grid), every grid that covers the box re-centred on a point
(CenteredGrid, virtualgrid.hpp:688-800: same cells, visited with an integer shift) and every zoom grid moved by the same
physical offset (OffsetGrid, :806-830). Here a level of that stack is a `_GLevel`: the sub/supersampled view of a stored grid
(`host.VirtualGrid`, evaluated on the device by the cross-grid kernels), the integer shift, and the header offset.

  GraficContext   copyContextWithCenteredIntermediate
  GraficMapper    particles/mapper/graficmapper.hpp: ids = cells of the levels in turn; flags propagate to the coarser real levels
  GraficMask      multilevelgrid/mask.hpp: the zoom regions carried through the virtual levels (ic_refmap / ic_pvar, `dump_mask`)
  save            io/grafic.hpp: ic_velc[xyz], ic_posc[xyz], ic_deltab, ic_refmap, ic_pvar_00001, ic_particle_ids per level
"""
import math
import os
import struct

import numpy as np

from . import lib as L
from .host import Field


def _find_power_of(base, p):
    """tools::findPowerOf (util_functions.hpp:52-54): log(p) / log(base), which must be an integer."""
    r = math.log(p) / math.log(base)
    if abs(r - round(r)) > 1e-8:
        raise L.GxError("Subsample and supersample cannot be converted into an integer number of grids with this base factor")
    return int(round(r))


class _GLevel:
    """One grid of the centred stack. `grid` is the stored grid or a sub/supersampled view of it; cell indices are those of
    `grid` (a CenteredGrid keeps the indices of the grid it wraps and only visits them in another order)."""

    def __init__(self, grid, resampled):
        self.grid, self.resampled = grid, resampled
        self.shift = None                        # CenteredGrid: integer offset per axis; None for an OffsetGrid
        self.pointOffset = None                  # OffsetGrid: the physical offset added to the lower corner
        self.offsetLower = tuple(grid.offsetLower)

    # -- geometry with the reference's arithmetic (grid.hpp:519-531, 656-677; virtualgrid.hpp:717-790) -----------------
    def _wrap(self, x):
        Ls = self.grid.periodicDomainSize
        x = np.fmod(x, Ls)
        return np.where(x < 0, x + Ls, x)

    def centroids(self, ids):
        g = self.grid
        T = g.dtype.type
        n = g.size
        ids = np.asarray(ids, dtype=np.int64)
        coords = np.stack(np.unravel_index(ids, (n, n, n)), axis=1).astype(T)
        if self.shift is not None:               # CenteredGrid: wrapPoint(underlying centroid + point offset)
            c = coords * g.cellSize + np.array(g.offsetLower, dtype=T) + T(g.cellSize / T(2))
            return self._wrap(c + np.array([T(s) * g.cellSize for s in self.shift], dtype=T))
        return coords * g.cellSize + np.array(self.offsetLower, dtype=T) + T(g.cellSize / T(2))

    def indexFromPoints(self, pts):
        g = self.grid
        T = g.dtype.type
        pts = np.asarray(pts, dtype=T)
        if self.shift is not None:               # CenteredGrid: underlying.getIndexFromPoint(point + inverse point offset)
            pts = pts + np.array([T(-s) * g.cellSize for s in self.shift], dtype=T)
            lower = np.array(g.offsetLower, dtype=T)
        else:
            lower = np.array(self.offsetLower, dtype=T)
        c = np.floor(self._wrap(pts - lower) / g.cellSize).astype(np.int64)
        return (c[:, 0] * g.size + c[:, 1]) * g.size + c[:, 2]

    def toOutputOrder(self, values):
        """values[index of `grid`] as an [x][y][z] cube -> the cube in the coordinates of the centred grid:
        out[x, y, z] = values[getIndexFromCoordinateNoWrap(x, y, z)] (virtualgrid.hpp:746-761)."""
        n = self.grid.size
        cube = np.asarray(values).reshape(n, n, n)
        if self.shift is not None and any(self.shift):
            cube = np.roll(cube, self.shift, axis=(0, 1, 2))
        return cube

    def cellIndexCube(self):
        n = self.grid.size
        return self.toOutputOrder(np.arange(n ** 3, dtype=np.int64))


class GraficContext:
    """MultiLevelGrid::copyContextWithCenteredIntermediate (multilevelgrid.hpp:337-473) with base factor 2."""

    def __init__(self, context, centre, subsample=1, supersample=1, log=None):
        log = log or (lambda s: None)
        extralow, extrahigh = _find_power_of(2, subsample), _find_power_of(2, supersample)
        real = [context.getGridForLevel(l) for l in range(context.getNumLevels())]
        for l, g in enumerate(real):
            g.levelIndex = l
        levels = []

        def add_sub(level, count):
            factor = 2 ** count
            neff = int(round(float(real[0].cellSize) / float(real[level].cellSize))) * real[0].size
            for _ in range(count):
                log("Adding virtual grid with effective resolution %d" % (neff // factor))
                levels.append(_GLevel(real[level].makeSubsampled(factor), True))
                factor //= 2

        for level, g in enumerate(real):
            if level == 0:
                add_sub(0, extralow)
                log("Adding real grid with resolution %d" % g.size)
            else:
                zoomfactor = float(real[level - 1].cellSize) / float(g.cellSize)
                add_sub(level, int(round(math.log10(zoomfactor) / math.log10(2.0)) - 1))
                log("Adding real grid with resolution %d" % (int(round(float(real[0].cellSize) / float(g.cellSize))) * real[0].size))
            levels.append(_GLevel(g, False))
        factor = 2
        neff = int(round(float(real[0].cellSize) / float(real[-1].cellSize))) * real[0].size
        for _ in range(extrahigh):
            log("Adding virtual grid with effective resolution %d" % (neff * factor))
            levels.append(_GLevel(real[-1].makeSupersampled(factor), True))
            factor *= 2
        # copyContextAndCenter: CenteredGrid for every level that covers the box, OffsetGrid below
        deepest = max(i for i, lv in enumerate(levels) if lv.grid.periodicDomainSize == lv.grid.thisGridSize)
        offset = None
        for i, lv in enumerate(levels):
            g = lv.grid
            T = g.dtype.type
            if i <= deepest:
                # Coordinate<int>((0.5 * thisGridSize - centre) / cellSize): truncation towards zero
                lv.shift = tuple(int(T(T(T(0.5) * g.thisGridSize) - T(centre[a])) / g.cellSize) for a in range(3))
                offset = tuple(T(s) * g.cellSize for s in lv.shift)
            else:
                lv.pointOffset = offset
                Ls = float(g.periodicDomainSize)
                lo = []
                for a in range(3):
                    x = math.fmod(float(T(g.offsetLower[a] + offset[a])), Ls)
                    lo.append(T(x + Ls if x < 0 else x))
                lv.offsetLower = tuple(lo)
        self.levels = levels

    def getNumLevels(self):
        return len(self.levels)

    def getIndexOfCellOnOtherLevel(self, currentLevel, otherLevel, ids):
        return self.levels[otherLevel].indexFromPoints(self.levels[currentLevel].centroids(ids))


class GraficMapper:
    """particle::mapper::GraficMapper (graficmapper.hpp:20-200): particle ids are the cells of the levels of the centred stack
    in turn, coarsest first."""

    def __init__(self, context, centre, subsample=1, supersample=1, log=None):
        self.log = log or (lambda s: None)
        self.ctx = GraficContext(context, centre, subsample, supersample)

    def size(self):
        return sum(lv.grid.size3 for lv in self.ctx.levels)

    def size_gas(self):
        return 0

    def unflagAllParticles(self):
        for lv in self.ctx.levels:
            lv.grid.unflagAllCells()

    def flagParticles(self, ids):
        ids = np.asarray(ids, dtype=np.int64)
        start = 0
        used = 0
        for lv in self.ctx.levels:
            end = start + lv.grid.size3
            m = (ids >= start) & (ids < end)
            cells = ids[m] - start
            used += int(m.sum())
            if lv.resampled and len(cells):
                self.log("%d input ids reference a GRAFIC intermediate grid - make sure you intended to do this !" % len(cells))
            lv.grid.flagCells(cells, append=True)
            start = end
        if used != len(ids):
            raise L.GxError("Ran out of grids when interpreting grafic cell IDs - check IDs?")
        self._propagateFlagsThroughHierarchy()

    def getFlaggedParticles(self):
        out, offset = [], 0
        for lv in self.ctx.levels:
            out.append(np.asarray(lv.grid.getFlaggedCells(), dtype=np.int64) + offset)
            offset += lv.grid.size3
        return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)

    def _propagateFlagsThroughHierarchy(self):
        real = [i for i, lv in enumerate(self.ctx.levels) if not lv.resampled]
        for k in range(len(real) - 1, 0, -1):
            this, coarser = real[k], real[k - 1]
            flags = self.ctx.levels[this].grid.getFlaggedCells()
            if len(flags) == 0:
                self.ctx.levels[coarser].grid.flagCells([], append=True)
                continue
            up = np.unique(self.ctx.getIndexOfCellOnOtherLevel(this, coarser, flags))
            self.ctx.levels[coarser].grid.flagCells(up, append=True)

    def references(self, grid):
        return any(lv.grid is grid or getattr(lv.grid, "base", None) is grid for lv in self.ctx.levels)

    def getCoarsestGrid(self):
        return self.ctx.levels[0].grid

    def getFinestGrid(self):
        return self.ctx.levels[-1].grid


class GraficMask:
    """multilevelgrid::GraficMask (mask.hpp:45-228): the cells replaced by each zoom (`zoomParticleArray`) carried through the
    whole centred stack, virtual levels included."""

    def __init__(self, gctx, inputMasks, log=None):
        self.ctx, self.input = gctx, [np.asarray(m, dtype=np.int64) for m in inputMasks]
        self.flagged = [np.zeros(0, dtype=np.int64) for _ in gctx.levels]
        self.deepest = -1
        self.log = log or (lambda s: None)

    def calculateMask(self):
        self._identifyLevelsOfInputMask()
        if self.deepest < 0:
            self.log("WARNING No zoom regions were ever opened. Grafic mask will not be generated in this case")
            return
        nl = self.ctx.getNumLevels()
        for level in range(self.deepest, -1, -1):                  # generateHierarchyAboveLevelInclusive
            if len(self.flagged[level]) == 0:
                self.flagged[level] = np.unique(self.ctx.getIndexOfCellOnOtherLevel(level + 1, level, self.flagged[level + 1]))
        for level in range(self.deepest + 1, nl - 1):              # generateHierarchyBelowLevelExclusive
            n3 = self.ctx.levels[level].grid.size3
            ids = np.arange(n3, dtype=np.int64)
            above = self.ctx.getIndexOfCellOnOtherLevel(level, level - 1, ids)
            self.flagged[level] = ids[np.isin(above, self.flagged[level - 1])]

    def _identifyLevelsOfInputMask(self):
        # a level takes the next input mask as soon as one of its cell indices occurs in it (mask.hpp:93-116): the masks are
        # matched by index range only, exactly as the reference does
        k = 0
        for level in range(self.ctx.getNumLevels() - 1):
            if k >= len(self.input):
                return
            if np.any(self.input[k] < self.ctx.levels[level].grid.size3):
                self.flagged[level] = self.input[k]
                self.deepest = level
                k += 1

    def isInMask(self, level):
        """1 / 0 per cell index of the level's grid (everything is refined on a level without flags)."""
        n3 = self.ctx.levels[level].grid.size3
        if len(self.flagged[level]) == 0:
            return np.ones(n3, dtype=np.float32)
        m = np.zeros(n3, dtype=np.float32)
        f = self.flagged[level]
        m[f[f < n3]] = 1.0       # a mask matched to a level by index range may hold indices beyond it (never looked up)
        return m

    def cube(self, level):
        """convertToField (mask.hpp:196-226): the mask in the coordinates of the centred grid."""
        return self.ctx.levels[level].toOutputOrder(self.isInMask(level))


# ---- writer ---------------------------------------------------------------------------------------------------------------
def _fortran_planes(f, cube_zyx, dtype):
    """MemMapFileWriter::getMemMapFortran per z-plane: int32 byte count, the plane, the count again (memmap.hpp)."""
    n = cube_zyx.shape[0]
    marker = struct.pack("<i", n * n * np.dtype(dtype).itemsize)
    for z in range(n):
        f.write(marker)
        f.write(np.ascontiguousarray(cube_zyx[z], dtype=dtype).tobytes())
        f.write(marker)


def _values_on_level(lv, src):
    """A stored real-space field on the cells of a (possibly sub/supersampled) level: Direct / SubSample / SuperSample
    evaluators (evaluator.hpp:59-205, 271-318) through the cross-grid kernels on the device."""
    if src.isFourier():
        src.toReal()
    g = lv.grid
    if lv.resampled:
        v = Field(g, fourier=False, stream=src.stream)
        v.addFieldFromDifferentGrid(src)
        src = v
    return src.getDataVector().reshape(-1)


def save(filename, context, particleGenerator, overdensityField, baryonField, cosmology, pvarValue, centre, subsample,
         supersample, inputMasks, log=None):
    """io::grafic::save (grafic.hpp:36-270). particleGenerator: the multi-level Zeldovich generator (offset fields per stored
    level); overdensityField: the dark-matter output field; baryonField: the baryon output field or None (ic_deltab then
    comes from the dark-matter field, as the reference's generators[baryon] does when there are no baryons)."""
    gctx = GraficContext(context, centre, subsample, supersample, log)
    mask = GraficMask(gctx, inputMasks, log)
    mask.calculateMask()
    lengthFactorHeader = 1.0 / float(cosmology.hubble)
    velFactor = math.pow(float(cosmology.scalefactor), 0.5)
    base = gctx.levels[0].grid
    iordOffset = 0
    names = ["ic_velcx", "ic_velcy", "ic_velcz", "ic_poscx", "ic_poscy", "ic_poscz", "ic_deltab", "ic_refmap", "ic_pvar_00001",
             "ic_particle_ids"]
    written = []
    for level, lv in enumerate(gctx.levels):
        g = lv.grid
        T = g.dtype.type
        n = g.size
        effective = int(round(float(base.cellSize * T(base.size)) / float(g.cellSize)))
        d = "%s_%d" % (filename, effective)
        os.makedirs(d, exist_ok=True)
        stored = g.base.levelIndex
        ratio = T(cosmology.zeldovichVelocityToOffsetRatio(g.dtype))
        psis = particleGenerator.getGeneratorForLevel(stored).getGeneratedFields()
        disp = [_values_on_level(lv, psis[a]) for a in range(3)]
        dsrc = (baryonField if baryonField is not None else overdensityField).getFieldForLevel(stored)
        deltab = _values_on_level(lv, dsrc)
        m = mask.isInMask(level)
        header = struct.pack("<iiiffffffff", n, n, n, np.float32(float(g.cellSize) * lengthFactorHeader),
                             *[np.float32(float(lv.offsetLower[a]) * lengthFactorHeader) for a in range(3)],
                             np.float32(cosmology.scalefactor), np.float32(cosmology.OmegaM0),
                             np.float32(cosmology.OmegaLambda0), np.float32(float(cosmology.hubble) * 100))
        assert len(header) == 44
        cubes = []
        for a in range(3):                                         # Coordinate<float>(particle.vel * velFactor)
            vel = (disp[a].astype(T) * ratio).astype(T)
            cubes.append((vel * T(velFactor)).astype(np.float32))
        for a in range(3):                                         # particle.pos * lengthFactorDisplacements (= 1)
            cubes.append(disp[a].astype(T).astype(np.float32))
        cubes.append(deltab.astype(np.float32))
        cubes.append(m.astype(np.float32))
        cubes.append((T(pvarValue) * m.astype(T)).astype(np.float32))   # float pvar = pvarValue * mask
        ids = np.arange(n ** 3, dtype=np.uint64) + np.uint64(iordOffset)
        for name, values in zip(names, cubes + [ids]):
            dtype = np.uint64 if name == "ic_particle_ids" else np.float32
            cube = lv.toOutputOrder(values)                        # [x][y][z] in the centred coordinates
            with open(os.path.join(d, name), "wb") as f:
                f.write(struct.pack("<i", 44))
                f.write(header)
                f.write(struct.pack("<i", 44))
                _fortran_planes(f, np.transpose(cube, (2, 1, 0)), dtype)   # file_index = i_y * n + i_x within plane i_z
        iordOffset += n ** 3
        written.append(d)
    return written


def read_grafic_level(path):
    """Reader for tests and tools: header fields and the ten cubes of one level directory as [z][y][x] arrays."""
    out = {}
    for name in sorted(os.listdir(path)):
        with open(os.path.join(path, name), "rb") as f:
            raw = f.read()
        assert struct.unpack("<i", raw[:4])[0] == 44
        hdr = struct.unpack("<iiiffffffff", raw[4:48])
        n = hdr[0]
        dtype = np.uint64 if name == "ic_particle_ids" else np.float32
        plane = n * n * np.dtype(dtype).itemsize
        body = np.frombuffer(raw, dtype=np.uint8, offset=52).reshape(n, plane + 8)
        out[name] = np.ascontiguousarray(body[:, 4:4 + plane]).view(dtype).reshape(n, n, n)
        out["header"] = hdr
    return out
