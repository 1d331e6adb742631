"""Python host layer for the multi-level (zoom) part of the hot path: the reference's class surface over the C ABI.

Every class forwards to the device library (include/genetic_b200.h); the control flow is the reference's, operation for
operation, so that results agree with it to rounding. Compile-time options follow the reference's default Makefile
(genetIC/Makefile:23): CUBIC_INTERPOLATION, FILTER_ON_COARSE_GRID, ZELDOVICH_GRADIENT_FOURIER_SPACE on,
VELOCITY_MODIFICATION_GRADIENT_FOURIER_SPACE off. Paths relative to genetIC/src/:

  multilevelgrid::MultiLevelGrid               simulation/multilevelgrid/multilevelgrid.hpp:30-330
  filters::FilterFamily                        simulation/filters/filterfamily.hpp:86-143
  fields::MultiLevelField / ConstraintField    simulation/field/multilevelfield.hpp:18-606
  particle::MultiLevelParticleGenerator        simulation/particles/multilevelgenerator.hpp:79-128
  modifications::*Modification                 simulation/modifications/linearmodification.hpp, quadraticmodification.hpp
  modifications::ModificationManager           simulation/modifications/modificationmanager.hpp
  modifications::spliceOneLevel                simulation/modifications/splice.hpp:38-118
  tools::numerics::conjugateGradient           tools/numerics/cg.hpp:14-58
"""
import math

import numpy as np

from . import lib as L
from .host import (Field, Grid, FilterProduct, LowPassFermiFilter, ComplementaryCovarianceFilter, ComplementaryFilter,
                   RandomFieldGenerator, ZeldovichParticleGenerator)

FRACTIONAL_K_SPLIT = 0.3  # Makefile:23
WHITENOISE, DM, BARYON, ALL = "whitenoise", "dm", "baryon", "all"  # particle::species (species.hpp)


_log = [None]


# The reference's log is one ostream: `std::setprecision(10)` in ICGenerator::getChi2 (ic.hpp:1596) and `setprecision(12)` in
# getCentre (ic.hpp:1384) stay in force for every number logged afterwards. fmtg() formats like that stream.
_precision = [6]


def set_log_precision(p):
    _precision[0] = int(p)


def fmtg(v):
    return "%.*g" % (_precision[0], float(v))


def set_logger(fn):
    """Where the reference's classes write to logging::entry() (warnings of the modifications, the chi^2 lines of
    ModificationManager::applyModifications); None = silent."""
    _log[0] = fn


def log(msg):
    if _log[0] is not None:
        _log[0](msg)


class MultiLevelGrid:
    """multilevelgrid::MultiLevelGrid: the grids of all levels, their weights and the covariance provider."""

    def __init__(self):
        self.grids = []
        self.weights = []
        self.levelsAreCombined = False
        self.allowStrays = False     # strays_on: output grids of later zoom levels become ResolutionMatchingGrids
        self.outputGrids = []
        self.spectra = {}     # species -> PowerSpectrum (cosmology::CAMB); evaluated in-kernel per grid

    def addLevel(self, grid):
        """multilevelgrid.hpp:112-137: weight = (cellSize / cellSize_0)^3."""
        self.weights.append(1.0 if not self.grids else float(grid.cellSize / self.grids[0].cellSize) ** 3.0)
        grid.levelIndex = len(self.grids)         # which level's fields a (view of this) grid's particles come from
        if self.allowStrays and self.grids:        # multilevelgrid.hpp:131-137
            from .host import ResolutionMatchingGrid
            self.outputGrids.append(ResolutionMatchingGrid(grid, self.outputGrids[-1]))
        else:
            self.outputGrids.append(grid)
        self.grids.append(grid)

    def addZoomLevel(self, zoomfac, n, corner):
        """ICGenerator::initZoomGridWithLowLeftCornerAt (ic.hpp:502-598) without the particle-mapper bookkeeping:
        appends a grid of n cells spanning 1/zoomfac of the level above, its lower corner at cell `corner` of that level."""
        up = self.grids[-1]
        T = up.dtype.type
        off = [T(up.offsetLower[a] + T(corner[a]) * up.cellSize) for a in range(3)]
        g = Grid(up.periodicDomainSize, n, dx=T(T(up.thisGridSize / T(zoomfac)) / T(n)), x0=off[0], y0=off[1], z0=off[2],
                 dtype=up.dtype)
        self.addLevel(g)
        return g

    def setPowerspectrumGenerator(self, species, spectrum):
        self.spectra[species] = spectrum

    def getNumLevels(self):
        return len(self.grids)

    def getGridForLevel(self, level):
        return self.grids[level]

    def getOutputGridForLevel(self, level):
        """The grid particles are output from (multilevelgrid.hpp:190-197): the stored grid unless strays are allowed."""
        return self.outputGrids[level]

    def getWeightForLevel(self, level):
        return self.weights[level]

    def getCovariance(self, level, species):
        return self.spectra.get(species if species in self.spectra else ALL)

    def setLevelsAreCombined(self):
        self.levelsAreCombined = True

    def getNumDof(self):
        """multilevelgrid.hpp:160-172"""
        dof = 0
        for i, g in enumerate(self.grids):
            dof += g.size3
            if i > 0:
                f = int(round(float(self.grids[i - 1].cellSize / g.cellSize))) ** 3
                dof -= g.size3 // f
        return dof


class FilterFamily:
    """filters::FilterFamily (filterfamily.hpp:86-143). Once the levels have been combined (the fine fields carry the
    low-frequency information too) every level but the finest gets the null filter and the finest the identity."""

    def __init__(self, context):
        ident = FilterProduct([])
        if context.levelsAreCombined:
            null = ident * 0.0
            nl = context.getNumLevels()
            self.filters = [null] * (nl - 1) + [ident]
            self.lpFilters, self.hpFilters = list(self.filters), list(self.filters)
            return
        self.filters, self.lpFilters, self.hpFilters = [ident], [ident], [ident]
        for level in range(context.getNumLevels() - 1):
            g = context.getGridForLevel(level)
            k_cut = FRACTIONAL_K_SPLIT * (float(g.size) * float(g.getFourierKmin()))
            filt = self.filters.pop()
            self.lpFilters.pop()
            self.filters.append(filt * LowPassFermiFilter(k_cut))
            self.lpFilters.append(LowPassFermiFilter(k_cut))
            self.filters.append(filt * ComplementaryCovarianceFilter(k_cut))
            self.hpFilters.append(ComplementaryCovarianceFilter(k_cut))
            self.lpFilters.append(ident)

    def getFilterForLevel(self, level):
        return self.filters[level]

    def getHighPassFilterForLevel(self, level):
        return self.hpFilters[level]

    def getLowPassFilterForLevel(self, level):
        return self.lpFilters[level]


def _apply_filter(field, filt):
    if filt.factors():
        field.applyFilter(filt)
    else:
        field.toFourier()  # forEachFourierCell converts first (field.hpp:741-746); the identity multiply is skipped


class MultiLevelField:
    """fields::MultiLevelField: one Field per level."""

    def __init__(self, context, fields=None, transferType=WHITENOISE, isCovector=False):
        self.context = context
        self.fieldsOnLevels = fields if fields is not None else [Field(g, fourier=True) for g in context.grids]
        self.transferType = transferType
        self.isCovector = isCovector

    def copy(self):
        return MultiLevelField(self.context, [f.copy() for f in self.fieldsOnLevels], self.transferType, self.isCovector)

    def getNumLevels(self):
        return len(self.fieldsOnLevels)

    def getFieldForLevel(self, level):
        return self.fieldsOnLevels[level]

    def getTransferType(self):
        return self.transferType

    def getFilters(self):
        return FilterFamily(self.context)

    def toFourier(self):
        for f in self.fieldsOnLevels:
            f.toFourier()

    def toReal(self):
        for f in self.fieldsOnLevels:
            f.toReal()

    def applyFilters(self, filters=None):
        filters = filters or self.getFilters()
        for level, f in enumerate(self.fieldsOnLevels):
            _apply_filter(f, filters.getFilterForLevel(level))

    # ---- linear algebra (multilevelfield.hpp:167-294) ---------------------------------------------------------
    def addScaled(self, other, scale):
        self.toFourier()
        for a, b in zip(self.fieldsOnLevels, other.fieldsOnLevels):
            assert b.fourier
            a.addScaledFourier(b, scale)

    def divide(self, ratio):
        """operator/=(ratio) (multilevelfield.hpp:172-179), as one multiplication by 1/ratio per element."""
        for f in self.fieldsOnLevels:
            f.lib.gx_scale(f.ptr, f.n, f.dt, 1.0 / float(ratio), f.stream)

    def reverse(self):
        for f in self.fieldsOnLevels:
            f *= -1.0

    def innerProduct(self, other):
        if not self.isCovector:
            raise L.GxError("The inner product can only be taken if one of the fields is a covector")
        if other.isCovector:
            v = other.copy()
            v.convertToVector()
            v.toFourier()
            return self.innerProduct(v)
        tot = 0.0
        for a, b in zip(self.fieldsOnLevels, other.fieldsOnLevels):
            assert a.fourier and b.fourier
            tot += a.innerProduct(b)
        return tot

    # ---- metric (multilevelfield.hpp:315-391) ------------------------------------------------------------------
    def convertToCovector(self):
        assert not self.isCovector
        self._applyMetric(True)
        self.isCovector = True

    def convertToVector(self):
        assert self.isCovector
        self._applyMetric(False)
        self.isCovector = False

    def _applyMetric(self, toCovector):
        filters = self.getFilters()
        ctx = self.context
        self.toFourier()
        nl = self.getNumLevels()
        new = []
        for level in range(nl):
            result = self.fieldsOnLevels[level].copy()
            f = filters.getFilterForLevel(level)
            if toCovector and level < nl - 1:
                window = ctx.getGridForLevel(level + 1)
                if f.factors():
                    result.applyFilterInWindow(f, window, True)
                    result.applyFilterInWindow(f, window, False)
                # (an identity filter leaves the windowed recombination equal to the field itself)
            else:
                _apply_filter(result, f * f)
            for source_level in range(nl):
                if source_level == level:
                    continue
                ratio = ctx.getWeightForLevel(level) / ctx.getWeightForLevel(source_level)
                flt = filters.getFilterForLevel(level) * filters.getFilterForLevel(source_level) * math.sqrt(ratio)
                result.addFieldFromDifferentGridWithFilter(self.fieldsOnLevels[source_level], flt)
            result.toFourier()
            new.append(result)
        self.fieldsOnLevels = new

    def getChi2(self):
        """multilevelfield.hpp:461-487"""
        if self.transferType != WHITENOISE:
            raise L.GxError("Cannot calculate the chi^2 after the output field has been generated")
        self.toFourier()
        c = self.copy()
        c.convertToCovector()
        return c.innerProduct(self)

    # ---- transfer functions (multilevelfield.hpp:410-454) ----------------------------------------------------------
    def enforceUnitVariance(self):
        for f in self.fieldsOnLevels:
            f.enforceUnitVariance()

    def applyPowerSpectrumFor(self, species):
        self.applyTransferRatio(self.transferType, species)
        self.transferType = species

    def applyTransferRatio(self, old, new):
        if old == new:
            return
        self.toFourier()
        for level, f in enumerate(self.fieldsOnLevels):
            if old != WHITENOISE:
                f.applyTransferFunction(self.context.getCovariance(level, old), -0.5)
            if new != WHITENOISE:
                f.applyTransferFunction(self.context.getCovariance(level, new), 0.5)


def draw(context, seed, drawInFourierSpace=False, parallel=False, reverse=False):
    """RandomFieldGenerator::draw into a new white-noise MultiLevelField (randomfieldgenerator.hpp:122-139)."""
    field = MultiLevelField(context)
    rfg = RandomFieldGenerator(field.fieldsOnLevels)
    rfg.seed(seed)
    rfg.setDrawInFourierSpace(drawInFourierSpace)
    rfg.setParallel(parallel)
    rfg.setReverseRandomDrawOrder(reverse)
    rfg.draw()
    return field


# ---- particle generator (multilevelgenerator.hpp:79-128) -------------------------------------------------------------
class MultiLevelParticleGenerator:
    """initialiseParticleGeneratorBasedOnTemplate for ZeldovichParticleGenerator: per level the overdensity and the three
    offset fields, high-pass filtered and completed with the low-pass filtered, tricubically interpolated level above."""

    def __init__(self, overdensityField):
        ctx = overdensityField.context
        nl = ctx.getNumLevels()
        overdensityField.toFourier()
        self.overdensityField = overdensityField
        # unfused form: the offsets must stay in Fourier space until the levels have been combined
        self.pGenerators = [_FourierOffsets(overdensityField.getFieldForLevel(l)) for l in range(nl)]
        if nl >= 2:
            filters = overdensityField.getFilters()
            for level in range(1, nl):
                hp = filters.getHighPassFilterForLevel(level)
                lp = filters.getLowPassFilterForLevel(level - 1)
                d = overdensityField.getFieldForLevel(level)
                _apply_filter(d, hp)
                for p in self.pGenerators[level].pOff:
                    _apply_filter(p, hp)
                d.addFieldFromDifferentGridWithFilter(overdensityField.getFieldForLevel(level - 1), lp)
                for p, q in zip(self.pGenerators[level].pOff, self.pGenerators[level - 1].pOff):
                    p.addFieldFromDifferentGridWithFilter(q, lp)
            ctx.setLevelsAreCombined()
        for g in self.pGenerators:
            for p in g.pOff:
                p.toReal()

    def getGeneratorForLevel(self, level):
        return self.pGenerators[level]


class _FourierOffsets:
    """ZeldovichParticleGenerator::calculateOffsetFields up to generateNewFourierFields (zeldovich.hpp:152-181), leaving
    the three fields in Fourier space (the single-level class goes straight to real space through the fused passes)."""

    def __init__(self, delta):
        delta.toFourier()
        self.grid = delta.grid
        self.pOff = []
        for axis in range(3):
            psi = Field(delta.grid, fourier=True, stream=delta.stream)
            delta.lib.gx_zeldovich_fourier(delta.ptr, psi.ptr, delta.n, delta.dt, float(delta.grid.thisGridSize), axis,
                                           delta.stream)
            self.pOff.append(psi)

    def getGeneratedFields(self):
        return self.pOff


# ---- modifications ---------------------------------------------------------------------------------------------------
def generateMultilevelCovectorFromHiresCovector(context, hires, transferType):
    """multilevelgrid.hpp:245-321. `hires`: Field on the finest grid (Fourier space, as calculateLocalisationCovector
    leaves it). Returns a covector MultiLevelField in real space."""
    nl = context.getNumLevels()
    fields = [None] * nl
    fields[-1] = hires
    for level in range(nl - 1):
        fields[level] = Field(context.getGridForLevel(level), fourier=False)
    if nl > 1:
        fields[-1].toReal()
        for level in range(nl - 2, -1, -1):
            hi, lo = fields[level + 1], fields[level]
            r = int(round(float(lo.grid.cellSize / hi.grid.cellSize)))
            lo.lib.gx_deinterp_add(lo.ptr, L.C.byref(lo.grid.desc()), hi.ptr, L.C.byref(hi.grid.desc()), lo.dt,
                                   1.0 / float(r ** 3), lo.stream)
    for level in range(nl):
        fields[level] *= context.getWeightForLevel(level) / context.getWeightForLevel(nl - 1)
    ret = MultiLevelField(context, fields, transferType, True)
    ret.applyFilters(FilterFamily(context))
    ret.toReal()
    return ret


def _warnIfWholeZoomGridIsFlagged(context):
    """Modification::Modification (modification.hpp:36-47)."""
    for level in range(1, context.getNumLevels()):
        g = context.getGridForLevel(level)
        if len(g.flags) == g.size3:
            log("WARNING: Region selected for modification is the entire zoom grid. This is likely because the cell "
                "selection extends beyond the zoom boundaries.")
            log("By design, modifications are meant to be defined inside a zoom region. Increase the size of your zoom grid "
                "or decrease your selection to avoid nasty surprises.")


def _flaggedCellsPhysicalSize(grid):
    """Grid::getFlaggedCellsPhysicalSize (grid.hpp:318-334): the largest side of the periodic window holding the flags."""
    from . import zoom as Z
    flags = grid.getFlaggedCells()
    if len(flags) == 0:
        return 0.0
    n = grid.size
    coords = np.stack(np.unravel_index(flags, (n, n, n)), axis=1)
    w = Z._Window(grid.simEquivalentSize, [int(v) for v in coords[0]])
    w.expandToIncludeAll(coords)
    return float(grid.dtype.type(max(w.sizes())) * grid.cellSize)


class LinearModification:
    """modifications::LinearModification (linearmodification.hpp:24-107)."""
    order = 1

    def __init__(self, context, cosmology):
        self.context, self.cosmology = context, cosmology
        _warnIfWholeZoomGridIsFlagged(context)
        g = context.getGridForLevel(context.getNumLevels() - 1)
        if len(g.flags) == 0:
            raise L.GxError("Linear modifications use exclusively information from the highest resolution grid "
                            "but no cells have been flagged on this grid.\n"
                            "Change flagged cells selection to use linear modifications.")
        self.flaggedCellsFinestGrid = g.flags.copy()
        self.target = None

    def setTarget(self, t):
        self.target = t

    def getTarget(self):
        return self.target

    def calculateLocalisationCovector(self, grid):
        data = np.zeros(grid.size3, dtype=grid.dtype)
        w = grid.dtype.type(1.0 / len(self.flaggedCellsFinestGrid))
        np.add.at(data, self.flaggedCellsFinestGrid, w)
        f = Field(grid, data, fourier=False)
        f.toFourier()
        return f

    def getCovector(self, forSpecies):
        g = self.context.getGridForLevel(self.context.getNumLevels() - 1)
        r = generateMultilevelCovectorFromHiresCovector(self.context, self.calculateLocalisationCovector(g), forSpecies)
        self.turnLocalisationCovectorIntoModificationCovector(r)
        r.toFourier()
        return r

    def calculateCurrentValue(self, field):
        return self.getCovector(field.getTransferType()).innerProduct(field)


class OverdensityModification(LinearModification):
    def turnLocalisationCovectorIntoModificationCovector(self, field):
        field.applyTransferRatio(field.getTransferType(), DM)   # linearmodification.hpp:122-125 (type is left unchanged)


class VelocityModification(OverdensityModification):
    """linearmodification.hpp:148-285, real-space 4th-order derivative (the default build)."""

    def __init__(self, context, cosmology, direction):
        super().__init__(context, cosmology)
        self.direction = direction

    def calculateLocalisationCovector(self, grid):
        n = grid.size
        T = grid.dtype.type
        w = T(1.0 / len(self.flaggedCellsFinestGrid))
        a, b = T(-w / 12. / grid.cellSize), T(w * 2. / 3. / grid.cellSize)
        data = np.zeros((n, n, n), dtype=grid.dtype)
        c = np.stack(np.unravel_index(self.flaggedCellsFinestGrid, (n, n, n)), axis=0)
        S = grid.simEquivalentSize

        def step(cc, s):  # getIndexFromIndexAndStep wraps with the simulation period (grid.hpp:555-574)
            cc = cc.copy()
            d = cc[self.direction] + s
            d = np.where(d > S - 1, d - S, d)
            cc[self.direction] = np.where(d < 0, d + S, d)
            return cc
        m1, p1 = step(c, -1), step(c, 1)
        m2, p2 = step(m1, -1), step(p1, 1)
        np.add.at(data, tuple(m2), a)
        np.add.at(data, tuple(m1), b)
        np.add.at(data, tuple(p1), -b)
        np.add.at(data, tuple(p2), -a)
        f = Field(grid, data, fourier=False)
        f.toFourier()
        return f

    def turnLocalisationCovectorIntoModificationCovector(self, field):
        if self.context.levelsAreCombined:
            raise L.GxError("Cannot create velocity covector for recombined fields (i.e. after output generation is complete)")
        super().turnLocalisationCovectorIntoModificationCovector(field)
        scale = self.cosmology.zeldovichVelocityToOffsetRatio(field.getFieldForLevel(0).dtype)
        for f in field.fieldsOnLevels:
            f.toFourier()
            f.lib.gx_scale_inv_k2(f.ptr, f.n, f.dt, float(f.grid.thisGridSize), -scale, 0, f.stream)


class PotentialModification(OverdensityModification):
    """linearmodification.hpp:128-146 with cosmology::densityToPotential (parameters.hpp:208-225)."""

    def turnLocalisationCovectorIntoModificationCovector(self, field):
        super().turnLocalisationCovectorIntoModificationCovector(field)
        c = self.cosmology
        prefac = 3. / 2. * c.OmegaM0 / c.scalefactor * 100. * 100. / 299792. / 299792.
        for f in field.fieldsOnLevels:
            f.toFourier()
            f.lib.gx_scale_inv_k2(f.ptr, f.n, f.dt, float(f.grid.thisGridSize), -prefac, 1, f.stream)


class AngMomentumModification(OverdensityModification):
    """linearmodification.hpp:288-411: specific angular momentum of the flagged Lagrangian patch about its centre,
    -(q - q_c) x grad(Phi) with the 4th-order real-space stencil."""

    def __init__(self, context, cosmology, direction):
        super().__init__(context, cosmology)
        if direction not in (0, 1, 2):
            raise L.GxError("Angular momentum direction must be 0 (x), 1 (y) or 2 (z)")
        self.direction = direction

    def calculateLocalisationCovector(self, grid):
        from . import zoom as Z
        n, S = grid.size, grid.simEquivalentSize
        flags = self.flaggedCellsFinestGrid
        qc = Z.getCentreWrapped(grid, flags)
        norm = len(flags) * float(grid.cellSize)
        a, b = 1. / 12. / norm, -2. / 3. / norm
        d, d1, d2 = self.direction, (self.direction + 1) % 3, (self.direction + 2) % 3
        c = np.stack(np.unravel_index(flags, (n, n, n)), axis=0)
        q = [(c[ax].astype(np.float64) * float(grid.cellSize) + float(grid.offsetLower[ax])) + float(grid.cellSize) / 2
             for ax in range(3)]
        dq = [Z.getWrappedOffset(grid, qc[ax], q[ax]) for ax in range(3)]
        coeff = {d: np.zeros(len(flags)), d2: dq[d1], d1: -dq[d2]}
        data = np.zeros((n, n, n), dtype=grid.dtype)

        def step(cc, ax, s):
            cc = cc.copy()
            v = cc[ax] + s
            v = np.where(v > S - 1, v - S, v)
            cc[ax] = np.where(v < 0, v + S, v)
            return cc
        for ax in range(3):
            if ax == d:
                continue
            m1, p1 = step(c, ax, -1), step(c, ax, 1)
            m2, p2 = step(m1, ax, -1), step(p1, ax, 1)
            np.add.at(data, tuple(m2), coeff[ax] * a)
            np.add.at(data, tuple(m1), coeff[ax] * b)
            np.add.at(data, tuple(p1), -coeff[ax] * b)
            np.add.at(data, tuple(p2), -coeff[ax] * a)
        f = Field(grid, data, fourier=False)
        f.toFourier()
        return f

    def turnLocalisationCovectorIntoModificationCovector(self, field):
        super().turnLocalisationCovectorIntoModificationCovector(field)
        scale = self.cosmology.zeldovichVelocityToOffsetRatio(field.getFieldForLevel(0).dtype)
        for f in field.fieldsOnLevels:
            f.toFourier()
            f.lib.gx_scale_inv_k2(f.ptr, f.n, f.dt, float(f.grid.thisGridSize), -scale, 0, f.stream)


class FilteredVarianceModification:
    """modifications::FilteredVarianceModification (quadraticmodification.hpp:82-221)."""
    order = 2

    def __init__(self, context, cosmology, filterscale, initNumberSteps=10, targetPrecision=0.001):
        if filterscale < 0.0:
            raise L.GxError("Trying to calculate filtered variance without initialising variance filtering scale."
                            " Use filtering_scale command to do this.")
        fine = context.getGridForLevel(context.getNumLevels() - 1)
        if filterscale < float(fine.cellSize):
            raise L.GxError("Variance high-pass filtering scale is smaller than the smallest pixel size.")
        _warnIfWholeZoomGridIsFlagged(context)
        window_size = _flaggedCellsPhysicalSize(fine)
        if filterscale > window_size:                       # checkFilterScale, quadraticmodification.hpp:128-149
            log("WARNING: High-pass filtering scale: %s h**-1 Mpc is greater than the rough window size: %s h**-1 Mpc used "
                "for modifications." % (fmtg(filterscale), fmtg(window_size)))
            log("This is prone to numerical errors when modifying the field. Decrease filtering scale to avoid it.")
        self.context, self.cosmology, self.scale = context, cosmology, float(filterscale)
        self.initNumberSteps, self.targetPrecision = initNumberSteps, targetPrecision
        self.flaggedCells = [g.flags.copy() for g in context.grids]
        self.target = None

    def setTarget(self, t):
        self.target = t

    def getTarget(self):
        return self.target

    def calculateCurrentValue(self, field):
        pushed = self.pushMultiLevelFieldThroughMatrix(field)
        pushed.toFourier()
        return pushed.innerProduct(field)

    def pushMultiLevelFieldThroughMatrix(self, field):
        result = MultiLevelField(self.context, [f.copy() for f in field.fieldsOnLevels], field.getTransferType(), False)
        result.applyTransferRatio(field.getTransferType(), DM)
        for level, f in enumerate(result.fieldsOnLevels):
            self._pushOneLevel(f, level)
        result.convertToCovector()
        result.applyTransferRatio(field.getTransferType(), DM)
        return result

    def _pushOneLevel(self, f, level):
        grid = self.context.getGridForLevel(level)
        assert np.array_equal(grid.flags, self.flaggedCells[level]), "flags changed since the modification was defined"
        mask = grid.flagMask()
        hp = ComplementaryFilter(2 * math.pi / self.scale)
        count = len(grid.flags)
        f.toReal()
        f.mulMask(mask)                  # windowOperator
        f.applyFilter(hp)                # filterOperator
        f.toReal()
        # varianceOperator: window, subtract the mean over the flagged cells, divide by their number
        s = L.C.c_double()
        if count:
            f.lib.gx_masked_sum(f.ptr, mask, f.n, f.dt, L.C.byref(s), f.stream)
            f.lib.gx_masked_shift_scale(f.ptr, mask, f.n, f.dt, s.value / count, 1.0 / count, f.stream)
        else:
            f.mulMask(mask)
        f.applyFilter(hp)
        f.toReal()
        f.mulMask(mask)


class ModificationManager:
    """modifications::ModificationManager (modificationmanager.hpp:20-385)."""

    def __init__(self, context, cosmology, outputField):
        self.context, self.cosmology, self.outputField = context, cosmology, outputField
        self.linearModificationList, self.quadraticModificationList = [], []

    def _fromName(self, name, **kw):
        nm = name.lower()
        if nm == "overdensity":
            return OverdensityModification(self.context, self.cosmology)
        if nm == "potential":
            return PotentialModification(self.context, self.cosmology)
        if nm in ("vx", "vy", "vz"):
            return VelocityModification(self.context, self.cosmology, "xyz".index(nm[1]))
        if nm in ("lx", "ly", "lz"):
            return AngMomentumModification(self.context, self.cosmology, "xyz".index(nm[1]))
        if nm == "variance":
            return FilteredVarianceModification(self.context, self.cosmology, kw.get("filterscale", -1.0),
                                                kw.get("initNumberSteps", 10), kw.get("precision", 0.001))
        raise L.GxError(name + " is an unknown modification name")

    def calculateCurrentValueByName(self, name, **kw):
        self.outputField.toFourier()
        return self._fromName(name, **kw).calculateCurrentValue(self.outputField)

    def addModificationToList(self, name, type_, target, **kw):
        m = self._fromName(name, **kw)
        if type_.lower() == "relative":
            target = target * m.calculateCurrentValue(self.outputField)
        elif type_.lower() != "absolute":
            raise L.GxError("Modification type must be either 'relative' or 'absolute'")
        m.setTarget(target)
        (self.linearModificationList if m.order == 1 else self.quadraticModificationList).append(m)

    def hasModifications(self):
        return bool(self.linearModificationList or self.quadraticModificationList)

    def clearModifications(self):
        log("Clearing modification list")
        self.linearModificationList, self.quadraticModificationList = [], []

    def applyModifications(self):
        logging = _log[0] is not None                # the chi^2 bookkeeping of modificationmanager.hpp:93-125 is log output only
        pre = self.outputField.getChi2() if logging else None
        log("")
        log("Applying modifications")
        self._applyModifications()
        if logging:
            post = self.outputField.getChi2()
            log("   Post-modification chi^2 = %s" % fmtg(post))
            log("  Modification Delta chi^2 = %s" % fmtg(post - pre))
            log("           d.o.f. in field = %d" % self.context.getNumDof())
            log("")

    def _applyModifications(self):
        covectors = [m.getCovector(self.outputField.getTransferType()) for m in self.linearModificationList]
        targets = [m.getTarget() for m in self.linearModificationList]
        self._orthonormalise(covectors, targets)
        self._applyLinear(covectors, targets)
        for q in self.quadraticModificationList:
            n0 = q.initNumberSteps
            test = self.outputField.copy()
            self._performIterations(test, covectors, q, n0)
            achieved = abs(q.calculateCurrentValue(test) - q.getTarget())
            n_steps = n0 * int(math.ceil(math.sqrt(achieved / (q.getTarget() * q.targetPrecision))))
            self._performIterations(self.outputField, covectors, q, n_steps if n_steps > n0 else n0)

    @staticmethod
    def _orthonormalise(alphas, targets):
        for i, ai in enumerate(alphas):
            for j in range(i):
                r = ai.innerProduct(alphas[j])
                ai.addScaled(alphas[j], -r)
                targets[i] -= r * targets[j]
            norm = math.sqrt(ai.innerProduct(ai))
            ai.divide(norm)
            targets[i] /= norm

    def _applyLinear(self, alphas, targets):
        existing = [a.innerProduct(self.outputField) for a in alphas]
        for a, t, e in zip(alphas, targets, existing):
            a.convertToVector()
            a.toFourier()
            self.outputField.addScaled(a, t - e)

    @staticmethod
    def _performIterations(field, alphas, q, n_steps):
        # The reference evaluates the current value (a push through the matrix + an inner product) and then pushes the same
        # field again (modificationmanager.hpp:309-323); the second push repeats the first bit for bit, so one is kept.
        def push_and_value():
            p = q.pushMultiLevelFieldThroughMatrix(field)
            p.toFourier()
            return p, p.innerProduct(field)
        pushed, cur = push_and_value()
        start, end = cur, q.getTarget()
        if n_steps in (0, 1) or start == end:          # tools::linspace (util_functions.hpp:78-92)
            targets = [end]
        else:
            step = (end - start) / n_steps
            targets = [start + c * step for c in range(n_steps + 1)]
        for i in range(n_steps):
            if i > 0:
                pushed, cur = push_and_value()
            norm = math.sqrt(pushed.innerProduct(pushed))
            for a in alphas:                             # addToOrthonormalFamily (:306-319); the alphas are vectors by now
                r = pushed.innerProduct(a)
                pushed.addScaled(a, -r)
            pushed.divide(math.sqrt(pushed.innerProduct(pushed)))
            mult = 0.5 * (targets[i + 1] - cur) / norm
            pushed.convertToVector()
            field.addScaled(pushed, mult)


# ---- splicing -----------------------------------------------------------------------------------------------------------
def conjugateGradient(Q, b, rtol=1e-6, atol=1e-12, log=None):
    """tools::numerics::conjugateGradient (cg.hpp:14-58) on real-space device fields."""
    residual = b.copy()
    direction = -residual
    x = Field(b.grid, fourier=False)
    scale = b.norm()
    if scale == 0.0:
        return x
    dimension = b.grid.size3
    for i in range(dimension + 1):
        Qd = Q(direction)
        alpha = -residual.innerProduct(direction) / direction.innerProduct(Qd)
        x.addScaled(direction, alpha)
        residual = Q(x)
        residual -= b
        norm = residual.norm()
        if norm < rtol * scale or norm < atol:
            break
        if log is not None:
            log.append(norm)
        beta = residual.innerProduct(Qd) / direction.innerProduct(Qd)
        direction *= beta
        direction -= residual
    return x


def spliceOneLevel(a, b, cov):
    """modifications::spliceOneLevel (splice.hpp:38-118): a, b white-noise Fields on one grid, cov the covariance Field
    (getPowerSpectrumForGrid). The flagged cells of the grid are the mask. Returns the spliced field (real space)."""
    grid = a.grid
    mask, maskCompl = grid.flagMask(False), grid.flagMask(True)
    a.toFourier()
    b.toFourier()
    pre = cov.copy()
    pre.setFourierCoefficient(0, 0, 0, 1.0 + 0.0j)

    def T(f, power):
        f.toFourier()
        f.applyTransferFunction(pre, power)
        f.toReal()
        return f

    delta_diff = b.copy()
    delta_diff -= a
    T(delta_diff, 0.5)
    z = delta_diff.copy()
    z.mulMask(mask)
    T(z, -1.0)
    z.mulMask(maskCompl)
    T(z, 0.5)

    def X(inp):
        v = inp.copy()
        T(v, 0.5)
        v.mulMask(maskCompl)
        T(v, -1.0)
        v.mulMask(maskCompl)
        T(v, 0.5)
        return v

    alpha = conjugateGradient(X, z)
    T(alpha, 0.5)
    bInDeltaBasis = T(b.copy(), 0.5)
    alpha.mulMask(maskCompl)
    alpha += bInDeltaBasis
    delta_diff.mulMask(mask)
    alpha -= delta_diff
    T(alpha, -0.5)
    return alpha
