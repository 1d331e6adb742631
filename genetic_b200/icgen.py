"""The paramfile command layer: `python -m genetic_b200 paramfile.txt` runs a genetIC parameter file on the device.

Reference: ICGenerator (genetIC/src/ic.hpp) and the command routes of genetIC/src/main.cpp:22-172, dispatched line by line
like tools::Dispatch::run_loop (tools/parser.hpp:156-207): first token = command (case-insensitive), `#` / `%` start a
comment line, surplus tokens are ignored, an error aborts the run with the reference's message format.

Everything numerical is sequencing over the C ABI (host.py / multilevel.py / zoom.py / mapper.py / output.py); this file
holds only the state machine of ICGenerator. Commands outside the accelerated path raise "not supported" errors instead
of being silently ignored: HDF5 / swift output, baryon_tf_on, the *_field variants of the dump commands.
"""
import ctypes as C
import math
import os
import sys
import time

import numpy as np

from . import lib as L
from .host import CosmologicalParameters, Grid, CAMB, PowerLawPowerSpectrum
from . import multilevel as GM, zoom as Z, mapper as PM, output as OUT, grafic as GF


class ICGenerator:
    """ICGenerator<GridDataType> (ic.hpp:62-1762)."""

    def __init__(self, dtype=np.float64, log=None, pUnderlying=None):
        self.dtype = np.dtype(dtype)
        # DummyICGenerator (dummyic.hpp): run for `mapper_relative_to` to rebuild the geometry (grids, flags, mapper) of
        # another parameter file; grids that this generator already has are shared, everything touching fields is a no-op
        self.pUnderlying = pUnderlying
        self.pInputMapper, self.pInputContext = None, None
        # ic.hpp:196-226
        T = self.dtype.type
        self.cosmology = CosmologicalParameters(OmegaM0=float(T(0.3)), OmegaLambda0=None, sigma8=None, redshift=0.0,
                                                ns=float(T(0.96)), hubble=float(T(0.701)), OmegaBaryons0=0.0,
                                                TCMB=float(T(2.725)))
        self.context = GM.MultiLevelGrid()
        self.context.zoomParticleArray = []
        self.spectrum_args = None
        self.seed, self.drawInFourierSpace, self.parallel, self.reverseOrder = None, False, False, False
        self.exactPowerSpectrum = False
        self.haveInitialisedRandomComponent = False
        self.outputField = None
        self.modificationManager = None
        self.pParticleGenerator = None
        self.outputFolder, self.outputFilename, self.outputFormat = "./", "", None
        self.nGadgetFiles = 1
        self.mruGadgetType, self.gadgetTypesForLevels, self.usedGadgetParticleTypes = 1, [], [False] * 6
        self.baryonsOnAllLevels = False
        self.supersample = self.subsample = self.supersampleGas = 1
        self.flaggedGadgetParticleType, self._flaggedCellsForType = None, None
        self.autopad = 0
        self.epsNorm = 0.01075
        self.x0 = self.y0 = self.z0 = 0.0
        self.variance_filterscale = -1.0
        self.initial_number_steps, self.precision = 10, 0.001
        self.velOffset = (0.0, 0.0, 0.0)
        self.pvarValue, self.centerOnTargetRegion = 1.0, False       # grafic output (ic.hpp:126,148)
        self.log = log if log is not None else (lambda msg: print(msg, flush=True))
        if pUnderlying is None:
            GM.set_logger(self.log)
            GM.set_log_precision(6)
        self.commands = self._routes()
        if pUnderlying is not None:
            noop = (lambda *a: None)
            for name in ("zerolevel", "zero_level", "zero_level_field", "dump_grid", "dump_grid_for_field", "dump_ps",
                         "dump_ps_field", "dump_mask", "dump_idfile", "dump_id_file", "modify", "done",
                         "apply_modifications", "calculate", "chi2", "reverse", "reverse_small_k", "import_level",
                         "dump_tipsy", "dump_tipsy_field", "dump_vx", "dump_grid_fourier", "splice"):
                types = self.commands[name][1:] if name in self.commands and self.commands[name][0].__name__ != "fail" else ()
                self.commands[name] = (noop,) + tuple(types)

    # ---- command table (main.cpp:22-172) ---------------------------------------------------------------------------
    def _routes(self):
        f, i, s = float, int, str
        r = {
            "tcmb": (self.setTCMB, f), "om": (self.setOmegaM0, f), "ob": (self.setOmegaB0, f), "ol": (self.setOmegaLambda0, f),
            "hubble": (self.setHubble, f), "s8": (self.setSigma8, f), "ns": (self.setns, f), "zin": (self.setZ0, f),
            "z": (self.setZ0, f),
            "seed": (self.setSeed, i), "random_seed_real_space": (self.setSeed, i),
            "seedfourier": (self.setSeedFourier, i), "random_seed_serial": (self.setSeedFourier, i),
            "seedfourier_reverse": (self.setSeedFourierReverseOrder, i),
            "seedfourier_parallel": (self.setSeedFourierParallel, i), "random_seed": (self.setSeedFourierParallel, i),
            "exact_power_spectrum_enforcement": (self.setExactPowerSpectrumEnforcement,),
            "fix_power": (self.setExactPowerSpectrumEnforcement,),
            "eps_norm": (self.setEpsNorm, f),
            "gadget_particle_type": (self.setGadgetParticleType, i), "gadget_num_files": (self.setGadgetNumFiles, i),
            "camb": (self.setCambDat, s), "powerlaw_amplitude": (self.setPowerLawAmplitude, f),
            "outdir": (self.setOutDir, s), "outname": (self.setOutName, s), "outformat": (self.setOutputFormat, s),
            "basegrid": (self.initBaseGrid, f, i), "base_grid": (self.initBaseGrid, f, i),
            "zoomgrid": (self.initZoomGrid, i, i), "zoom_grid": (self.initZoomGrid, i, i),
            "zoomgrid_with_origin_at": (self.initZoomGridWithLowLeftCornerAt, i, i, i, i, i),
            "zoom_grid_with_origin_at": (self.initZoomGridWithLowLeftCornerAt, i, i, i, i, i),
            "idfile": (self.loadID, s), "id_file": (self.loadID, s),
            "append_idfile": (self.appendID, s), "merge_id_file": (self.appendID, s),
            "dump_idfile": (self.dumpID, s), "dump_id_file": (self.dumpID, s),
            "centre_on": (self.centreParticle, i), "center_on": (self.centreParticle, i),
            "centre": (self.setCentre, f, f, f), "center": (self.setCentre, f, f, f),
            "select_sphere": (self.selectSphere, f), "select_ellipse": (self.selectEllipse, f, f, f),
            "select_cube": (self.selectCube, f), "select_nearest": (self.selectNearest,),
            "expand_flagged_region": (self.expandFlaggedRegion, i),
            "modify": (self.modify, s, s, f), "calculate": (self.calculate, s),
            "filtering_scale": (self.setVarianceFilteringScale, f),
            "clear_modifications": (self.clearModifications,), "done": (self.done,),
            "apply_modifications": (self.applyModifications,), "chi2": (self.getFieldChi2,),
            "reverse": (self.reverse,), "splice": (self.splice, i),
            "dump_grid": (self.dumpGrid, i), "dump_vx": (self.dumpVelocityX, i), "dump_grid_fourier": (self.dumpGridFourier, i),
            "dump_ps": (self.dumpPS, i), "dump_tipsy": (self.saveTipsyArray, s),
            "zerolevel": (self.zeroLevel, i), "zero_level": (self.zeroLevel, i),
            "baryons_all_levels": (self.setBaryonsOnAllLevels,),
            "velocity_offset": (self.setOffset, f, f, f),
        }
        r.update({"mapper_relative_to": (self.setInputMapper, s), "adapt_mask": (self.adaptMask, i), "gadget_flagged_particle_type": (self.setFlaggedGadgetParticleType, i),
                  "supersample": (self.setSupersample, i), "subsample": (self.setSubsample, i),
                  "supersample_gas": (self.setSupersampleGas, i)})
        r["strays_on"] = (self.setStraysOn,)
        r["autopad"] = (self.setAutopad, i)
        r.update({"pvar": (self.setpvarValue, f), "center_grafic_output": (self.setCenteringOnRegion,),
                  "centre_output": (self.setCenteringOnRegion,), "center_output": (self.setCenteringOnRegion,),
                  "dump_mask": (self.dumpMask,), "import_level": (self.importLevel, i, s),
                  "reverse_small_k": (self.reverseSmallK, f)})
        for name in ("num_neighbors", "num_neighbours",
                     "dump_grid_for_field", "dump_grid_fourier_for_field",
                     "dump_ps_field", "dump_tipsy_field",
                     "baryon_tf_on", "zero_level_field"):
            r[name] = (self._unsupported(name),)
        return r

    def _unsupported(self, name):
        def fail(*_):
            raise L.GxError("command %s is outside the accelerated path and not supported by genetic_b200" % name)
        return fail

    def run_line(self, line):
        """Dispatch::run (parser.hpp:156-181)."""
        tok = line.split()
        if not tok or tok[0][0] in "#%":
            return
        name = tok[0].lower()
        if name not in self.commands:
            raise L.GxError("Unknown command " + tok[0])
        fn, *types = self.commands[name]
        if len(tok) - 1 < len(types) and fn.__name__ != "fail":
            raise L.GxError("Insufficient arguments (expected %d)" % len(types))
        try:
            args = [t(v) for t, v in zip(types, tok[1:])]
        except ValueError:
            raise L.GxError("Error while parsing arguments")
        fn(*args)

    # ---- cosmology (ic.hpp:228-343) ----------------------------------------------------------------------------------
    def _T(self, v):
        """CosmologicalParameters<T> holds its members in the field type: `-f` rounds them to float32."""
        return float(self.dtype.type(v))

    def setTCMB(self, v):
        self.cosmology.TCMB = self._T(v)

    def setOmegaM0(self, v):
        self.cosmology.OmegaM0 = self._T(v)

    def setOmegaB0(self, v):
        self.cosmology.OmegaBaryons0 = self._T(v)

    def setOmegaLambda0(self, v):
        self.cosmology.OmegaLambda0 = self._T(v)

    def setHubble(self, v):
        self.cosmology.hubble = self._T(v)

    def setSigma8(self, v):
        self.cosmology.sigma8 = self._T(v)

    def setns(self, v):
        self.cosmology.ns = self._T(v)

    def setZ0(self, z):
        """ic.hpp:339-342: redshift = z; scalefactor = 1. / (redshift + 1.), each stored in T."""
        self.cosmology.setRedshift(self._T(z))
        self.cosmology.scalefactor = self._T(1. / (self.cosmology.redshift + 1.))

    def setEpsNorm(self, v):
        self.epsNorm = v

    def setAutopad(self, nCells):
        self.autopad = nCells

    def setStraysOn(self):
        """ic.hpp:256-260: zoom levels opened from now on keep flagged cells that fall outside their window (the
        particles are interpolated from the level above)."""
        self.context.allowStrays = True

    def setSupersample(self, factor):
        if factor <= 1:
            raise L.GxError("Supersampling factor must be greater then one")
        self.supersample = factor

    def setSupersampleGas(self, factor):
        if factor <= 1:
            raise L.GxError("Supersampling factor must be greater then one")
        self.supersampleGas = factor

    def setSubsample(self, factor):
        if factor <= 0:
            raise L.GxError("Subsampling factor must be greater then zero")
        self.subsample = factor

    def setpvarValue(self, value):
        """ic.hpp:344-347: the value of the passive variable written where the grafic mask is set."""
        self.pvarValue = value

    def setCenteringOnRegion(self):
        """ic.hpp:349-353 (`center_output`): output centred on the point selected at `done` time."""
        self.centerOnTargetRegion = True

    def getBoxCentre(self):
        """ic.hpp:1585-1589"""
        half = float(self.context.getGridForLevel(0).thisGridSize) / 2.0
        return (half, half, half)

    def _graficCentre(self):
        return (self.x0, self.y0, self.z0) if self.centerOnTargetRegion else self.getBoxCentre()

    def dumpMask(self):
        """ic.hpp:973-994: mask-<level>.npy for every level of the centred grafic stack, virtual levels included."""
        gctx = GF.GraficContext(self.context, self._graficCentre(), self.subsample, self.supersample)
        mask = GF.GraficMask(gctx, self.context.zoomParticleArray, self.log)
        mask.calculateMask()
        for level, lv in enumerate(gctx.levels):
            g = lv.grid
            np.save("%s/mask-%d.npy" % (self.outputFolder, level), mask.cube(level).astype(g.dtype))
            with open("%s/mask-info-%d.txt" % (self.outputFolder, level), "w") as fh:
                fh.write("%g %g %g %g\n" % (*[float(v) for v in lv.offsetLower], float(g.thisGridSize)))
                fh.write("The line above contains information about grid level %d\n" % level)
                fh.write("It gives the x-offset, y-offset and z-offset of the low-left corner and also the box length\n")

    def importLevel(self, level, filename):
        """ic.hpp:809-821 + Field::loadGridData (field.hpp:868-891): the real-space field of one level read from a .npy file."""
        self.initialiseRandomComponentIfUninitialised()
        self.log("Importing random field on level %d from %s" % (level, filename))
        fld = self.outputField.getFieldForLevel(level)
        data = np.load(filename)
        n = fld.grid.size
        if data.shape != (n, n, n):
            raise L.GxError("Incorrect size for imported numpy array")
        if data.dtype != self.dtype:
            raise L.GxError("Incorrect data type for imported numpy array")
        fld.setFourier(False)
        fld.setDataVector(data)
        fld.toFourier()
        fld.applyTransferFunction(self.context.getCovariance(level, GM.DM), -0.5)   # back to white noise
        self.log("... success!")

    def setBaryonsOnAllLevels(self):
        self.baryonsOnAllLevels = True

    def setOffset(self, vx, vy, vz):
        """velocity_offset (ic.hpp:285-289): km/s -> km a^1/2 s^-1."""
        T = self.dtype.type
        conversion = T(math.pow(float(T(self.cosmology.scalefactor)), -0.5))
        self.velOffset = tuple(T(T(v) * conversion) for v in (vx, vy, vz))

    def setVarianceFilteringScale(self, v):
        self.variance_filterscale = v

    # ---- seeds (ic.hpp:615-660) --------------------------------------------------------------------------------------
    def _seed(self, seed, fourier, parallel, reverse):
        self.seed, self.drawInFourierSpace, self.parallel, self.reverseOrder = seed, fourier, parallel, reverse

    def setSeed(self, seed):
        self._seed(seed, False, False, False)

    def setSeedFourier(self, seed):
        self._seed(seed, True, False, False)

    def setSeedFourierParallel(self, seed):
        self._seed(seed, True, True, False)

    def setSeedFourierReverseOrder(self, seed):
        self.log("*** Warning: seed_field_fourier_reverse and seedfourier_reverse are deprecated commands and should be avoided.***")
        self._seed(seed, True, False, True)

    def setExactPowerSpectrumEnforcement(self):
        self.exactPowerSpectrum = True

    # ---- files and formats (ic.hpp:662-760) --------------------------------------------------------------------------
    def setGadgetNumFiles(self, n):
        self.nGadgetFiles = n

    def setGadgetParticleType(self, t):
        if t >= 6 or t < 0:
            raise L.GxError("Gadget particle type must be between 0 and 5")
        if t != self.mruGadgetType:
            if self.usedGadgetParticleTypes[t]:
                raise L.GxError("Cannot reuse a gadget particle type; this type was already used in a different section of the ICs")
            self.mruGadgetType = t
        if self.context.getNumLevels() == 0:
            raise L.GxError("Can't set a gadget particle type until a level has been initialised to set it for")
        self.gadgetTypesForLevels[-1] = t

    def setCambDat(self, path):
        self.spectrum_args = ("camb", path)
        self._install_spectrum()

    def setPowerLawAmplitude(self, amplitude):
        self.spectrum_args = ("powerlaw", amplitude)
        self._install_spectrum()

    def _install_spectrum(self):
        kind, arg = self.spectrum_args
        c = self.cosmology
        cos = CosmologicalParameters(c.OmegaM0, c.OmegaLambda0, c.sigma8, c.redshift, c.ns, c.hubble, c.OmegaBaryons0, c.TCMB)
        for species in (GM.DM, GM.ALL):
            spec = CAMB(cos, arg, species, self.dtype) if kind == "camb" else PowerLawPowerSpectrum(cos, arg, self.dtype)
            self.context.setPowerspectrumGenerator(species, spec)

    def setOutDir(self, path):
        self.outputFolder = path

    def setOutName(self, name):
        self.outputFilename = name

    def setOutputFormat(self, fmt):
        """io::operator>>(OutputFormat) accepts a name or its number (io.hpp)."""
        names = {"gadget2": "gadget2", "2": "gadget2", "gadget3": "gadget3", "3": "gadget3", "tipsy": "tipsy", "4": "tipsy",
                 "grafic": "grafic", "5": "grafic"}
        if fmt.lower() not in names:
            if fmt.lower() in ("gadgethdf", "6", "swift", "7"):
                raise L.GxError("output format %s is not supported by genetic_b200 (tipsy, gadget2, gadget3, grafic are)" % fmt)
            raise L.GxError("Unknown output format")
        self.outputFormat = names[fmt.lower()]

    def getOutputPath(self):
        if not self.outputFilename:
            name = "IC_%s_z%s_%d" % ("float" if self.dtype == np.float32 else "doub", "%g" % self.cosmology.redshift,
                                     self.context.getGridForLevel(0).size)
            return self.outputFolder + "/" + name
        return self.outputFolder + "/" + self.outputFilename

    # ---- grids (ic.hpp:365-612) --------------------------------------------------------------------------------------
    def initBaseGrid(self, boxSize, n):
        if self.context.getNumLevels() > 0:
            raise L.GxError("Cannot re-initialize the base grid")
        if self.haveInitialisedRandomComponent:
            raise L.GxError("Trying to initialize a grid after the random field was already drawn")
        self.log("Initialized the base grid:\n  Box length            = %s Mpc/h\n  n                     = %d\n"
                 "  dx                    = %s" % (GM.fmtg(boxSize), n, GM.fmtg(boxSize / n)))
        self.context.addLevel(Grid(boxSize, n, dtype=self.dtype))
        self._levelAdded()

    def _levelAdded(self):
        if self.pUnderlying is not None:
            # DummyICGenerator::addLevelToContext (dummyic.hpp:28-58): share the grid object of the underlying generator
            level = self.context.getNumLevels() - 1
            mine = self.context.grids[level]
            uctx = self.pUnderlying.context
            if uctx.getNumLevels() > level:
                shared = uctx.getGridForLevel(level)
                if shared.size != mine.size:
                    raise L.GxError("Trying to match particles between incompatible simulation setups (wrong grid n)")
                if shared.thisGridSize != mine.thisGridSize:
                    raise L.GxError("Trying to match particles between incompatible simulation setups (wrong grid size)")
                if any(abs(float(a) - float(b)) > 1e-12 * max(1.0, abs(float(a))) for a, b in zip(shared.offsetLower, mine.offsetLower)):
                    raise L.GxError("Trying to match particles between incompatible simulation setups (wrong grid origin)")
                shared.flagCells(mine.getFlaggedCells(), append=True)
                shared.levelIndex = level
                if self.context.outputGrids[level] is mine:
                    self.context.outputGrids[level] = shared
                self.context.grids[level] = shared
            self.gadgetTypesForLevels.append(1)
            return
        self.gadgetTypesForLevels.append(self.mruGadgetType)
        if len(self.gadgetTypesForLevels) > 1:
            self.usedGadgetParticleTypes[self.mruGadgetType] = True

    def initZoomGrid(self, zoomfac, n):
        self._zoom(zoomfac, n, None)

    def initZoomGridWithLowLeftCornerAt(self, x0, y0, z0, zoomfac, n):
        self._zoom(zoomfac, n, (x0, y0, z0))

    def _zoom(self, zoomfac, n, corner):
        if self.haveInitialisedRandomComponent:
            raise L.GxError("Trying to initialize a grid after the random field was already drawn")
        if self.context.getNumLevels() < 1:
            raise L.GxError("Cannot initialise a zoom grid before initialising the base grid")
        up = self.context.getGridForLevel(self.context.getNumLevels() - 1)
        out = self.context.getOutputGridForLevel(self.context.getNumLevels() - 1)
        nflag = len(out.getFlaggedCells())
        if nflag == 0:
            raise L.GxError("Cannot initialise zoom without marking particles to be replaced")
        if corner is None:
            corner = Z.zoomCorner(up, zoomfac, self.context.allowStrays, out)
            if Z.flagsNearZoomBorder(up, corner, zoomfac, out=out):
                self.log("WARNING: Opening a zoom where flagged particles are within 3 pixels of the edge. This is prone to "
                         "numerical errors.")
        g = Z.initZoomGrid(self.context, zoomfac, n, corner=corner)
        self._levelAdded()
        kept = len(self.context.zoomParticleArray[-1])
        missed = self.context.zoomMissedParticles
        if missed > 0:
            self.log("WARNING: the requested zoom particles do not all fit in the requested zoom window")
            if self.context.allowStrays:
                self.log("         of %d particles, %d will be interpolated from LR grid (stray particle mode)" % (nflag, missed))
            else:
                self.log("         of %d particles, %d have been omitted" % (nflag, missed))
            self.log("         to make a new zoom flag list of %d" % kept)
        self.log("Initialized a zoom region:\n  Subbox length         = %s Mpc/h\n  n                     = %d\n"
                 "  dx                    = %s\n  Zoom factor           = %d\n  Num particles         = %d\n"
                 "  Low-left corner in parent grid = (%d, %d, %d)\n"
                 "  Low-left corner (h**-1 Mpc)    = %s, %s, %s\n  Total particles       = %d"
                 % (GM.fmtg(g.thisGridSize), g.size, GM.fmtg(g.cellSize), zoomfac, nflag, *corner, *[GM.fmtg(v) for v in g.offsetLower],
                    self._mapper().size()))

    def _mapper(self, gasFirst=None):
        if self.outputFormat == "grafic":
            # ensureParticleGeneratorInitialised (ic.hpp:1140-1147): grafic writes the grids in turn; ids are their cells
            key = ("grafic", self.context.getNumLevels(), self.subsample, self.supersample)
            if getattr(self, "_mapper_key", None) != key:
                self._mapper_cache = GF.GraficMapper(self.context, (self.x0, self.y0, self.z0), self.subsample,
                                                     self.supersample, self.log)
                self._mapper_key = key
            return self._mapper_cache
        if gasFirst is None:
            gasFirst = self.outputFormat == "tipsy"
        # the reference rebuilds the mapper whenever one of these changes (updateParticleMapper call sites); here it is
        # rebuilt on demand and kept while they stay the same
        key = (self.context.getNumLevels(), tuple(self.gadgetTypesForLevels), float(self.cosmology.OmegaBaryons0),
               float(self.cosmology.OmegaM0), self.baryonsOnAllLevels, gasFirst, self.supersample, self.subsample,
               self.supersampleGas, self.flaggedGadgetParticleType, self.autopad, tuple(id(z) for z in self.context.zoomParticleArray))
        if getattr(self, "_mapper_key", None) != key:
            self._mapper_cache = PM.updateParticleMapper(self.context, self.gadgetTypesForLevels, self.cosmology,
                                                         self.baryonsOnAllLevels, gasFirst=gasFirst,
                                                         supersample=self.supersample, subsample=self.subsample,
                                                         supersampleGas=self.supersampleGas,
                                                         flaggedGadgetParticleType=self.flaggedGadgetParticleType,
                                                         flaggedCells=self._flaggedCellsForType, autopad=self.autopad)
            self._mapper_key = key
        return self._mapper_cache

    # ---- flags (ic.hpp:1236-1524) ------------------------------------------------------------------------------------
    def _ids(self, filename):
        with open(filename) as f:
            return np.unique(np.array([int(t) for t in f.read().split()], dtype=np.int64))   # io::getBuffer + sortAndEraseDuplicate

    def loadID(self, filename):
        self._mapper().unflagAllParticles()                # clearCellFlags (ic.hpp:1238-1245)
        if self.pInputMapper is not None:
            self.pInputMapper.unflagAllParticles()
        self.appendID(filename)

    def appendID(self, filename):
        ids = self._ids(filename)
        if self.pInputMapper is not None:
            # flagCellsCorrespondingToParticles (ic.hpp:1247-1261): ids refer to the other setup; flag there, then carry the
            # flags onto the grids only this setup has, and back
            self.pInputMapper.flagParticles(ids)
            PM.extendParticleListToUnreferencedGrids(self.pInputMapper, self.context)
            PM.extendParticleListToUnreferencedGrids(self._mapper(), self.pInputContext)
        else:
            self._mapper().flagParticles(ids)
        self._getCentre()

    def setInputMapper(self, fname):
        """mapper_relative_to (ic.hpp:1077-1112): the particle ids of later id files refer to the setup of another
        parameter file, whose geometry is rebuilt here by a field-less run of its commands."""
        if self.context.getNumLevels() == 0:
            raise L.GxError("Mapper relative command cannot be used before a basegrid has been initialised.")
        if not os.path.isfile(fname):
            raise L.GxError("Cannot open IC paramfile for relative_to command")
        self.log("")
        self.log("+ Input mapper: computing geometry from " + fname)
        pseudo = ICGenerator(self.dtype, lambda msg: self.log("\n".join("| " + l for l in (str(msg).splitlines() or [""]))),
                             pUnderlying=self)                       # logging::IndentWhileInScope
        pseudo.log("")
        cwd = os.getcwd()
        fname = os.path.abspath(fname)
        os.chdir(os.path.dirname(fname))
        try:
            with open(fname) as f:
                for line in f:
                    if line.strip():
                        pseudo.run_line(line.rstrip("\n"))
        finally:
            os.chdir(cwd)
        self.log("")
        self.pInputMapper, self.pInputContext = pseudo._mapper(), pseudo.context

    def dumpID(self, filename):
        with open(filename, "w") as f:
            for i in self._mapper().getFlaggedParticles():     # io::dumpBuffer: one id per line
                f.write("%d\n" % i)

    def _getCentre(self):
        """ICGenerator::getCentre (ic.hpp:1372-1386): the centre of the flagged cells on the deepest level that has any."""
        level = 0
        for l in range(self.context.getNumLevels() - 1, -1, -1):      # deepestLevelwithFlaggedCells, multilevelgrid.hpp
            if len(self.context.getGridForLevel(l).getFlaggedCells()) > 0:
                level = l
                break
        g = self.context.getGridForLevel(level)
        self.x0, self.y0, self.z0 = Z.getCentreWrapped(g, g.getFlaggedCells())
        GM.set_log_precision(12)          # setprecision(12), sticky (ic.hpp:1384)
        self.log("Centre of region is %s %s %s" % (GM.fmtg(self.x0), GM.fmtg(self.y0), GM.fmtg(self.z0)))

    def centreParticle(self, pid):
        g = self.context.getGridForLevel(0)
        c = np.unravel_index(int(pid), (g.size,) * 3)
        T = g.dtype.type
        self.x0, self.y0, self.z0 = (float(T(T(T(c[a]) * g.cellSize) + g.offsetLower[a]) + g.cellSize / T(2)) for a in range(3))

    def setCentre(self, x, y, z):
        self.x0, self.y0, self.z0 = x, y, z

    def _centre(self):
        return (self.x0, self.y0, self.z0)

    def selectSphere(self, radius):
        Z.selectSphere(self.context, self._centre(), radius)

    def selectEllipse(self, a, b, c):
        Z.selectEllipse(self.context, self._centre(), a, b, c)

    def selectCube(self, side):
        Z.selectCube(self.context, self._centre(), side)

    def selectNearest(self):
        Z.selectNearest(self.context, self._centre())

    def expandFlaggedRegion(self, nCells):
        """ic.hpp:1527-1543."""
        if self.context.getNumLevels() < 1:
            raise L.GxError("No grid has yet been initialised")
        self.log("Expand flagged region by %d cells" % nCells)
        for level, g in enumerate(self.context.grids):
            before = len(g.getFlaggedCells())
            flags = self._expanded(g.getFlaggedCells(), g.size, nCells)
            g.flagCells(flags)
            self.log("  - level %d increased number of flagged cells by %d (now %d)" % (level, len(flags) - before, len(flags)))

    @staticmethod
    def _expanded(flags, n, nCells):
        return PM._expandFlaggedRegion(flags, n, nCells)

    def adaptMask(self, nPadCells):
        """ic.hpp:1545-1575: working up from the deepest zoom, the cells each level replaces become the footprint of the
        level below's zoom cells on it, padded by nPadCells."""
        nl = self.context.getNumLevels()
        if nl < 3:
            raise L.GxError("Adapting the mask requires there to be more than one zoom level")
        z = self.context.zoomParticleArray
        for level in range(nl - 2, 0, -1):
            # the output grids (the stored grids unless strays are on: then the whole extent of the level above at this
            # level's resolution), with independent flags
            this, above = self.context.getOutputGridForLevel(level), self.context.getOutputGridForLevel(level - 1)
            src = Grid(this.periodicDomainSize, this.size, dx=this.cellSize, x0=this.offsetLower[0], y0=this.offsetLower[1],
                       z0=this.offsetLower[2], dtype=this.dtype)
            src.flagCells(z[level])
            cells = Z.proxyFlags(src, above)                                 # makeProxyGridToMatch: section + subsample
            z[level - 1] = self._expanded(cells, above.size, nPadCells)

    def setFlaggedGadgetParticleType(self, t):
        """ic.hpp:703-714: the flagged cells of the finest level become a separate gadget particle type."""
        if t >= 6 or t < 0:
            raise L.GxError("Gadget particle type must be between 0 and 5")
        if self.usedGadgetParticleTypes[t]:
            raise L.GxError("Cannot reuse a gadget particle type; this type was already used in a different section of the ICs")
        self.flaggedGadgetParticleType = t
        self.usedGadgetParticleTypes[t] = True
        # cells of the mapper's finest grid, i.e. of the OUTPUT grid of the finest level (ic.hpp:1182-1183)
        self._flaggedCellsForType = np.array(self.context.getOutputGridForLevel(self.context.getNumLevels() - 1).getFlaggedCells())

    # ---- random component, modifications (ic.hpp:1333-1346,1589-1745) ------------------------------------------------
    def initialiseRandomComponentIfUninitialised(self):
        if self.haveInitialisedRandomComponent:
            return
        if self.context.getNumLevels() == 0:
            raise L.GxError("No grid has been initialised")
        if self.seed is None:
            raise L.GxError("No random seed has been specified")
        self.log("Drawing random numbers")
        self.outputField = GM.draw(self.context, self.seed, self.drawInFourierSpace, self.parallel, self.reverseOrder)
        if self.exactPowerSpectrum:
            self.outputField.enforceUnitVariance()
        self.modificationManager = GM.ModificationManager(self.context, self.cosmology, self.outputField)
        self.haveInitialisedRandomComponent = True

    def getFieldChi2(self):
        self.initialiseRandomComponentIfUninitialised()
        GM.set_log_precision(10)          # std::setprecision(10), sticky on the reference's log stream (ic.hpp:1596)
        self.log("Calculated chi^2 = %s (dof = %d)" % (GM.fmtg(self.outputField.getChi2()), self.context.getNumDof()))

    def _modkw(self):
        return dict(filterscale=self.variance_filterscale, initNumberSteps=self.initial_number_steps, precision=self.precision)

    def calculate(self, name):
        self.initialiseRandomComponentIfUninitialised()
        self.log("%s: calculated value = %s" % (name, GM.fmtg(self.modificationManager.calculateCurrentValueByName(name, **self._modkw()))))

    def modify(self, name, type_, target):
        self.initialiseRandomComponentIfUninitialised()
        # ICGenerator::modify(string, string, float target): the target passes through a C float (ic.hpp:1606)
        self.modificationManager.addModificationToList(name, type_, float(np.float32(target)), **self._modkw())

    def clearModifications(self):
        if self.modificationManager is not None:
            self.modificationManager.clearModifications()
        else:
            self.log("Clearing modification list")

    def applyModifications(self):
        self.initialiseRandomComponentIfUninitialised()
        self.modificationManager.applyModifications()

    def reverse(self):
        self.initialiseRandomComponentIfUninitialised()
        self.outputField.reverse()

    def reverseSmallK(self, kmax):
        """ic.hpp:1690-1708: the sign of the modes below kmax reversed on every level."""
        self.initialiseRandomComponentIfUninitialised()
        for level in range(self.context.getNumLevels()):
            fld = self.outputField.getFieldForLevel(level)
            fld.toFourier()
            fld.lib.gx_reverse_small_k(fld.ptr, fld.n, fld.dt, float(fld.grid.thisGridSize), float(kmax), fld.stream)

    def splice(self, newSeed):
        """ic.hpp:1706-1743."""
        self.initialiseRandomComponentIfUninitialised()
        if self.outputField.getTransferType() != GM.WHITENOISE:
            raise L.GxError("It is too late in the IC generation process to perform splicing; try moving the splice command earlier")
        if self.context.getNumLevels() > 1:
            self.log("Splicing operations on multiple levels work, but may have edge artefacts. Use experimentally/at your own risk!")
        self.log("Constructing new random field for exterior of splice")
        newField = GM.draw(self.context, newSeed)           # a fresh RandomFieldGenerator: real space, serial (:1718-1727)
        self.log("Finished constructing new random field. Beginning splice operation.")
        for level in range(self.context.getNumLevels()):
            orig = self.outputField.getFieldForLevel(level)
            cov = self.context.getCovariance(level, GM.DM).getPowerSpectrumForGrid(orig.grid)
            spliced = GM.spliceOneLevel(newField.getFieldForLevel(level), orig, cov)
            spliced.toFourier()
            self.outputField.fieldsOnLevels[level] = spliced

    def zeroLevel(self, level):
        self.log("*** Warning: your script calls zeroLevel(%d). This is intended for testing purposes only!" % level)
        self.initialiseRandomComponentIfUninitialised()
        if level > self.outputField.getNumLevels() - 1:
            raise L.GxError("Specified level does not exist.")
        f = self.outputField.getFieldForLevel(level)
        f.lib.gx_memset_zero(f.ptr, f.nbytes, f.stream)

    # ---- output (ic.hpp:820-900,1275-1331,1633-1648) -----------------------------------------------------------------
    def done(self):
        if self.outputFormat is None:
            raise L.GxError("No output format specified!")
        self.initialiseRandomComponentIfUninitialised()
        if self.modificationManager.hasModifications():
            self.applyModifications()
        self.write()

    def write(self):
        self.initialiseRandomComponentIfUninitialised()
        if self.outputField.getTransferType() == GM.WHITENOISE:
            self.outputField.applyPowerSpectrumFor(GM.ALL)                      # applyPowerSpec (ic.hpp:802-818)
        if self._streamableUnigrid():
            return self._writeUnigridStreaming()
        if self.pParticleGenerator is None:
            self.log("Calculating particles from overdensity fields...")
            self.pParticleGenerator = GM.MultiLevelParticleGenerator(self.outputField)
        if self.outputFormat == "grafic":
            # ic.hpp:1312-1324: the grids in turn, centred on the selected point with `center_output`
            self.outputField.toReal()
            GF.save(self.getOutputPath() + ".grafic", self.context, self.pParticleGenerator, self.outputField, None,
                    self.cosmology, self.pvarValue, self._graficCentre(), self.subsample, self.supersample,
                    self.context.zoomParticleArray, self.log)
            self.log("Finished writing initial conditions")
            return
        t0 = time.perf_counter()
        mapper = self._mapper()
        t1 = time.perf_counter()
        self.log("Writing output; number dm particles=%d, number gas particles=%d" % (mapper.size() - mapper.size_gas(), mapper.size_gas()))
        boxlen = float(self.context.getGridForLevel(0).periodicDomainSize)
        offset = any(float(v) != 0.0 for v in self.velOffset)
        if self.outputFormat == "tipsy" or offset:
            parts = PM.MappedParticles(mapper, self.pParticleGenerator, self.cosmology, out_dtype=self.dtype, epsNorm=self.epsNorm)
            if offset:                                                            # OffsetEvaluator (offsetgenerator.hpp:38-43)
                parts.vel += np.array(self.velOffset, dtype=self.dtype)
        else:
            parts = PM.MappedParticles(mapper, self.pParticleGenerator, self.cosmology, out_dtype=np.float32, epsNorm=self.epsNorm)
        t2 = time.perf_counter()
        if self.outputFormat == "tipsy":
            OUT.saveTipsy(self.getOutputPath() + ".tipsy", boxlen, parts, self.cosmology, aux_dir=".")
        else:
            OUT.saveGadget(self.getOutputPath() + ".gadget", boxlen, parts, self.cosmology, int(self.outputFormat[-1]),
                           self.nGadgetFiles)
        if os.environ.get("GX_TIMING", "0") == "1":
            self.log("[timing] mapper %.3f s, particle evaluation + copy %.3f s, file %.3f s"
                     % (t1 - t0, t2 - t1, time.perf_counter() - t2))
        self.log("Finished writing initial conditions")

    # ---- large single-level ICs: two fields on the device, particles streamed into the memory-mapped output file --------
    def _streamableUnigrid(self):
        """One level, dark matter only, no resampling or retyping: particle id = cell index, so the offsets can be formed
        one component at a time (the fused transform of gx_zeldovich_offset) and evaluated by id range straight into the
        file -- delta + one psi field on the device instead of delta + three, nothing of size n^3 on the host. Used from
        512^3 upwards (GX_STREAM_UNIGRID=1/0 forces it on / off); a 2048^3 f64 IC needs 2 x 68.8 GB of the 180 GB."""
        env = os.environ.get("GX_STREAM_UNIGRID", "")
        if env == "0" or self.context.getNumLevels() != 1 or self.pParticleGenerator is not None or self.outputFormat == "grafic":
            return False
        if float(self.cosmology.OmegaBaryons0) > 0 or self.supersample > 1 or self.subsample > 1 or \
                self.flaggedGadgetParticleType is not None or self.nGadgetFiles != 1:
            return False
        return env == "1" or self.context.getGridForLevel(0).size >= 512

    def _writeUnigridStreaming(self, chunk=1 << 24):
        import ctypes as C
        from .host import Field, _dt
        lib = L.load()
        grid = self.context.getGridForLevel(0)
        delta = self.outputField.getFieldForLevel(0)
        delta.toFourier()
        n, n3 = grid.size, grid.size3
        T = self.dtype.type
        boxlen = float(grid.periodicDomainSize)
        mass, eps = PM.particle_mass(grid, self.cosmology), PM.particle_eps(grid, self.epsNorm)
        self.log("Writing output; number dm particles=%d, number gas particles=0" % n3)
        tipsy = self.outputFormat == "tipsy"
        offset = any(float(v) != 0.0 for v in self.velOffset)
        wide = tipsy or offset                       # unit conversions / the offset are applied in T before the float32 store
        out_dtype = self.dtype if wide else np.dtype(np.float32)
        esz, odt = out_dtype.itemsize, _dt(out_dtype)
        if tipsy:
            # TipsyOutput::operator() (tipsy.hpp:207-271); the mass sum over identical particles, sequentially in double
            ft = np.dtype(self.dtype).type               # FloatType tot_mass: float under -f (tipsy.hpp:218)
            tot, m = ft(0.0), ft(mass)
            for c0 in range(0, n3, 1 << 22):            # tot_mass += mass, one particle after the other (tipsy.hpp:217-225)
                tot = np.cumsum(np.concatenate([[tot], np.full(min(1 << 22, n3 - c0), m, dtype=ft)]).astype(ft), dtype=ft)[-1]
            tot = float(tot)
            mass_factor = float(self.cosmology.OmegaM0) / tot
            pos_factor = 1. / boxlen
            dKpcUnit = boxlen * 1000 / float(self.cosmology.hubble)
            dMsolUnit = 1e10 / float(self.cosmology.hubble) / mass_factor
            dKmsUnit = float(np.sqrt(4.3022682e-6 * dMsolUnit / dKpcUnit))
            vel_factor = float(self.cosmology.scalefactor) ** -0.5 / dKmsUnit
            fn = self.getOutputPath() + ".tipsy"
            with open(fn, "wb") as f:
                f.write(OUT.struct.pack("<diiiii4x", float(self.cosmology.scalefactor), OUT._wrap_int32(n3), 3, 0,
                                        OUT._wrap_int32(n3), 0))
                f.truncate(32 + n3 * 36)
            rec = np.memmap(fn, dtype=np.float32, mode="r+", offset=32, shape=(n3, 9))
            with open("tipsy.param", "w") as f:
                f.write("dKpcUnit = %g\ndMsolUnit = %g\ndHubble0 = %g\nbComove = 1 \n"
                        % (dKpcUnit, dMsolUnit, 0.1 * float(self.cosmology.hubble) * dKpcUnit / dKmsUnit))
            for c0 in range(0, n3, chunk):
                rec[c0:c0 + chunk, 0] = np.float32(float(mass) * mass_factor)
                rec[c0:c0 + chunk, 7] = np.float32(float(eps) * pos_factor)
                rec[c0:c0 + chunk, 8] = 0.0
        else:
            version = int(self.outputFormat[-1])
            fn = self.getOutputPath() + ".gadget" + str(version)
            counts = [0] * 6
            counts[self.gadgetTypesForLevels[0]] = n3
            masses = [0.0] * 6
            masses[self.gadgetTypesForLevels[0]] = float(mass)
            header = OUT._gadget_header(version, masses, counts, counts, 1, boxlen, self.cosmology, False)
            blocks = [("pos", 12), ("vel", 12), ("ids", 8)]
            with open(fn, "wb") as f:
                OUT._fortran(f, header)
                where = {}
                for name, width in blocks:
                    nbytes = n3 * width
                    marker = OUT.struct.pack("<i", nbytes if nbytes <= 0x7fffffff else 0x7fffffff)
                    f.write(marker)
                    where[name] = f.tell()
                    f.seek(nbytes, 1)
                    f.write(marker)
            posm = np.memmap(fn, dtype=np.float32, mode="r+", offset=where["pos"], shape=(n3, 3))
            velm = np.memmap(fn, dtype=np.float32, mode="r+", offset=where["vel"], shape=(n3, 3))
            idm = np.memmap(fn, dtype=np.uint64, mode="r+", offset=where["ids"], shape=(n3,))
            for c0 in range(0, n3, chunk):
                idm[c0:c0 + chunk] = np.arange(c0, min(n3, c0 + chunk), dtype=np.uint64)
        ratio = self.cosmology.zeldovichVelocityToOffsetRatio(grid.dtype)
        m = min(chunk, n3)
        dpos, dvel = C.c_void_p(), C.c_void_p()
        lib.gx_malloc(C.byref(dpos), m * esz)
        lib.gx_malloc(C.byref(dvel), m * esz)
        hp, hv = np.empty(m, dtype=out_dtype), np.empty(m, dtype=out_dtype)
        psi = Field(grid, fourier=True, stream=delta.stream)
        try:
            for axis in range(3):
                lib.gx_zeldovich_offset(delta.ptr, psi.ptr, n, delta.dt, float(grid.thisGridSize), axis, delta.stream)
                for c0 in range(0, n3, chunk):
                    cnt = min(chunk, n3 - c0)
                    lib.gx_emit_component(psi.ptr, n, psi.dt, axis, float(grid.cellSize), float(grid.offsetLower[axis]),
                                          float(grid.periodicDomainSize), ratio, dpos, dvel, odt, 1, c0, cnt, delta.stream)
                    lib.gx_memcpy_d2h(hp.ctypes.data_as(C.c_void_p), dpos, cnt * esz, delta.stream)
                    lib.gx_memcpy_d2h(hv.ctypes.data_as(C.c_void_p), dvel, cnt * esz, delta.stream)
                    lib.gx_synchronize(delta.stream)
                    p, v = hp[:cnt], hv[:cnt]
                    if offset:
                        v = v + T(self.velOffset[axis])
                    if tipsy:
                        rec[c0:c0 + cnt, 1 + axis] = p.astype(np.float64) * pos_factor - 0.5
                        rec[c0:c0 + cnt, 4 + axis] = v.astype(np.float64) * vel_factor
                    else:
                        posm[c0:c0 + cnt, axis] = p
                        velm[c0:c0 + cnt, axis] = v
        finally:
            lib.gx_free(dpos)
            lib.gx_free(dvel)
        for mm in ((rec,) if tipsy else (posm, velm, idm)):
            mm.flush()
        self.log("Finished writing initial conditions")

    def _dumpGridData(self, level, field, prefix="grid"):
        """ic.hpp:822-851 (+ numpy.hpp): the n^3 array and the info text file."""
        g = field.grid
        name = "%s/%s-%d" % (self.outputFolder, prefix, level)
        data = field.getDataVector()
        np.save(name + ".npy", data)
        with open("%s/%s-info-%d%s.txt" % (self.outputFolder, prefix, level, "-fourier" if field.isFourier() else ""), "w") as f:
            f.write("%g %g %g %g\n" % (*[float(v) for v in g.offsetLower], float(g.thisGridSize)))
            f.write("The line above contains information about grid level %d\n" % level)
            f.write("It gives the x-offset, y-offset and z-offset of the low-left corner and also the box length\n")

    def dumpGrid(self, level):
        self.initialiseRandomComponentIfUninitialised()
        self.outputField.toReal()
        self._dumpGridData(level, self.outputField.getFieldForLevel(level))

    def dumpGridFourier(self, level):
        """ic.hpp:879-885 + getComplexFourierField (fourier.hpp:685-716): the full complex n^3 Fourier cube."""
        self.initialiseRandomComponentIfUninitialised()
        self.outputField.toFourier()
        f = self.outputField.getFieldForLevel(level)
        half = f.getDataVector()
        n = f.grid.size
        full = np.zeros((n, n, n), dtype=half.dtype)
        full[:, :, :n // 2 + 1] = half
        idx = (-np.arange(n)) % n
        kz = np.arange(n // 2 + 1, n)
        full[:, :, kz] = np.conj(half[idx][:, idx][:, :, n - kz])
        np.save("%s/grid-%d.npy" % (self.outputFolder, level), full)

    def dumpPS(self, level):
        """ic.hpp:943-962: <outname>_<level>.ps, a host-side diagnostic of the level's Fourier field."""
        self.initialiseRandomComponentIfUninitialised()
        f = self.outputField.getFieldForLevel(level)
        f.toFourier()
        cov = self.context.getCovariance(level, GM.ALL).getPowerSpectrumForGrid(f.grid)
        bins = np.zeros((4, 100), dtype=np.float64)        # the binning reduction runs on the device
        f.lib.gx_power_spectrum_bins(f.ptr, cov.ptr, f.grid.size, f.dt, float(f.grid.thisGridSize), 100,
                                     bins.ctypes.data_as(C.c_void_p), None)
        OUT.writePowerSpectrum(bins, f.grid, "%s_%d.ps" % (self.getOutputPath(), level))

    def saveTipsyArray(self, filename):
        """ic.hpp:853-862: the overdensity of every particle's cell in mapper order (`dump_tipsy`), a host-side diagnostic."""
        self.initialiseRandomComponentIfUninitialised()
        self.outputField.toReal()
        OUT.saveFieldTipsyArray(filename, PM.field_on_mapper(self._mapper(), self.outputField.fieldsOnLevels))

    def dumpVelocityX(self, level):
        if self.pParticleGenerator is None:
            raise L.GxError("dump_vx: the particle generator has not been initialised (call done first)")
        self._dumpGridData(level, self.pParticleGenerator.getGeneratorForLevel(level).getGeneratedFields()[0], "vx")


def run_paramfile(paramfile, dtype=np.float64, log=None):
    """runGenetic (main.cpp:199-229): runs from the paramfile's directory, stops at the first error."""
    paramfile = os.path.abspath(paramfile)
    if not os.path.isfile(paramfile):
        raise L.GxError("could not open parameter file " + paramfile)
    gen = ICGenerator(dtype, log)
    timing = os.environ.get("GX_TIMING", "0") == "1"       # per-command wall clock in the log
    cwd = os.getcwd()
    os.chdir(os.path.dirname(paramfile))
    try:
        with open(paramfile) as f:
            for lineno, line in enumerate(f, 1):
                line = line.rstrip("\n")
                if not line.strip():
                    continue
                try:
                    t0 = time.perf_counter()
                    gen.run_line(line)
                    if timing and time.perf_counter() - t0 > 0.01:
                        gen.log("[timing] %-40s %.3f s" % (line.strip()[:40], time.perf_counter() - t0))
                except (L.GxError, OSError, ValueError) as e:
                    raise L.GxError('Error "%s" on line %d ("%s")' % (e, lineno, line)) from e
    finally:
        os.chdir(cwd)
    return gen


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv or argv[0] in ("-h", "--help"):
        print("Usage: python -m genetic_b200 paramfile [-f]\n\n The paramfile is a text file of genetIC commands.\n"
              " If option -f is specified, float (32-bit) is used instead of double (64-bit) internally.")
        return 0 if argv else 1
    try:
        run_paramfile(argv[0], np.float32 if "-f" in argv[1:] else np.float64)
    except L.GxError as e:
        print(str(e), file=sys.stderr)
        return 1
    return 0
