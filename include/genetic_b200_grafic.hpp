// C++ host layer, part 4: grafic (RAMSES) output, its particle mapper and its refinement masks over the C ABI.
//
// The reference's classes, restated with device-resident fields (paths relative to genetIC/src/):
//   io::grafic::GraficOutput / save                       io/grafic.hpp:36-270
//   particle::mapper::GraficMapper                        simulation/particles/mapper/graficmapper.hpp:20-200
//   multilevelgrid::GraficMask                            simulation/multilevelgrid/mask.hpp:45-228
//   MultiLevelGrid::copyContextWithCenteredIntermediate   simulation/multilevelgrid/multilevelgrid.hpp:337-473
//   grids::CenteredGrid / OffsetGrid / SubSampleGrid / SuperSampleGrid     simulation/grid/virtualgrid.hpp
// A level of the centred intermediate stack is a GLevel: the geometry of a (possibly sub/supersampled) view of a stored grid,
// the integer shift of a CenteredGrid (same cells, visited in another order) or the moved lower corner of an OffsetGrid.
// Fields on virtual levels are made on the device by the cross-grid kernels (gx_interp_add), as the particle mappers do.
// genetic_b200/grafic.py is the same flow in Python; both are pinned on the reference's goldens test_12, test_08a-f,
// mapper_test_02a/b.
#ifndef GENETIC_B200_GRAFIC_HPP
#define GENETIC_B200_GRAFIC_HPP
#include "genetic_b200_zoom.hpp"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <sys/stat.h>

namespace gx {
namespace grafic {
using zoom::Int3;

inline int findPowerOf(size_t base, size_t p) {  // tools::findPowerOf, util_functions.hpp:52-54
  const double r = std::log((double) p) / std::log((double) base);
  if (std::fabs(r - std::round(r)) > 1e-8)
    throw std::runtime_error("Subsample and supersample cannot be converted into an integer number of grids with this base factor");
  return (int) std::lround(r);
}

// Grid::upscaleCellFlagVector / downscaleCellFlagVector (grid.hpp:384-436); the result is sorted and unique
inline std::vector<size_t> scaleCells(const std::vector<size_t> &cells, size_t src_n, size_t dst_n) {
  std::vector<size_t> out;
  if (dst_n >= src_n) {
    const size_t f = dst_n / src_n;
    out.reserve(cells.size() * f * f * f);
    for (size_t id : cells) {
      const size_t x = id / (src_n * src_n), y = (id / src_n) % src_n, z = id % src_n;
      for (size_t a = 0; a < f; ++a)
        for (size_t b = 0; b < f; ++b)
          for (size_t c = 0; c < f; ++c) out.push_back(((x * f + a) * dst_n + (y * f + b)) * dst_n + (z * f + c));
    }
  } else {
    const size_t f = src_n / dst_n;
    out.reserve(cells.size());
    for (size_t id : cells) {
      const size_t x = id / (src_n * src_n), y = (id / src_n) % src_n, z = id % src_n;
      out.push_back(((x / f) * dst_n + (y / f)) * dst_n + (z / f));
    }
  }
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
  return out;
}

template <typename T> struct GLevel {
  std::shared_ptr<grids::Grid<T>> grid;  // geometry of this level: the stored grid itself or a sub/supersampled view of it
  grids::Grid<T> *stored = nullptr;
  size_t storedLevel = 0;
  bool resampled = false, centered = false;
  Int3 shift{0, 0, 0};
  T offsetLower[3] = {0, 0, 0};  // header / centroid offset (OffsetGrid: wrapped lower corner + point offset)

  T wrap(T x) const {
    const T Ls = grid->periodicDomainSize;
    x = std::fmod(x, Ls);
    return x < 0 ? x + Ls : x;
  }
  std::array<T, 3> centroid(size_t id) const {  // grid.hpp:656-669; virtualgrid.hpp:779-782
    const size_t n = grid->size;
    const size_t c[3] = {id / (n * n), (id / n) % n, id % n};
    std::array<T, 3> p;
    for (int a = 0; a < 3; ++a) {
      if (centered) p[a] = wrap((T(c[a]) * grid->cellSize + grid->offsetLower[a] + grid->cellSize / T(2)) + T(shift[a]) * grid->cellSize);
      else p[a] = T(c[a]) * grid->cellSize + offsetLower[a] + grid->cellSize / T(2);
    }
    return p;
  }
  size_t indexFromPoint(std::array<T, 3> p) const {  // grid.hpp:671-677; virtualgrid.hpp:784-787
    size_t c[3];
    for (int a = 0; a < 3; ++a) {
      T x = p[a];
      if (centered) x = x + T(-shift[a]) * grid->cellSize;
      const T lower = centered ? grid->offsetLower[a] : offsetLower[a];
      c[a] = (size_t) std::floor(wrap(x - lower) / grid->cellSize);
    }
    return (c[0] * grid->size + c[1]) * grid->size + c[2];
  }
  // getIndexFromCoordinateNoWrap of the centred grid (virtualgrid.hpp:746-761): the cell visited at output coordinate (x, y, z)
  size_t indexAtOutput(size_t x, size_t y, size_t z) const {
    if (!centered) return (x * grid->size + y) * grid->size + z;
    const long n = (long) grid->size;
    auto w = [n](long v) { v %= n; return (size_t) (v < 0 ? v + n : v); };
    return (w((long) x - shift[0]) * grid->size + w((long) y - shift[1])) * grid->size + w((long) z - shift[2]);
  }
  std::vector<size_t> getFlaggedCells() const {
    std::vector<size_t> v;
    stored->getFlaggedCells(v);
    std::sort(v.begin(), v.end());
    return resampled ? scaleCells(v, stored->size, grid->size) : v;
  }
  void flagCellsAppend(const std::vector<size_t> &ids) {  // Grid::flagCells: union with the flags already there (grid.hpp:269-275)
    std::vector<size_t> add = resampled ? scaleCells(ids, grid->size, stored->size) : ids, have;
    stored->getFlaggedCells(have);
    have.insert(have.end(), add.begin(), add.end());
    std::sort(have.begin(), have.end());
    have.erase(std::unique(have.begin(), have.end()), have.end());
    stored->flagCells(have);
  }
};

template <typename T> class GraficContext {
public:
  std::vector<GLevel<T>> levels;
  GraficContext(multilevelgrid::MultiLevelGrid<T> &ctx, const std::array<T, 3> &centre, size_t subsample, size_t supersample) {
    const int extralow = findPowerOf(2, subsample), extrahigh = findPowerOf(2, supersample);
    const size_t nl = ctx.getNumLevels();
    auto view = [&](size_t level, size_t up, size_t down) {
      auto &g = ctx.getGridForLevel(level);
      GLevel<T> lv;
      lv.stored = &g; lv.storedLevel = level; lv.resampled = up != 1 || down != 1;
      lv.grid = std::make_shared<grids::Grid<T>>(g.periodicDomainSize, g.size * up / down, g.cellSize * T(down) / T(up),
                                                 g.offsetLower[0], g.offsetLower[1], g.offsetLower[2]);
      for (int a = 0; a < 3; ++a) lv.offsetLower[a] = g.offsetLower[a];
      levels.push_back(lv);
    };
    auto addSub = [&](size_t level, int count) {
      size_t factor = (size_t) 1 << count;
      for (int i = 0; i < count; ++i) { view(level, 1, factor); factor /= 2; }
    };
    for (size_t level = 0; level < nl; ++level) {
      if (level == 0) addSub(0, extralow);
      else {
        const double zoomfactor = (double) ctx.getGridForLevel(level - 1).cellSize / (double) ctx.getGridForLevel(level).cellSize;
        addSub(level, (int) (std::round(std::log10(zoomfactor) / std::log10(2.0)) - 1));
      }
      view(level, 1, 1);
    }
    size_t factor = 2;
    for (int i = 0; i < extrahigh; ++i) { view(nl - 1, factor, 1); factor *= 2; }
    // copyContextAndCenter: CenteredGrid for every level that covers the box, OffsetGrid below
    size_t deepest = 0;
    for (size_t i = 0; i < levels.size(); ++i) if (levels[i].grid->coversFullSimulation()) deepest = i;
    T offset[3] = {0, 0, 0};
    for (size_t i = 0; i < levels.size(); ++i) {
      auto &lv = levels[i];
      const auto &g = *lv.grid;
      if (i <= deepest) {
        lv.centered = true;
        for (int a = 0; a < 3; ++a) {
          lv.shift[a] = (long) ((T(0.5) * g.thisGridSize - centre[a]) / g.cellSize);  // Coordinate<int>(...): truncation
          offset[a] = T(lv.shift[a]) * g.cellSize;
        }
      } else {
        for (int a = 0; a < 3; ++a) lv.offsetLower[a] = lv.wrap(g.offsetLower[a] + offset[a]);
      }
    }
  }
  size_t getNumLevels() const { return levels.size(); }
  size_t getIndexOfCellOnOtherLevel(size_t cur, size_t other, size_t id) const {  // multilevelgrid.hpp:481-484
    return levels[other].indexFromPoint(levels[cur].centroid(id));
  }
};

// graficmapper.hpp: ids are the cells of the levels in turn; flags propagate to the coarser real levels
template <typename T> class GraficMapper {
public:
  GraficContext<T> ctx;
  GraficMapper(multilevelgrid::MultiLevelGrid<T> &c, const std::array<T, 3> &centre, size_t subsample, size_t supersample)
      : ctx(c, centre, subsample, supersample) {}
  size_t size() const { size_t s = 0; for (auto &lv : ctx.levels) s += lv.grid->size3; return s; }
  void unflagAllParticles() { for (auto &lv : ctx.levels) lv.stored->flagCells({}); }
  void flagParticles(const std::vector<size_t> &ids) {
    size_t start = 0, i = 0;
    for (auto &lv : ctx.levels) {
      const size_t end = start + lv.grid->size3;
      std::vector<size_t> cells;
      while (i < ids.size() && ids[i] < end) { cells.push_back(ids[i] - start); ++i; }
      lv.flagCellsAppend(cells);
      start = end;
    }
    if (i != ids.size()) throw std::runtime_error("Ran out of grids when interpreting grafic cell IDs - check IDs?");
    std::vector<size_t> real;
    for (size_t l = 0; l < ctx.levels.size(); ++l) if (!ctx.levels[l].resampled) real.push_back(l);
    for (size_t k = real.size(); k-- > 1;) {  // propagateFlagsThroughHierarchy
      std::vector<size_t> up;
      for (size_t f : ctx.levels[real[k]].getFlaggedCells()) up.push_back(ctx.getIndexOfCellOnOtherLevel(real[k], real[k - 1], f));
      std::sort(up.begin(), up.end());
      up.erase(std::unique(up.begin(), up.end()), up.end());
      ctx.levels[real[k - 1]].flagCellsAppend(up);
    }
  }
  void getFlaggedParticles(std::vector<size_t> &out) const {
    size_t offset = 0;
    for (auto &lv : ctx.levels) {
      for (size_t c : lv.getFlaggedCells()) out.push_back(c + offset);
      offset += lv.grid->size3;
    }
  }
};

// mask.hpp:45-228
template <typename T> class GraficMask {
public:
  GraficMask(const GraficContext<T> &c, const std::vector<std::vector<size_t>> &inputMasks) : ctx(&c), input(inputMasks), flagged(c.getNumLevels()) {
    for (auto &m : input) std::sort(m.begin(), m.end());
  }
  void calculateMask(const std::function<void(const std::string &)> &log = nullptr) {
    size_t k = 0;
    const size_t nl = ctx->getNumLevels();
    for (size_t level = 0; level + 1 < nl && k < input.size(); ++level)  // identifyLevelsOfInputMask: matched by index range
      if (!input[k].empty() && input[k].front() < ctx->levels[level].grid->size3) { flagged[level] = input[k]; deepest = (int) level; ++k; }
    if (deepest < 0) {
      if (log) log("WARNING No zoom regions were ever opened. Grafic mask will not be generated in this case");
      return;
    }
    for (int level = deepest; level >= 0; --level)  // generateHierarchyAboveLevelInclusive
      if (flagged[level].empty()) {
        for (size_t i : flagged[level + 1]) flagged[level].push_back(ctx->getIndexOfCellOnOtherLevel(level + 1, level, i));
        std::sort(flagged[level].begin(), flagged[level].end());
        flagged[level].erase(std::unique(flagged[level].begin(), flagged[level].end()), flagged[level].end());
      }
    for (size_t level = (size_t) deepest + 1; level + 1 < nl; ++level)  // generateHierarchyBelowLevelExclusive
      for (size_t i = 0; i < ctx->levels[level].grid->size3; ++i)
        if (std::binary_search(flagged[level - 1].begin(), flagged[level - 1].end(), ctx->getIndexOfCellOnOtherLevel(level, level - 1, i)))
          flagged[level].push_back(i);
  }
  T isInMask(size_t level, size_t cell) const {
    if (flagged[level].empty()) return T(1);
    return std::binary_search(flagged[level].begin(), flagged[level].end(), cell) ? T(1) : T(0);
  }
private:
  const GraficContext<T> *ctx;
  std::vector<std::vector<size_t>> input, flagged;
  int deepest = -1;
};

// ---- io/numpy.hpp: a C-ordered n^3 array as .npy -------------------------------------------------------------------------
template <typename T> void saveNpy(const std::string &fn, const std::vector<T> &data, size_t n) {
  std::ofstream o(fn, std::ios::binary);
  char dict[128];
  snprintf(dict, sizeof dict, "{'descr': '<f%d', 'fortran_order': False, 'shape': (%zu, %zu, %zu), }", (int) sizeof(T), n, n, n);
  std::string h(dict);
  const size_t total = 10 + h.size() + 1;
  h.append((64 - total % 64) % 64, ' ');
  h.push_back('\n');
  const unsigned short hl = (unsigned short) h.size();
  o.write("\x93NUMPY\x01\x00", 8);
  o.write(reinterpret_cast<const char *>(&hl), 2);
  o.write(h.data(), (std::streamsize) h.size());
  o.write(reinterpret_cast<const char *>(data.data()), (std::streamsize) (n * n * n * sizeof(T)));
}

// io::numpy::LoadArrayFromNumpy for the files saveNpy / numpy.save write: version 1.0 header, C order, '<f8' or '<f4'
template <typename T> std::vector<T> loadNpy(const std::string &fn, size_t n) {
  std::ifstream f(fn, std::ios::binary);
  if (!f.is_open()) throw std::runtime_error("File " + fn + " not found");
  char magic[10];
  f.read(magic, 10);
  if (std::memcmp(magic, "\x93NUMPY", 6) != 0) throw std::runtime_error("Not a numpy file: " + fn);
  const size_t hl = (unsigned char) magic[8] | ((size_t) (unsigned char) magic[9] << 8);
  std::string h(hl, ' ');
  f.read(&h[0], (std::streamsize) hl);
  char shape[64];
  snprintf(shape, sizeof shape, "(%zu, %zu, %zu)", n, n, n);
  if (h.find(shape) == std::string::npos) throw std::runtime_error("Incorrect size for imported numpy array");
  if (h.find(sizeof(T) == 8 ? "<f8" : "<f4") == std::string::npos) throw std::runtime_error("Incorrect data type for imported numpy array");
  if (h.find("'fortran_order': False") == std::string::npos) throw std::runtime_error("Imported numpy array must be in C order");
  std::vector<T> data(n * n * n);
  f.read((char *) data.data(), (std::streamsize) (data.size() * sizeof(T)));
  if (!f) throw std::runtime_error("Imported numpy array is truncated");
  return data;
}

// ICGenerator::dumpMask (ic.hpp:973-994): mask-<level>.npy in the coordinates of the centred grids
template <typename T>
void dumpMask(multilevelgrid::MultiLevelGrid<T> &context, const std::array<T, 3> &centre, size_t subsample, size_t supersample,
              const std::vector<std::vector<size_t>> &zoomParticleArray, const std::string &folder) {
  GraficContext<T> gctx(context, centre, subsample, supersample);
  GraficMask<T> mask(gctx, zoomParticleArray);
  mask.calculateMask();
  for (size_t level = 0; level < gctx.getNumLevels(); ++level) {
    const auto &lv = gctx.levels[level];
    const size_t n = lv.grid->size;
    std::vector<T> cube(n * n * n);
    for (size_t x = 0; x < n; ++x)
      for (size_t y = 0; y < n; ++y)
        for (size_t z = 0; z < n; ++z) cube[(x * n + y) * n + z] = mask.isInMask(level, lv.indexAtOutput(x, y, z));
    saveNpy<T>(folder + "/mask-" + std::to_string(level) + ".npy", cube, n);
  }
}

// ---- io::grafic::save (grafic.hpp:36-270) ---------------------------------------------------------------------------------
#pragma pack(push, 1)
struct io_header_grafic {
  int nx, ny, nz;
  float dx, xOffset, yOffset, zOffset, scalefactor, omegaM, omegaL, h0;
};
#pragma pack(pop)

template <typename T> std::vector<T> valuesOnLevel(const GLevel<T> &lv, fields::Field<T> &src) {
  src.toReal();
  if (!lv.resampled) return src.getDataVector();
  fields::Field<T> v(*lv.grid, false);
  v.addFieldFromDifferentGrid(src);
  return v.getDataVector();
}

template <typename T>
void save(const std::string &filename, multilevelgrid::MultiLevelGrid<T> &context, particle::MultiLevelParticleGenerator<T> &generator,
          std::vector<std::shared_ptr<fields::Field<T>>> &overdensity, const cosmology::CosmologicalParameters<T> &cosmo, T pvarValue,
          const std::array<T, 3> &centre, size_t subsample, size_t supersample, const std::vector<std::vector<size_t>> &zoomParticleArray) {
  GraficContext<T> gctx(context, centre, subsample, supersample);
  GraficMask<T> mask(gctx, zoomParticleArray);
  mask.calculateMask();
  const T lengthFactorHeader = T(1.) / cosmo.hubble;
  const T velFactor = (T) std::pow((double) cosmo.scalefactor, 0.5);
  const T ratio = cosmology::zeldovichVelocityToOffsetRatio<T>(cosmo);
  const auto &base = *gctx.levels[0].grid;
  size_t iordOffset = 0;
  static const char *names[10] = {"ic_velcx", "ic_velcy", "ic_velcz", "ic_poscx", "ic_poscy", "ic_poscz", "ic_deltab", "ic_refmap",
                                  "ic_pvar_00001", "ic_particle_ids"};
  for (size_t level = 0; level < gctx.getNumLevels(); ++level) {
    const auto &lv = gctx.levels[level];
    const auto &g = *lv.grid;
    const size_t n = g.size;
    const long effective = std::lround((double) (base.cellSize * T(base.size)) / (double) g.cellSize);
    const std::string dir = filename + "_" + std::to_string(effective);
    mkdir(dir.c_str(), 0777);
    auto &psis = generator.getGeneratedFieldsForLevel(lv.storedLevel);
    std::vector<T> disp[3];
    for (int a = 0; a < 3; ++a) disp[a] = valuesOnLevel<T>(lv, *psis[a]);
    const std::vector<T> deltab = valuesOnLevel<T>(lv, *overdensity[lv.storedLevel]);
    io_header_grafic h;
    h.nx = h.ny = h.nz = (int) n;
    h.dx = (float) (g.cellSize * lengthFactorHeader);
    h.xOffset = (float) (lv.offsetLower[0] * lengthFactorHeader);
    h.yOffset = (float) (lv.offsetLower[1] * lengthFactorHeader);
    h.zOffset = (float) (lv.offsetLower[2] * lengthFactorHeader);
    h.scalefactor = (float) cosmo.scalefactor; h.omegaM = (float) cosmo.OmegaM0; h.omegaL = (float) cosmo.OmegaLambda0;
    h.h0 = (float) (cosmo.hubble * 100);
    std::vector<float> plane(n * n);
    std::vector<size_t> idplane(n * n);
    for (int file = 0; file < 10; ++file) {
      std::ofstream o(dir + "/" + names[file], std::ios::binary);
      const int hl = (int) sizeof(io_header_grafic);
      o.write((const char *) &hl, 4); o.write((const char *) &h, hl); o.write((const char *) &hl, 4);
      for (size_t z = 0; z < n; ++z) {  // one Fortran record per z-plane, file_index = i_y * n + i_x
        for (size_t y = 0; y < n; ++y)
          for (size_t x = 0; x < n; ++x) {
            const size_t i = lv.indexAtOutput(x, y, z);
            const T m = mask.isInMask(level, i);
            float v = 0;
            if (file < 3) v = (float) ((disp[file][i] * ratio) * velFactor);
            else if (file < 6) v = (float) disp[file - 3][i];
            else if (file == 6) v = (float) deltab[i];
            else if (file == 7) v = (float) m;
            else if (file == 8) v = (float) (pvarValue * (T) (float) m);
            plane[y * n + x] = v;
            idplane[y * n + x] = i + iordOffset;
          }
        const int bytes = (int) (n * n * (file == 9 ? sizeof(size_t) : sizeof(float)));
        o.write((const char *) &bytes, 4);
        o.write(file == 9 ? (const char *) idplane.data() : (const char *) plane.data(), bytes);
        o.write((const char *) &bytes, 4);
      }
    }
    iordOffset += n * n * n;
  }
}

}  // namespace grafic
}  // namespace gx
#endif
