// genetic_b200_zoom.hpp -- C++ host layer, part 2: zoom levels end to end.
//
// Cell selection and zoom-window placement (ICGenerator::select*, initZoomGrid, ic.hpp:398-598,1447-1524; Window<int>,
// simulation/window.hpp), the particle mappers (simulation/particles/mapper/onelevelmapper.hpp, twolevelmapper.hpp,
// gasmapper.hpp; ICGenerator::updateParticleMapper, ic.hpp:1129-1221), the multi-level particle generator
// (simulation/particles/multilevelgenerator.hpp:79-128), the multi-level tipsy and gadget2/3 writers (io/tipsy.hpp:207-271,
// io/gadget.hpp:116-556) and the paramfile commands that drive them. Same design as the Python mirror
// (genetic_b200/zoom.py, mapper.py, output.py, icgen.py): a mapper is resolved once into its layout -- (layer, cell) per
// output particle, host index arithmetic -- and every layer's particles are evaluated on the device at its cell list
// (gx_emit_cells). Not in C++ yet (Python only): sub/supersampling, strays, autopad, mapper_relative_to, modifications.
//
// Status: the host logic (selection, window placement, layouts, id flagging, masses) is checked on CPU against the Python
// layer by tests/test_cpp_zoom.py through `host/genetic_b200_zoom --layout-only`; on the B200 the driver's gadget3 file of a
// two-level zoom is byte-identical to the reference binary's (profiles/r1s25_*, tests/test_cpp_zoom.py -m gpu). The writers
// (tipsy with gas records, gadget2/3 with gas, per-level types, variable masses, multi-file) are checked on CPU byte for
// byte against the Python writers through `--particles-from`.
#ifndef GENETIC_B200_ZOOM_HPP
#define GENETIC_B200_ZOOM_HPP

#include "genetic_b200.hpp"

#include <algorithm>
#include <iostream>
#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <numeric>

namespace gx {

// ---------------------------------------------------------------------------------------------------------------------
namespace zoom {
using Int3 = std::array<long, 3>;

template <typename T> Int3 coordinateFromIndex(const grids::Grid<T> &g, size_t id) {
  const long n = (long) g.size;
  return {(long) (id / (size_t) (n * n)), (long) ((id / (size_t) n) % (size_t) n), (long) (id % (size_t) n)};
}
template <typename T> T centroid(const grids::Grid<T> &g, int axis, long c) {  // grid.hpp:662-669
  return (T(c) * g.cellSize + g.offsetLower[axis]) + g.cellSize / T(2);
}
template <typename T> T wrappedOffset(const grids::Grid<T> &base, T x0, T x1) {  // Grid::getWrappedOffset, grid.hpp:445-454
  T d = x0 - x1;
  if (d > base.periodicDomainSize / 2) d -= base.periodicDomainSize;
  if (d < -base.periodicDomainSize / 2) d += base.periodicDomainSize;
  return d;
}

// ICGenerator::select (ic.hpp:1458-1484): on every level flag the cells whose centroid offset from the centre is included
template <typename T>
void select(multilevelgrid::MultiLevelGrid<T> &ctx, const std::array<T, 3> &centre, const std::function<bool(T, T, T)> &inc) {
  const auto &base = ctx.getGridForLevel(0);
  for (size_t level = 0; level < ctx.getNumLevels(); ++level) {
    auto &g = ctx.getGridForLevel(level);
    const long n = (long) g.size;
    std::vector<T> d[3];
    for (int a = 0; a < 3; ++a) {
      d[a].resize(n);
      for (long c = 0; c < n; ++c) d[a][c] = wrappedOffset(base, centroid(g, a, c), centre[a]);
    }
    std::vector<size_t> flags;
    for (long x = 0; x < n; ++x)
      for (long y = 0; y < n; ++y)
        for (long z = 0; z < n; ++z)
          if (inc(d[0][x], d[1][y], d[2][z])) flags.push_back((size_t) ((x * n + y) * n + z));
    g.flagCells(flags);
  }
}
template <typename T> void selectSphere(multilevelgrid::MultiLevelGrid<T> &ctx, const std::array<T, 3> &c, float radius) {
  const T r2 = T(radius * radius);  // selectSphere(float radius), ic.hpp:1490-1496
  select<T>(ctx, c, [r2](T x, T y, T z) { return x * x + y * y + z * z < r2; });
}
template <typename T> void selectCube(multilevelgrid::MultiLevelGrid<T> &ctx, const std::array<T, 3> &c, float side) {
  const T h = T(side / 2);  // ic.hpp:1519-1524
  select<T>(ctx, c, [h](T x, T y, T z) { return std::abs(x) < h && std::abs(y) < h && std::abs(z) < h; });
}
template <typename T> void selectEllipse(multilevelgrid::MultiLevelGrid<T> &ctx, const std::array<T, 3> &c, float a, float b, float cc) {
  const T a2 = T(a * a), b2 = T(b * b), c2 = T(cc * cc);  // selectEllipse(float a, b, c), ic.hpp:1503-1513
  select<T>(ctx, c, [a2, b2, c2](T x, T y, T z) { return x * x / a2 + y * y / b2 + z * z / c2 < T(1.); });
}
template <typename T> void selectNearest(multilevelgrid::MultiLevelGrid<T> &ctx, const std::array<T, 3> &centre) {  // ic.hpp:1447-1455
  for (size_t l = 0; l < ctx.getNumLevels(); ++l) ctx.getGridForLevel(l).flagCells({});
  auto &g = ctx.getGridForLevel(ctx.getNumLevels() - 1);
  long c[3];
  for (int a = 0; a < 3; ++a) {
    T x = std::fmod(centre[a] - g.offsetLower[a], g.periodicDomainSize);
    if (x < 0) x += g.periodicDomainSize;
    c[a] = (long) std::floor(x / g.cellSize);
  }
  g.flagCells({(size_t) ((c[0] * (long) g.size + c[1]) * (long) g.size + c[2])});
}

// Window<int> (simulation/window.hpp): a periodic cuboid [lower, upper) per axis
struct IntWindow {
  long wrap;
  Int3 lo, hi;
  IntWindow(long wrap_, Int3 p) : wrap(wrap_), lo(p), hi{p[0] + 1, p[1] + 1, p[2] + 1} {}
  long w(long v) const { if (v < 0) v += wrap; if (v >= wrap) v -= wrap; return v; }
  static bool within(long l, long h, long v) { return h >= l ? (l <= v && v < h) : (v < h || v >= l); }
  void expandToInclude(const Int3 &p) {
    for (int a = 0; a < 3; ++a) {
      const long l = lo[a], h = hi[a], v = p[a];
      if (within(l, h, v)) continue;
      const long sizeLower = w(h - v - 1) + 1, sizeUpper = w(v - l) + 1;
      const bool canLower = within(v, h, l), canUpper = within(l, v + 1, h - 1);
      if (canLower && (sizeLower < sizeUpper || !canUpper)) lo[a] = v; else hi[a] = v + 1;
    }
  }
  Int3 sizes() const { return {w(hi[0] - lo[0] - 1) + 1, w(hi[1] - lo[1] - 1) + 1, w(hi[2] - lo[2] - 1) + 1}; }
  void wrapCorners() { for (int a = 0; a < 3; ++a) { lo[a] = w(lo[a]); hi[a] = w(hi[a] - 1) + 1; } }
  void shift(const Int3 &d) { for (int a = 0; a < 3; ++a) { lo[a] += d[a]; hi[a] += d[a]; } wrapCorners(); }
  void expandSymmetricallyToSize(long n) {
    const Int3 old = lo;
    if (n == wrap) { lo = {0, 0, 0}; hi = {n, n, n}; return; }
    const long up = n % 2 == 0 ? n / 2 - 1 : n / 2;
    for (int a = 0; a < 3; ++a) {
      const long c = w(lo[a] + w(hi[a] - 1 - lo[a]) / 2);
      lo[a] = w(c - n / 2);
      hi[a] = w(c + up) + 1;
    }
    const long ref[3] = {old[0], old[2], old[2]};  // (the reference compares the y corner with the old *z* corner; kept)
    for (int a = 0; a < 3; ++a) if (lo[a] > ref[a]) { lo[a] -= 1; hi[a] -= 1; }
  }
  long wrappedOffset(long a, long b) const { long d = b - a; if (d < -(wrap / 2)) d += wrap; if (d > wrap / 2) d -= wrap; return d; }
};

// ICGenerator::initZoomGrid (ic.hpp:398-479): lower-left corner, in cells of `grid`, of the window of grid.size / zoomfac
// cells that holds the flagged cells
template <typename T> Int3 zoomCorner(const grids::Grid<T> &grid, size_t zoomfac) {
  std::vector<size_t> flags;
  grid.getFlaggedCells(flags);
  if (flags.empty()) throw std::runtime_error("Cannot initialise zoom without marking particles to be replaced");
  const long n = (long) grid.size, wrap = (long) grid.simEquivalentSize;
  IntWindow win(wrap, coordinateFromIndex(grid, flags[0]));
  for (size_t id : flags) win.expandToInclude(coordinateFromIndex(grid, id));
  const long n_user = n / (long) zoomfac;
  const Int3 s = win.sizes();
  const long n_required = std::max(s[0], std::max(s[1], s[2]));
  if (n_required > n_user)
    throw std::runtime_error("Zoom particles do not fit in specified sub-box. Decrease zoom, or choose different particles");
  if (n_required < n_user) win.expandSymmetricallyToSize(n_user);
  Int3 low;
  for (int a = 0; a < 3; ++a) low[a] = std::lround((double) grid.offsetLower[a] / (double) grid.cellSize);
  win.shift(low);
  IntWindow above(wrap, low);
  above.hi = {low[0] + n, low[1] + n, low[2] + n};
  above.wrapCorners();
  Int3 move{0, 0, 0};
  for (int a = 0; a < 3; ++a) {
    if (!IntWindow::within(above.lo[a], above.hi[a], win.lo[a])) move[a] = win.wrappedOffset(win.lo[a], above.lo[a]);
    if (!IntWindow::within(above.lo[a], above.hi[a], win.hi[a] - 1)) move[a] = win.wrappedOffset(win.hi[a], above.hi[a]);
  }
  win.shift(move);
  win.shift({-above.lo[0], -above.lo[1], -above.lo[2]});
  return win.lo;
}

// ICGenerator::initZoomGrid (ic.hpp:461-474): a flagged cell within `borderSafety` cells of the edge of the zoom window placed at
// `corner` (Window<int>::containsWithBorderSafety, window.hpp:296-302)
template <typename T> bool flagsNearZoomBorder(const grids::Grid<T> &grid, const Int3 &corner, size_t zoomfac, long borderSafety = 3) {
  const long n = (long) grid.size;
  IntWindow w((long) grid.simEquivalentSize, corner);
  w.hi = {corner[0] + n / (long) zoomfac, corner[1] + n / (long) zoomfac, corner[2] + n / (long) zoomfac};
  w.wrapCorners();
  std::vector<size_t> flags;
  grid.getFlaggedCells(flags);
  for (size_t id : flags) {
    const Int3 c = coordinateFromIndex(grid, id);
    for (int a = 0; a < 3; ++a) {
      const long lo = w.lo[a] + borderSafety, hi = w.hi[a] - borderSafety;
      const bool ok = hi >= lo ? (c[a] < hi && c[a] >= lo) : (c[a] < hi || c[a] >= lo);
      if (!ok) return true;
    }
  }
  return false;
}

// Grid::makeProxyGridToMatch(target)->getFlaggedCells (grid.hpp:195-228; SuperSampleGrid / SectionOfGrid / SubSampleGrid flag
// proxies, virtualgrid.hpp): the flagged cells of `source` as cells of `target` -- refined first if the target is finer,
// shifted into the target's frame (periodic in the simulation box, cells outside the target dropped), coarsened last
template <typename T> std::vector<size_t> proxyFlags(const grids::Grid<T> &source, const grids::Grid<T> &target) {
  auto ratio = [](double a, double b) {
    const double r = a / b;
    if (std::fabs(r - std::round(r)) > 1e-8 * std::max(1.0, std::fabs(r)) || std::round(r) < 1)
      throw std::runtime_error("The ratio of cell sizes is not a positive integer");
    return (long) std::lround(r);
  };
  std::vector<size_t> flags;
  source.getFlaggedCells(flags);
  long n = (long) source.size;
  double dx = (double) source.cellSize;
  std::vector<std::array<long, 3>> c;
  c.reserve(flags.size());
  for (size_t id : flags) c.push_back({(long) (id / (size_t) (n * n)), (long) ((id / (size_t) n) % (size_t) n), (long) (id % (size_t) n)});
  if ((double) target.cellSize < dx) {
    const long r = ratio(dx, (double) target.cellSize);
    std::vector<std::array<long, 3>> fine;
    fine.reserve(c.size() * (size_t) (r * r * r));
    for (auto &q : c)
      for (long a = 0; a < r; ++a)
        for (long b = 0; b < r; ++b)
          for (long d = 0; d < r; ++d) fine.push_back({q[0] * r + a, q[1] * r + b, q[2] * r + d});
    c.swap(fine);
    dx /= (double) r;
    n *= r;
  }
  const long sub = ratio((double) target.cellSize, dx), pre = (long) target.size * sub;
  long off[3];
  bool shifted = false;
  for (int a = 0; a < 3; ++a) {
    const double r = ((double) target.offsetLower[a] - (double) source.offsetLower[a]) / dx;
    if (std::fabs(r - std::round(r)) > 1e-6 * std::max(1.0, std::fabs(r))) throw std::runtime_error("Grid offsets do not differ by a whole number of cells");
    off[a] = (long) std::lround(r);
    shifted = shifted || off[a] != 0;
  }
  const long wrap = std::lround((double) source.periodicDomainSize / dx), m = (long) target.size;
  std::vector<size_t> out;
  for (auto q : c) {
    bool keep = true;
    if (shifted || pre != n)
      for (int a = 0; a < 3; ++a) {
        q[a] = ((q[a] - off[a]) % wrap + wrap) % wrap;
        keep = keep && q[a] < pre;
      }
    if (keep) out.push_back((size_t) (((q[0] / sub) * m + (q[1] / sub)) * m + (q[2] / sub)));
  }
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
  return out;
}

// the trimming of initZoomGridWithLowLeftCornerAt (ic.hpp:531-554): flagged cells whose centroid lies outside the new window
template <typename T>
std::vector<size_t> trimToZoomWindow(const grids::Grid<T> &up, const std::vector<size_t> &flags, const Int3 &corner, long nCoarse) {
  const long n = (long) up.size;
  Int3 hiCell{corner[0] + nCoarse, corner[1] + nCoarse, corner[2] + nCoarse};
  if (up.coversFullSimulation()) { for (int a = 0; a < 3; ++a) hiCell[a] = ((hiCell[a] - 1) % n + n) % n + 1; }
  else if (std::max(hiCell[0], std::max(hiCell[1], hiCell[2])) > n)
    throw std::runtime_error("The newly initialised zoom level is not fully contained within the previous level");
  const T eps = up.cellSize * T(1e-6), L = up.periodicDomainSize;
  auto wrap = [L](T v) { if (v < 0) v += L; if (v >= L) v -= L; return v; };
  T lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = wrap(centroid(up, a, corner[a]) - eps);
    const T h = centroid(up, a, hiCell[a] - 1) + eps;
    hi[a] = std::nextafter(wrap(std::nextafter(h, -std::numeric_limits<T>::infinity())), std::numeric_limits<T>::infinity());
  }
  std::vector<size_t> kept;
  for (size_t id : flags) {
    const Int3 c = coordinateFromIndex(up, id);
    bool in = true;
    for (int a = 0; a < 3 && in; ++a) {
      const T x = centroid(up, a, c[a]);
      in = hi[a] >= lo[a] ? (x < hi[a] && x >= lo[a]) : (x < hi[a] || x >= lo[a]);
    }
    if (in) kept.push_back(id);
  }
  return kept;
}
}  // namespace zoom

// ---------------------------------------------------------------------------------------------------------------------
// log output with the indentation of logging::IndentWhileInScope (the geometry run of `mapper_relative_to` is shown as "| ...")
inline std::string &logIndent() { static std::string s; return s; }
inline void logLine(const std::string &text) {
  size_t a = 0;
  while (true) {
    const size_t b = text.find('\n', a);
    std::cout << logIndent() << text.substr(a, b == std::string::npos ? std::string::npos : b - a) << std::endl;
    if (b == std::string::npos) break;
    a = b + 1;
  }
}

namespace particle { namespace mapper {

template <typename T> class OneLevelParticleMapper;

// (layer, cell) of every particle in output order; a layer is one OneLevelParticleMapper
struct Layout {
  std::vector<int32_t> layer;
  std::vector<int64_t> cell;
  size_t size() const { return layer.size(); }
};

template <typename T> class ParticleMapper {
public:
  using Ptr = std::shared_ptr<ParticleMapper<T>>;
  virtual ~ParticleMapper() = default;
  virtual size_t size() const = 0;
  virtual size_t size_gas() const { return 0; }
  virtual grids::Grid<T> &getFinestGrid() = 0;
  virtual int getFinestLayer() const = 0;
  virtual const Layout &layout() const = 0;
  virtual void collectLayers(std::map<int, const OneLevelParticleMapper<T> *> &out) const = 0;
  virtual void unflagAllParticles() = 0;
  virtual void flagParticles(const std::vector<size_t> &orderedParticleIndices) = 0;
  virtual void getFlaggedParticles(std::vector<size_t> &out) const = 0;
  virtual std::pair<Ptr, Ptr> splitMass(T massRatio, const std::vector<const grids::Grid<T> *> &toGrids) = 0;
  std::map<int, const OneLevelParticleMapper<T> *> layers() const { std::map<int, const OneLevelParticleMapper<T> *> m; collectLayers(m); return m; }
};

inline int nextLayerId() { static int id = 0; return id++; }

template <typename T> class OneLevelParticleMapper : public ParticleMapper<T> {  // onelevelmapper.hpp
public:
  using Ptr = typename ParticleMapper<T>::Ptr;
  grids::Grid<T> *grid;
  size_t level;
  unsigned gadgetParticleType;
  bool massScaled;
  T massScale;  // MassScaledGrid (virtualgrid.hpp:657-676)
  int layer;
  OneLevelParticleMapper(grids::Grid<T> &g, size_t level_, unsigned type = 1, bool scaled = false, T scale = 1)
      : grid(&g), level(level_), gadgetParticleType(type), massScaled(scaled), massScale(scale), layer(nextLayerId()) {
    lay.layer.assign(g.size3, layer);
    lay.cell.resize(g.size3);
    std::iota(lay.cell.begin(), lay.cell.end(), (int64_t) 0);
  }
  size_t size() const override { return grid->size3; }
  grids::Grid<T> &getFinestGrid() override { return *grid; }
  int getFinestLayer() const override { return layer; }
  const Layout &layout() const override { return lay; }
  void collectLayers(std::map<int, const OneLevelParticleMapper<T> *> &out) const override { out[layer] = this; }
  void unflagAllParticles() override { grid->flagCells({}); }
  void flagParticles(const std::vector<size_t> &ids) override {  // Grid::flagCells is a union in the reference (grid.hpp:269-275)
    std::vector<size_t> cur;
    grid->getFlaggedCells(cur);
    for (size_t i : ids) { if (i >= grid->size3) throw std::runtime_error("Particle ID out of range"); cur.push_back(i); }
    std::sort(cur.begin(), cur.end());
    cur.erase(std::unique(cur.begin(), cur.end()), cur.end());
    grid->flagCells(cur);
  }
  void getFlaggedParticles(std::vector<size_t> &out) const override {
    std::vector<size_t> f;
    grid->getFlaggedCells(f);
    std::sort(f.begin(), f.end());
    out.insert(out.end(), f.begin(), f.end());
  }
  std::pair<Ptr, Ptr> splitMass(T massRatio, const std::vector<const grids::Grid<T> *> &toGrids) override {  // :155-173
    if (std::find(toGrids.begin(), toGrids.end(), grid) != toGrids.end())
      return {std::make_shared<OneLevelParticleMapper<T>>(*grid, level, 0u, true, massScaled ? massScale * massRatio : massRatio),
              std::make_shared<OneLevelParticleMapper<T>>(*grid, level, gadgetParticleType, true,
                                                          massScaled ? massScale * (T(1.0) - massRatio) : T(1.0) - massRatio)};
    return {nullptr, std::make_shared<OneLevelParticleMapper<T>>(*grid, level, gadgetParticleType, massScaled, massScale)};
  }
  T cellMassFrac() const {  // grid.hpp:112, virtualgrid.hpp:672
    const T frac = T(std::pow((double) T(grid->cellSize / grid->periodicDomainSize), 3.0));
    return massScaled ? massScale * frac : frac;
  }
private:
  Layout lay;
};

// Grid::insertCubeIdsIntoVector (grid.hpp:704-731): ids of the cells of `fine` whose centres lie in the cube of side dxc
// centred on x0c, x outermost
template <typename T>
void insertCubeIds(const grids::Grid<T> &fine, const T x0c[3], T dxc, std::vector<int64_t> &out) {
  long a[3], b[3];
  for (int d = 0; d < 3; ++d) {
    T x = std::fmod(x0c[d] - fine.offsetLower[d], fine.periodicDomainSize);
    if (x < 0) x += fine.periodicDomainSize;
    a[d] = (long) std::floor((x - dxc / 2 + fine.cellSize / 2) / fine.cellSize);
    b[d] = (long) std::floor((x + dxc / 2 - fine.cellSize / 2) / fine.cellSize);
    if (a[d] < 0 || b[d] >= (long) fine.size) throw std::out_of_range("Requested zoom region falls outside high resolution grid");
  }
  const long n = (long) fine.size;
  for (long x = a[0]; x <= b[0]; ++x)
    for (long y = a[1]; y <= b[1]; ++y)
      for (long z = a[2]; z <= b[2]; ++z) out.push_back((int64_t) ((x * n + y) * n + z));
}

template <typename T> class TwoLevelParticleMapper : public ParticleMapper<T> {  // twolevelmapper.hpp
public:
  using Ptr = typename ParticleMapper<T>::Ptr;
  Ptr pLevel1;
  std::shared_ptr<OneLevelParticleMapper<T>> pLevel2;
  bool skipLevel1;
  std::vector<size_t> level1CellsToReplace, level1ParticlesToReplace;
  std::vector<int64_t> zoomParticleArrayHiresUnsorted;
  size_t n_hr_per_lr = 0, firstLevel2Particle = 0, totalParticles = 0;

  TwoLevelParticleMapper(Ptr l1, std::shared_ptr<OneLevelParticleMapper<T>> l2, std::vector<size_t> cells, bool skip = false)
      : pLevel1(l1), pLevel2(l2), skipLevel1(skip), level1CellsToReplace(std::move(cells)) {
    if (level1CellsToReplace.empty()) throw std::runtime_error("TwoLevelParticleMapper: no cells to replace");
    std::sort(level1CellsToReplace.begin(), level1CellsToReplace.end());
    level1CellsToReplace.erase(std::unique(level1CellsToReplace.begin(), level1CellsToReplace.end()), level1CellsToReplace.end());
    const Layout &L1 = pLevel1->layout();
    const int finest = pLevel1->getFinestLayer();
    std::vector<char> keep(L1.size(), skip ? 0 : 1);
    for (size_t i = 0; i < L1.size(); ++i)
      if (L1.layer[i] == finest && std::binary_search(level1CellsToReplace.begin(), level1CellsToReplace.end(), (size_t) L1.cell[i])) {
        level1ParticlesToReplace.push_back(i);
        keep[i] = 0;
      }
    if (level1ParticlesToReplace.size() != level1CellsToReplace.size())
      throw std::runtime_error("Consistency error when relating cells to particles");
    grids::Grid<T> &g1 = pLevel1->getFinestGrid();
    const grids::Grid<T> &g2 = *pLevel2->grid;
    const double r = (double) g1.cellSize / (double) g2.cellSize;
    if (std::abs(r - std::round(r)) > 1e-8 || std::round(r) < 1)
      throw std::runtime_error("TwoLevelParticleMapper: the cell size ratio is not a positive integer");
    const size_t ratio = (size_t) std::llround(r);
    n_hr_per_lr = ratio * ratio * ratio;
    firstLevel2Particle = skip ? 0 : pLevel1->size() - level1ParticlesToReplace.size();
    totalParticles = firstLevel2Particle + n_hr_per_lr * level1ParticlesToReplace.size();
    // calculateHiresParticleList (:866-893): blocks in ascending coarse cell id
    for (size_t id : level1CellsToReplace) {
      const zoom::Int3 c = zoom::coordinateFromIndex(g1, id);
      const T cen[3] = {zoom::centroid(g1, 0, c[0]), zoom::centroid(g1, 1, c[1]), zoom::centroid(g1, 2, c[2])};
      const size_t before = zoomParticleArrayHiresUnsorted.size();
      insertCubeIds(g2, cen, g1.cellSize, zoomParticleArrayHiresUnsorted);
      if (zoomParticleArrayHiresUnsorted.size() - before != n_hr_per_lr)
        throw std::runtime_error("Zoomed cells do not hold n_hr_per_lr fine cells");
    }
    for (size_t i = 0; i < L1.size(); ++i)
      if (keep[i]) { lay.layer.push_back(L1.layer[i]); lay.cell.push_back(L1.cell[i]); }
    for (int64_t c : zoomParticleArrayHiresUnsorted) { lay.layer.push_back(pLevel2->layer); lay.cell.push_back(c); }
  }
  size_t size() const override { return totalParticles; }
  grids::Grid<T> &getFinestGrid() override { return *pLevel2->grid; }
  int getFinestLayer() const override { return pLevel2->layer; }
  const Layout &layout() const override { return lay; }
  void collectLayers(std::map<int, const OneLevelParticleMapper<T> *> &out) const override {
    pLevel1->collectLayers(out);
    pLevel2->collectLayers(out);
  }
  void unflagAllParticles() override { pLevel1->unflagAllParticles(); pLevel2->unflagAllParticles(); }

  void flagParticles(const std::vector<size_t> &ids) override {  // :358-452
    const size_t nZoom = level1ParticlesToReplace.size();
    if (pLevel1->size() < nZoom) throw std::runtime_error("Zoom particle list is longer than the grid it refers to");
    const size_t firstHires = pLevel1->size() - nZoom;
    std::vector<size_t> level1, level2;
    size_t last = 0, skipped = 0;
    for (size_t id : ids) {
      if (id < last) throw std::runtime_error("The particle list must be in ascending order");
      last = id;
      if (id < firstHires) {
        while (skipped < nZoom && level1ParticlesToReplace[skipped] <= id + skipped) ++skipped;
        level1.push_back(id + skipped);
      } else {
        const size_t lr = (id - firstHires) / n_hr_per_lr;
        if (lr >= nZoom) throw std::runtime_error("Particle ID out of range");
        level1.push_back(level1ParticlesToReplace[lr]);
        level2.push_back((size_t) zoomParticleArrayHiresUnsorted[id - firstHires]);
      }
    }
    std::sort(level1.begin(), level1.end());
    std::sort(level2.begin(), level2.end());
    pLevel1->flagParticles(level1);
    pLevel2->flagParticles(level2);
  }

  void getFlaggedParticles(std::vector<size_t> &out) const override {  // :455-494
    std::vector<size_t> g1, g2, res;
    pLevel1->getFlaggedParticles(g1);
    size_t zi = 0;
    for (size_t i : g1) {
      while (zi < level1ParticlesToReplace.size() && level1ParticlesToReplace[zi] < i) ++zi;
      if (zi == level1ParticlesToReplace.size() || level1ParticlesToReplace[zi] != i) res.push_back(i - zi);
    }
    pLevel2->getFlaggedParticles(g2);
    std::vector<size_t> order(zoomParticleArrayHiresUnsorted.size());
    std::iota(order.begin(), order.end(), (size_t) 0);
    std::stable_sort(order.begin(), order.end(), [this](size_t a, size_t b) { return zoomParticleArrayHiresUnsorted[a] < zoomParticleArrayHiresUnsorted[b]; });
    size_t k = 0;
    for (size_t c : g2) {
      while (k < order.size() && (size_t) zoomParticleArrayHiresUnsorted[order[k]] < c) ++k;
      if (k == order.size() || (size_t) zoomParticleArrayHiresUnsorted[order[k]] != c) continue;
      res.push_back(order[k] + firstLevel2Particle);
    }
    std::sort(res.begin(), res.end());
    out.insert(out.end(), res.begin(), res.end());
  }

  std::pair<Ptr, Ptr> splitMass(T massRatio, const std::vector<const grids::Grid<T> *> &toGrids) override {  // :542-566
    auto n1 = pLevel1->splitMass(massRatio, toGrids);
    auto n2 = pLevel2->splitMass(massRatio, toGrids);
    bool newskip = skipLevel1;
    Ptr gas1 = n1.first;
    if (!gas1) { gas1 = pLevel1; newskip = true; }
    Ptr gas;
    if (n2.first)
      gas = std::make_shared<TwoLevelParticleMapper<T>>(gas1, std::static_pointer_cast<OneLevelParticleMapper<T>>(n2.first), level1CellsToReplace, newskip);
    Ptr dm = std::make_shared<TwoLevelParticleMapper<T>>(n1.second, std::static_pointer_cast<OneLevelParticleMapper<T>>(n2.second), level1CellsToReplace, skipLevel1);
    return {gas, dm};
  }
private:
  Layout lay;
};

template <typename T> class AddGasMapper : public ParticleMapper<T> {  // gasmapper.hpp
public:
  using Ptr = typename ParticleMapper<T>::Ptr;
  Ptr firstMap, secondMap;
  bool gasFirst;
  size_t nFirst, nSecond;
  AddGasMapper(Ptr first, Ptr second, bool gasFirst_) : firstMap(first), secondMap(second), gasFirst(gasFirst_), nFirst(first->size()), nSecond(second->size()) {
    lay = firstMap->layout();
    const Layout &b = secondMap->layout();
    lay.layer.insert(lay.layer.end(), b.layer.begin(), b.layer.end());
    lay.cell.insert(lay.cell.end(), b.cell.begin(), b.cell.end());
    for (auto &kv : (gasFirst ? secondMap : firstMap)->layers())
      if (kv.second->gadgetParticleType == 0)
        throw std::runtime_error("You have asked for some gadget DM particles to be output with type 0, but have also set non-zero baryon density. Type 0 is reserved for gas particles in hydro simulations, and genetIC particle types refer only to the dark matter part of the output.");
  }
  size_t size() const override { return nFirst + nSecond; }
  size_t size_gas() const override { return gasFirst ? nFirst : nSecond; }
  grids::Grid<T> &getFinestGrid() override { return (gasFirst ? secondMap : firstMap)->getFinestGrid(); }
  int getFinestLayer() const override { return (gasFirst ? secondMap : firstMap)->getFinestLayer(); }
  const Layout &layout() const override { return lay; }
  void collectLayers(std::map<int, const OneLevelParticleMapper<T> *> &out) const override { firstMap->collectLayers(out); secondMap->collectLayers(out); }
  void unflagAllParticles() override { firstMap->unflagAllParticles(); secondMap->unflagAllParticles(); }
  void flagParticles(const std::vector<size_t> &ids) override {  // ids refer to the dark-matter part only (:149-166)
    std::vector<size_t> part;
    for (size_t i : ids) { if (gasFirst) { if (i >= nFirst) part.push_back(i - nFirst); } else if (i < nFirst) part.push_back(i); }
    (gasFirst ? secondMap : firstMap)->flagParticles(part);
  }
  void getFlaggedParticles(std::vector<size_t> &out) const override {
    std::vector<size_t> r;
    (gasFirst ? secondMap : firstMap)->getFlaggedParticles(r);
    for (size_t i : r) out.push_back(gasFirst ? i + nFirst : i);
  }
  std::pair<Ptr, Ptr> splitMass(T, const std::vector<const grids::Grid<T> *> &) override { throw std::runtime_error("AddGasMapper cannot be split again"); }
  bool isGas(size_t i) const { return gasFirst ? i < nFirst : i >= nFirst; }
private:
  Layout lay;
};

// ICGenerator::updateParticleMapper (ic.hpp:1129-1221) without sub/supersampling and padding
template <typename T>
typename ParticleMapper<T>::Ptr updateParticleMapper(multilevelgrid::MultiLevelGrid<T> &ctx, const std::vector<std::vector<size_t>> &zoomParticleArray,
                                                    const std::vector<unsigned> &gadgetTypes, const cosmology::CosmologicalParameters<T> &cosmo,
                                                    bool baryonsOnAllLevels, bool gasFirst, int flaggedGadgetParticleType = -1,
                                                    const std::vector<size_t> *flaggedCells = nullptr) {
  const size_t nl = ctx.getNumLevels();
  typename ParticleMapper<T>::Ptr m = std::make_shared<OneLevelParticleMapper<T>>(ctx.getGridForLevel(0), 0, gadgetTypes[0]);
  for (size_t level = 1; level < nl; ++level) {
    auto fine = std::make_shared<OneLevelParticleMapper<T>>(ctx.getGridForLevel(level), level, gadgetTypes[level]);
    m = std::make_shared<TwoLevelParticleMapper<T>>(m, fine, zoomParticleArray[level - 1]);
  }
  if (flaggedGadgetParticleType >= 0 && flaggedCells) {  // ic.hpp:1176-1186: the flagged cells of the finest grid, retyped
    auto flagged = std::make_shared<OneLevelParticleMapper<T>>(ctx.getGridForLevel(nl - 1), nl - 1, (unsigned) flaggedGadgetParticleType);
    m = std::make_shared<TwoLevelParticleMapper<T>>(m, flagged, *flaggedCells);
  }
  if (cosmo.OmegaBaryons0 > 0) {
    std::vector<const grids::Grid<T> *> to;
    if (baryonsOnAllLevels) for (size_t l = 0; l < nl; ++l) to.push_back(&ctx.getGridForLevel(l));
    else to.push_back(&ctx.getGridForLevel(nl - 1));
    auto split = m->splitMass(cosmo.OmegaBaryons0 / cosmo.OmegaM0, to);
    m = gasFirst ? std::make_shared<AddGasMapper<T>>(split.first, split.second, true)
                 : std::make_shared<AddGasMapper<T>>(split.second, split.first, false);
  }
  return m;
}
}}  // namespace particle::mapper

// ---------------------------------------------------------------------------------------------------------------------
namespace particle {
// initialiseParticleGeneratorBasedOnTemplate (multilevelgenerator.hpp:79-128): per level the three Zeldovich offset fields
// in Fourier space, high-pass filtered and completed with the low-pass filtered, tricubically interpolated level above
template <typename T> class MultiLevelParticleGenerator {
public:
  using TField = fields::Field<T>;
  MultiLevelParticleGenerator(multilevelgrid::MultiLevelGrid<T> &ctx, std::vector<std::shared_ptr<TField>> &overdensity) {
    const size_t nl = ctx.getNumLevels();
    pOff.resize(nl);
    for (size_t l = 0; l < nl; ++l) {
      auto &d = *overdensity[l];
      d.toFourier();
      const auto &g = d.getGrid();
      for (int axis = 0; axis < 3; ++axis) {
        auto psi = std::make_shared<TField>(g, true);
        check(gx_zeldovich_fourier(d.devicePointer(), psi->devicePointer(), (int) g.size, dtype_of<T>(), (double) g.thisGridSize, axis, nullptr));
        pOff[l].push_back(psi);
      }
    }
    if (nl >= 2) {
      filters::FilterFamily<T> fam(ctx);
      for (size_t l = 1; l < nl; ++l) {
        const auto &hp = fam.getHighPassFilterForLevel(l);
        const auto &lp = fam.getLowPassFilterForLevel(l - 1);
        overdensity[l]->applyFilter(hp);
        for (auto &p : pOff[l]) p->applyFilter(hp);
        overdensity[l]->addFieldFromDifferentGridWithFilter(*overdensity[l - 1], lp);
        for (int a = 0; a < 3; ++a) pOff[l][a]->addFieldFromDifferentGridWithFilter(*pOff[l - 1][a], lp);
      }
      ctx.setLevelsAreCombined();
    }
    for (auto &lev : pOff) for (auto &p : lev) p->toReal();
  }
  std::vector<std::shared_ptr<TField>> &getGeneratedFieldsForLevel(size_t l) { return pOff[l]; }
private:
  std::vector<std::vector<std::shared_ptr<TField>>> pOff;
};

// every particle of a mapper, evaluated layer by layer on the device at the layer's cell list (gx_emit_cells) and placed in
// mapper order on the host; O = float for gadget (Coordinate<float>(particle.pos)), T for tipsy (converted in double first)
template <typename T, typename O> struct MappedParticles {
  std::vector<O> pos, vel;          // [N][3]
  std::vector<T> mass, eps;
  std::vector<int32_t> gadgetType;
  std::vector<char> isGas;
  MappedParticles(const mapper::ParticleMapper<T> &m, MultiLevelParticleGenerator<T> *gen, const cosmology::CosmologicalParameters<T> &cosmo,
                  T epsNorm = T(0.01075)) {
    const mapper::Layout &lay = m.layout();
    const size_t N = lay.size();
    mass.resize(N); eps.resize(N); gadgetType.resize(N); isGas.assign(N, 0);
    if (auto *g = dynamic_cast<const mapper::AddGasMapper<T> *>(&m)) for (size_t i = 0; i < N; ++i) isGas[i] = g->isGas(i);
    if (gen) { pos.resize(3 * N); vel.resize(3 * N); }
    const T ratio = cosmology::zeldovichVelocityToOffsetRatio(cosmo);
    for (auto &kv : m.layers()) {
      const auto *one = kv.second;
      const grids::Grid<T> &g = *one->grid;
      const T boxMass = T(T(27.744948) * cosmo.OmegaM0) * T(std::pow((double) g.periodicDomainSize, 3.0));   // zeldovich.hpp:55-62
      const T pmass = boxMass * one->cellMassFrac();
      const T peps = T(g.cellSize * T(1.0)) * epsNorm;
      std::vector<size_t> where;
      std::vector<long long> cells;
      for (size_t i = 0; i < N; ++i)
        if (lay.layer[i] == kv.first) { where.push_back(i); cells.push_back((long long) lay.cell[i]); mass[i] = pmass; eps[i] = peps; gadgetType[i] = (int32_t) one->gadgetParticleType; }
      if (!gen || where.empty()) continue;
      const size_t cnt = where.size();
      void *dcells = nullptr, *dpos = nullptr, *dvel = nullptr;
      check(gx_malloc(&dcells, cnt * sizeof(long long)));
      check(gx_malloc(&dpos, 3 * cnt * sizeof(O)));
      check(gx_malloc(&dvel, 3 * cnt * sizeof(O)));
      check(gx_memcpy_h2d(dcells, cells.data(), cnt * sizeof(long long), nullptr));
      auto &psi = gen->getGeneratedFieldsForLevel(one->level);
      for (int axis = 0; axis < 3; ++axis)
        check(gx_emit_cells(psi[axis]->devicePointer(), (int) g.size, dtype_of<T>(), axis, (double) g.cellSize, (double) g.offsetLower[axis],
                            (double) g.periodicDomainSize, (double) ratio, (const long long *) dcells, (long long) cnt,
                            (char *) dpos + axis * sizeof(O), (char *) dvel + axis * sizeof(O), dtype_of<O>(), 3, nullptr));
      std::vector<O> p(3 * cnt), v(3 * cnt);
      check(gx_memcpy_d2h(p.data(), dpos, p.size() * sizeof(O), nullptr));
      check(gx_memcpy_d2h(v.data(), dvel, v.size() * sizeof(O), nullptr));
      check(gx_synchronize(nullptr));
      gx_free(dcells); gx_free(dpos); gx_free(dvel);
      for (size_t k = 0; k < cnt; ++k)
        for (int a = 0; a < 3; ++a) { pos[3 * where[k] + a] = p[3 * k + a]; vel[3 * where[k] + a] = v[3 * k + a]; }
    }
  }
};
}  // namespace particle

// ---------------------------------------------------------------------------------------------------------------------
namespace io {
namespace tipsy {
// TipsyOutput::operator() with a mapper (tipsy.hpp:207-271): gas records first, then dark; tipsy.param next to the file
template <typename T>
void saveMapped(const std::string &filename, double boxLength, const particle::MappedParticles<T, T> &P, const cosmology::CosmologicalParameters<T> &cosmo) {
  const size_t N = P.mass.size();
  T tot_mass = 0;
  for (size_t i = 0; i < N; ++i) tot_mass += P.mass[i];
  const double mass_factor = cosmo.OmegaM0 / tot_mass, pos_factor = 1. / boxLength;
  const double dKpcUnit = boxLength * 1000 / cosmo.hubble, dMsolUnit = 1e10 / cosmo.hubble / mass_factor;
  const double dKmsUnit = std::sqrt(4.3022682e-6 * dMsolUnit / dKpcUnit);
  const double vel_factor = std::pow((double) cosmo.scalefactor, -0.5) / dKmsUnit;
  size_t ngas = 0;
  for (size_t i = 0; i < N; ++i) ngas += P.isGas[i] ? 1 : 0;
  struct { double scalefactor; int n, ndim, ngas, ndark, nstar, pad; } header;
  header.scalefactor = cosmo.scalefactor; header.n = (int) N; header.ndim = 3; header.ngas = (int) ngas; header.ndark = (int) (N - ngas);
  header.nstar = 0; header.pad = 0;
  std::ofstream f(filename, std::ios::binary);
  if (!f.is_open()) throw std::runtime_error("Could not open " + filename + " for writing");
  f.write(reinterpret_cast<const char *>(&header), sizeof(header));
  const double a = cosmo.scalefactor, adec = 1. / 150.;
  const float temp = (float) (a > adec ? cosmo.TCMB * adec / (a * a) : cosmo.TCMB / a);   // cosmology::getTemperature
  for (size_t i = 0; i < N; ++i) {
    float rec[12] = {0};
    rec[0] = (float) (P.mass[i] * mass_factor);
    for (int d = 0; d < 3; ++d) { rec[1 + d] = (float) (P.pos[3 * i + d] * pos_factor - 0.5); rec[4 + d] = (float) (P.vel[3 * i + d] * vel_factor); }
    if (P.isGas[i]) { rec[8] = temp; rec[9] = (float) (P.eps[i] * pos_factor); f.write(reinterpret_cast<const char *>(rec), 12 * sizeof(float)); }
    else { rec[7] = (float) (P.eps[i] * pos_factor); f.write(reinterpret_cast<const char *>(rec), 9 * sizeof(float)); }
  }
}
}  // namespace tipsy

namespace gadget {
// gadget::save<float> (gadget.hpp:340-556): Fortran records header / pos / vel / ids / [mass] / [u], by gadget type
template <typename T>
void save(const std::string &name, double boxLength, const particle::MappedParticles<T, float> &P, const cosmology::CosmologicalParameters<T> &cosmo,
          int gadgetVersion, int nFiles = 1) {
  if (gadgetVersion != 2 && gadgetVersion != 3) throw std::runtime_error("Unknown gadget format");
  const size_t N = P.mass.size();
  std::vector<size_t> order(N);
  std::iota(order.begin(), order.end(), (size_t) 0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return P.gadgetType[a] < P.gadgetType[b]; });
  size_t nPartPerType[6] = {0};
  double masses[6] = {0};
  T minMass[6], maxMass[6];
  for (int t = 0; t < 6; ++t) { minMass[t] = std::numeric_limits<T>::max(); maxMass[t] = 0; }
  for (size_t i = 0; i < N; ++i) {
    const int t = P.gadgetType[i];
    nPartPerType[t]++;
    minMass[t] = std::min(minMass[t], P.mass[i]);
    maxMass[t] = std::max(maxMass[t], P.mass[i]);
  }
  bool variableMass = false;
  for (int t = 0; t < 6; ++t) if (nPartPerType[t]) { if (minMass[t] != maxMass[t]) variableMass = true; masses[t] = (double) minMass[t]; }
  if (variableMass) for (int t = 0; t < 6; ++t) masses[t] = 0;
  std::vector<size_t> remaining(nPartPerType, nPartPerType + 6);
  size_t start = 0;
  auto record = [](std::ofstream &f, const void *data, size_t bytes) {
    const int marker = bytes > (size_t) std::numeric_limits<int>::max() ? std::numeric_limits<int>::max() : (int) bytes;
    f.write(reinterpret_cast<const char *>(&marker), 4);
    f.write(reinterpret_cast<const char *>(data), (std::streamsize) bytes);
    f.write(reinterpret_cast<const char *>(&marker), 4);
  };
  for (int file = 0; file < nFiles; ++file) {
    const size_t nThis = file == nFiles - 1 ? N - (N / nFiles) * (nFiles - 1) : N / nFiles;
    size_t room = nThis, thisFile[6];
    for (int t = 0; t < 6; ++t) { thisFile[t] = std::min(remaining[t], room); room -= thisFile[t]; remaining[t] -= thisFile[t]; }
    unsigned char h[256];
    std::memset(h, 0, sizeof(h));
    auto put = [&](size_t off, const void *v, size_t b) { std::memcpy(h + off, v, b); };
    for (int t = 0; t < 6; ++t) {
      const int np = (int) thisFile[t];
      const unsigned lo = (unsigned) (nPartPerType[t] & 0xffffffffu), hi = (unsigned) (nPartPerType[t] >> 32);
      put(4 * t, &np, 4); put(24 + 8 * t, &masses[t], 8); put(96 + 4 * t, &lo, 4); put(168 + 4 * t, &hi, 4);
    }
    const double time = cosmo.scalefactor, z = cosmo.redshift, box = boxLength, om = cosmo.OmegaM0, ol = cosmo.OmegaLambda0, hub = cosmo.hubble;
    const int baryons = nPartPerType[0] > 0 ? 1 : 0, one = 1, zero = 0;
    put(72, &time, 8); put(80, &z, 8); put(88, &baryons, 4); put(92, &baryons, 4); put(120, &baryons, 4); put(124, &nFiles, 4);
    put(128, &box, 8); put(136, &om, 8); put(144, &ol, 8); put(152, &hub, 8);
    if (gadgetVersion == 3) { put(196, &zero, 4); put(200, &one, 4); }   // flag_doubleprecision = 0 (float output), flag_ic_info = 1
    std::string fn = name + std::to_string(gadgetVersion) + (nFiles == 1 ? "" : "." + std::to_string(file));
    std::ofstream f(fn, std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("Could not open " + fn + " for writing");
    record(f, h, 256);
    std::vector<float> buf(3 * nThis);
    for (size_t k = 0; k < nThis; ++k) for (int d = 0; d < 3; ++d) buf[3 * k + d] = P.pos[3 * order[start + k] + d];
    record(f, buf.data(), buf.size() * 4);
    for (size_t k = 0; k < nThis; ++k) for (int d = 0; d < 3; ++d) buf[3 * k + d] = P.vel[3 * order[start + k] + d];
    record(f, buf.data(), buf.size() * 4);
    std::vector<unsigned long> ids(nThis);
    for (size_t k = 0; k < nThis; ++k) ids[k] = (unsigned long) order[start + k];
    record(f, ids.data(), ids.size() * sizeof(unsigned long));
    if (variableMass) {
      std::vector<float> m(nThis);
      for (size_t k = 0; k < nThis; ++k) m[k] = (float) P.mass[order[start + k]];
      record(f, m.data(), m.size() * 4);
    }
    if (cosmo.OmegaBaryons0 > 0) { std::vector<float> u(thisFile[0], 0.0f); record(f, u.data(), u.size() * 4); }
    start += nThis;
  }
}
}  // namespace gadget
}  // namespace io

// ---------------------------------------------------------------------------------------------------------------------
namespace ic {
// ICGenerator with zoom levels: the commands of a paramfile that define the geometry (base_grid, zoom_grid[_with_origin_at],
// centre, select_sphere / cube / nearest, id_file, merge_id_file, dump_id_file), seeds, cosmology, gadget particle types, gas
// (Ob, baryons_all_levels), output (tipsy, gadget2, gadget3) and `done`.
template <typename T> class ZoomICGenerator {
public:
  cosmology::CosmologicalParameters<T> cosmo;
  multilevelgrid::MultiLevelGrid<T> context;
  std::vector<std::vector<size_t>> zoomParticleArray;
  std::vector<unsigned> gadgetTypesForLevels;
  unsigned mruGadgetType = 1;
  bool usedGadgetParticleTypes[6] = {false, false, false, false, false, false};
  ZoomICGenerator<T> *pUnderlying = nullptr;   // DummyICGenerator (dummyic.hpp): geometry only, grids shared with this generator
  int flaggedGadgetParticleType = -1;          // gadget_flagged_particle_type (ic.hpp:703-714)
  std::vector<size_t> flaggedCellsForType;
  std::array<T, 3> velOffset{0, 0, 0};   // velocity_offset (ic.hpp:285-289), km a^1/2 s^-1
  std::string cambFile, outputFolder = ".", outputFilename = "IC", outputFormat, dir = ".";
  unsigned long seed = 0;
  int noiseMode = -1, nGadgetFiles = 1;
  bool reverse = false, fixPower = false, baryonsOnAllLevels = false;
  bool layoutOnly = false;               // stop before anything touches the device (CPU check of the host logic)
  std::string particlesFile;             // CPU check of the writers: positions / velocities [N][3] doubles each, in mapper order
  std::array<T, 3> centre{0, 0, 0};
  typename particle::mapper::ParticleMapper<T>::Ptr pMapper;

  void setRedshift(T z) { cosmo.redshift = z; cosmo.scalefactor = T(1.) / (z + T(1.)); }

  typename particle::mapper::ParticleMapper<T>::Ptr mapper() {
    return particle::mapper::updateParticleMapper<T>(context, zoomParticleArray, gadgetTypesForLevels, cosmo, baryonsOnAllLevels, outputFormat == "tipsy",
                                                     flaggedGadgetParticleType, &flaggedCellsForType);
  }

  // DummyICGenerator::addLevelToContext (dummyic.hpp:28-58): a level the underlying generator also has is the same Grid object,
  // so that flags set through either setup are seen by both
  void levelAdded() {
    if (!pUnderlying) return;
    const size_t level = context.getNumLevels() - 1;
    gadgetTypesForLevels.back() = 1;
    if (pUnderlying->context.getNumLevels() <= level) return;
    auto shared = pUnderlying->context.getGridPtr(level);
    const auto &mine = context.getGridForLevel(level);
    if (shared->size != mine.size) throw std::runtime_error("Trying to match particles between incompatible simulation setups (wrong grid n)");
    if (shared->thisGridSize != mine.thisGridSize)
      throw std::runtime_error("Trying to match particles between incompatible simulation setups (wrong grid size)");
    for (int a = 0; a < 3; ++a)
      if (std::fabs((double) shared->offsetLower[a] - (double) mine.offsetLower[a]) > 1e-12 * std::max(1.0, std::fabs((double) shared->offsetLower[a])))
        throw std::runtime_error("Trying to match particles between incompatible simulation setups (wrong grid origin)");
    {  // the flags this setup put on its own copy of the level are kept (dummyic.hpp:50-52)
      std::vector<size_t> have, add;
      shared->getFlaggedCells(have);
      mine.getFlaggedCells(add);
      have.insert(have.end(), add.begin(), add.end());
      std::sort(have.begin(), have.end());
      have.erase(std::unique(have.begin(), have.end()), have.end());
      shared->flagCells(have);
    }
    context.replaceGrid(level, shared);
  }

  void initBaseGrid(T boxSize, size_t n) {
    if (context.getNumLevels() > 0) throw std::runtime_error("Cannot re-initialize the base grid");
    {
      char buf[256];
      snprintf(buf, sizeof buf, "Initialized the base grid:\n  Box length            = %g Mpc/h\n  n                     = %zu\n"
               "  dx                    = %g", (double) boxSize, n, (double) boxSize / (double) n);
      logLine(buf);
    }
    context.addLevel(boxSize, n);
    gadgetTypesForLevels.push_back(mruGadgetType);   // (the type of the base level is only reserved once a zoom is opened, ic.hpp:611-613)
    levelAdded();
  }

  void initZoomGrid(size_t zoomfac, size_t n, const zoom::Int3 *userCorner = nullptr) {  // ic.hpp:398-598
    if (context.getNumLevels() < 1) throw std::runtime_error("Cannot initialise a zoom grid before initialising the base grid");
    auto &up = context.getGridForLevel(context.getNumLevels() - 1);
    std::vector<size_t> flags;
    up.getFlaggedCells(flags);
    if (flags.empty()) throw std::runtime_error("Cannot initialise zoom without marking particles to be replaced");
    const zoom::Int3 corner = userCorner ? *userCorner : zoom::zoomCorner(up, zoomfac);
    if (!userCorner && zoom::flagsNearZoomBorder(up, corner, zoomfac))
      logLine("WARNING: Opening a zoom where flagged particles are within 3 pixels of the edge. This is prone to numerical errors.");
    if (std::min(corner[0], std::min(corner[1], corner[2])) < 0)
      throw std::runtime_error("The newly initialised zoom level is not fully contained within the previous level");
    zoomParticleArray.push_back(zoom::trimToZoomWindow(up, flags, corner, (long) (up.size / zoomfac)));
    context.addLevel(up.thisGridSize / T(zoomfac), n, up.offsetLower[0] + T(corner[0]) * up.cellSize,
                     up.offsetLower[1] + T(corner[1]) * up.cellSize, up.offsetLower[2] + T(corner[2]) * up.cellSize);
    gadgetTypesForLevels.push_back(mruGadgetType);
    if (!pUnderlying) usedGadgetParticleTypes[mruGadgetType] = true;
    // flags of the level above carried onto the new grid (makeProxyGridToMatch, grid.hpp:195-228)
    auto &g = context.getGridForLevel(context.getNumLevels() - 1);
    {
      char buf[512];
      snprintf(buf, sizeof buf, "Initialized a zoom region:\n  Subbox length         = %g Mpc/h\n  n                     = %zu\n"
               "  dx                    = %g\n  Zoom factor           = %zu\n  Num particles         = %zu\n"
               "  Low-left corner in parent grid = (%ld, %ld, %ld)\n  Low-left corner (h**-1 Mpc)    = %g, %g, %g",
               (double) g.thisGridSize, g.size, (double) g.cellSize, zoomfac, flags.size(), corner[0], corner[1], corner[2],
               (double) g.offsetLower[0], (double) g.offsetLower[1], (double) g.offsetLower[2]);
      zoomLogPending = buf;  // completed with the particle count once the flags of the new level are in place
    }
    const long ratio = std::lround((double) up.cellSize / (double) g.cellSize), nu = (long) up.size, nf = (long) g.size;
    const long W = (long) up.simEquivalentSize;
    std::vector<char> mask(up.size3, 0);
    for (size_t id : flags) mask[id] = 1;
    std::vector<size_t> fine;
    for (long x = 0; x < nf; ++x) {
      const long ux = (x / ratio + corner[0]) % W;
      if (ux >= nu) continue;
      for (long y = 0; y < nf; ++y) {
        const long uy = (y / ratio + corner[1]) % W;
        if (uy >= nu) continue;
        for (long z = 0; z < nf; ++z) {
          const long uz = (z / ratio + corner[2]) % W;
          if (uz < nu && mask[(size_t) ((ux * nu + uy) * nu + uz)]) fine.push_back((size_t) ((x * nf + y) * nf + z));
        }
      }
    }
    g.flagCells(fine);
    levelAdded();
    if (outputFormat != "grafic") zoomLogPending += "\n  Total particles       = " + std::to_string(mapper()->size());
    logLine(zoomLogPending);
  }
  std::string zoomLogPending;

  static std::vector<size_t> readIds(const std::string &fn) {
    std::ifstream f(fn);
    if (!f.is_open()) throw std::runtime_error("File " + fn + " not found");
    std::vector<size_t> ids;
    size_t v;
    while (f >> v) ids.push_back(v);
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    return ids;
  }

  void runParamFile(const std::string &filename) {
    std::ifstream f(filename);
    if (!f.is_open()) throw std::runtime_error("Paramfile " + filename + " not found");
    dir = filename.find('/') == std::string::npos ? "." : filename.substr(0, filename.rfind('/'));
    std::string line;
    int lineno = 0;
    while (std::getline(f, line)) {
      ++lineno;
      std::istringstream ss(line);
      std::string cmd;
      if (!(ss >> cmd) || cmd[0] == '#' || cmd[0] == '%') continue;
      for (auto &c : cmd) c = (char) std::tolower(c);
      try {
        runCommand(cmd, ss);
      } catch (std::exception &e) {
        throw std::runtime_error("Error \"" + std::string(e.what()) + "\" on line " + std::to_string(lineno) + " (\"" + line + "\")");
      }
    }
  }

  // one command of a paramfile (lower-cased name, the rest of the line)
  void runCommand(const std::string &cmd, std::istringstream &ss) {
    auto path = [this](const std::string &p) { return p[0] == '/' ? p : dir + "/" + p; };
    auto num = [&]() { double v; if (!(ss >> v)) throw std::runtime_error("Insufficient arguments for command " + cmd); return v; };
    auto str = [&]() { std::string v; if (!(ss >> v)) throw std::runtime_error("Insufficient arguments for command " + cmd); return v; };
    {
      {
        if (cmd == "om") cosmo.OmegaM0 = (T) num();
        else if (cmd == "ol") cosmo.OmegaLambda0 = (T) num();
        else if (cmd == "ob") cosmo.OmegaBaryons0 = (T) num();
        else if (cmd == "s8") cosmo.sigma8 = (T) num();
        else if (cmd == "ns") cosmo.ns = (T) num();
        else if (cmd == "hubble") cosmo.hubble = (T) num();
        else if (cmd == "tcmb") cosmo.TCMB = (T) num();
        else if (cmd == "zin" || cmd == "z") setRedshift((T) num());
        else if (cmd == "camb") cambFile = path(str());
        else if (cmd == "basegrid" || cmd == "base_grid") { const T box = (T) num(); initBaseGrid(box, (size_t) num()); }
        else if (cmd == "zoomgrid" || cmd == "zoom_grid") { const size_t zf = (size_t) num(); initZoomGrid(zf, (size_t) num()); }
        else if (cmd == "zoomgrid_with_origin_at" || cmd == "zoom_grid_with_origin_at") {
          zoom::Int3 c{(long) num(), 0, 0}; c[1] = (long) num(); c[2] = (long) num();
          const size_t zf = (size_t) num();
          initZoomGrid(zf, (size_t) num(), &c);
        }
        else if (cmd == "centre" || cmd == "center") { centre[0] = (T) num(); centre[1] = (T) num(); centre[2] = (T) num(); }
        else if (cmd == "select_sphere") zoom::selectSphere<T>(context, centre, (float) num());
        else if (cmd == "select_cube") zoom::selectCube<T>(context, centre, (float) num());
        else if (cmd == "select_ellipse") { const float a = (float) num(), b = (float) num(); zoom::selectEllipse<T>(context, centre, a, b, (float) num()); }
        else if (cmd == "select_nearest") zoom::selectNearest<T>(context, centre);
        else if (cmd == "idfile" || cmd == "id_file") { auto m = mapper(); m->unflagAllParticles(); m->flagParticles(readIds(path(str()))); }
        else if (cmd == "append_idfile" || cmd == "merge_id_file") mapper()->flagParticles(readIds(path(str())));
        else if (cmd == "dump_idfile" || cmd == "dump_id_file") {
          std::vector<size_t> ids;
          mapper()->getFlaggedParticles(ids);
          std::ofstream o(path(str()));
          for (size_t i : ids) o << i << std::endl;
        }
        else if (cmd == "gadget_particle_type") {
          const unsigned t = (unsigned) num();
          if (t >= 6) throw std::runtime_error("Gadget particle type must be between 0 and 5");
          if (t != mruGadgetType) {  // ic.hpp:690-703
            if (usedGadgetParticleTypes[t])
              throw std::runtime_error("Cannot reuse a gadget particle type; this type was already used in a different section of the ICs");
            mruGadgetType = t;
          }
          if (gadgetTypesForLevels.empty()) throw std::runtime_error("Can't set a gadget particle type until a level has been initialised to set it for");
          gadgetTypesForLevels.back() = t;
        }
        else if (cmd == "gadget_flagged_particle_type") {  // ic.hpp:703-714: the flagged cells of the finest level as a separate type
          const unsigned t = (unsigned) num();
          if (t >= 6) throw std::runtime_error("Gadget particle type must be between 0 and 5");
          if (usedGadgetParticleTypes[t])
            throw std::runtime_error("Cannot reuse a gadget particle type; this type was already used in a different section of the ICs");
          flaggedGadgetParticleType = (int) t;
          usedGadgetParticleTypes[t] = true;
          flaggedCellsForType.clear();
          context.getGridForLevel(context.getNumLevels() - 1).getFlaggedCells(flaggedCellsForType);
        }
        else if (cmd == "gadget_num_files") nGadgetFiles = (int) num();
        else if (cmd == "baryons_all_levels") baryonsOnAllLevels = true;
        else if (cmd == "random_seed" || cmd == "seedfourier_parallel") { seed = (unsigned long) num(); noiseMode = GX_NOISE_FOURIER_PARALLEL; }
        else if (cmd == "random_seed_serial" || cmd == "seedfourier") { seed = (unsigned long) num(); noiseMode = GX_NOISE_FOURIER_SERIAL; }
        else if (cmd == "seedfourier_reverse") { seed = (unsigned long) num(); noiseMode = GX_NOISE_FOURIER_SERIAL; reverse = true; }
        else if (cmd == "random_seed_real_space" || cmd == "seed") { seed = (unsigned long) num(); noiseMode = GX_NOISE_REAL; }
        else if (cmd == "fix_power" || cmd == "exact_power_spectrum_enforcement") fixPower = true;
        else if (cmd == "outname") outputFilename = str();
        else if (cmd == "outdir") outputFolder = path(str());
        else if (cmd == "outformat") {
          outputFormat = str();
          if (outputFormat == "2") outputFormat = "gadget2"; else if (outputFormat == "3") outputFormat = "gadget3"; else if (outputFormat == "4") outputFormat = "tipsy";
          if (outputFormat != "tipsy" && outputFormat != "gadget2" && outputFormat != "gadget3") throw std::runtime_error("Unknown output format");
        }
        else if (cmd == "done") done();
        else throw std::runtime_error("Unknown command " + cmd);
      }
    }
  }

  // ICGenerator::done -> write (ic.hpp:1633-1648, 1275-1331)
  void done() {
    if (outputFormat.empty()) throw std::runtime_error("No output format specified!");
    if (context.getNumLevels() == 0) throw std::runtime_error("No grid has been initialised");
    pMapper = mapper();
    if (layoutOnly) return;
    const double boxlen0 = (double) context.getGridForLevel(0).periodicDomainSize;
    if (!particlesFile.empty()) {  // writers only: the particles come from a file instead of the device
      std::ifstream pf(particlesFile, std::ios::binary);
      if (!pf.is_open()) throw std::runtime_error("Could not open " + particlesFile);
      const size_t N = pMapper->size();
      std::vector<double> buf(6 * N);
      pf.read(reinterpret_cast<char *>(buf.data()), (std::streamsize) (buf.size() * sizeof(double)));
      if ((size_t) pf.gcount() != buf.size() * sizeof(double)) throw std::runtime_error("particle file has the wrong size");
      const std::string out = outputFolder + "/" + outputFilename;
      if (outputFormat == "tipsy") {
        particle::MappedParticles<T, T> P(*pMapper, nullptr, cosmo);
        P.pos.assign(buf.begin(), buf.begin() + 3 * N);
        P.vel.assign(buf.begin() + 3 * N, buf.end());
        io::tipsy::saveMapped<T>(out + ".tipsy", boxlen0, P, cosmo);
      } else {
        particle::MappedParticles<T, float> P(*pMapper, nullptr, cosmo);
        P.pos.resize(3 * N); P.vel.resize(3 * N);
        for (size_t i = 0; i < 3 * N; ++i) { P.pos[i] = (float) T(buf[i]); P.vel[i] = (float) T(buf[3 * N + i]); }
        io::gadget::save<T>(out + ".gadget", boxlen0, P, cosmo, outputFormat == "gadget2" ? 2 : 3, nGadgetFiles);
      }
      return;
    }
    if (noiseMode < 0) throw std::runtime_error("The random number generator has not been seeded");
    const size_t nl = context.getNumLevels();
    std::vector<std::shared_ptr<fields::Field<T>>> delta;
    for (size_t l = 0; l < nl; ++l) delta.push_back(std::make_shared<fields::Field<T>>(context.getGridForLevel(l), true));
    fields::RandomFieldGenerator<T> rfg(delta);
    rfg.seed(seed);
    rfg.setDrawInFourierSpace(noiseMode != GX_NOISE_REAL);
    rfg.setParallel(noiseMode == GX_NOISE_FOURIER_PARALLEL);
    rfg.setReverseRandomDrawOrder(reverse);
    rfg.draw();
    cosmology::CAMB<T> spectrum(cosmo, cambFile, cosmology::species::all);
    for (auto &d : delta) { if (fixPower) d->enforceUnitVariance(); d->applyTransferFunction(spectrum, 0.5); }
    particle::MultiLevelParticleGenerator<T> gen(context, delta);
    const double boxlen = (double) context.getGridForLevel(0).periodicDomainSize;
    const std::string base = outputFolder + "/" + outputFilename;
    if (outputFormat == "tipsy") {
      particle::MappedParticles<T, T> P(*pMapper, &gen, cosmo);
      io::tipsy::saveMapped<T>(base + ".tipsy", boxlen, P, cosmo);
    } else {
      particle::MappedParticles<T, float> P(*pMapper, &gen, cosmo);
      io::gadget::save<T>(base + ".gadget", boxlen, P, cosmo, outputFormat == "gadget2" ? 2 : 3, nGadgetFiles);
    }
  }
};
}  // namespace ic

}  // namespace gx
#endif
