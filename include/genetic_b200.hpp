// genetic_b200.hpp -- C++ host layer over the C ABI (include/genetic_b200.h): the reference's class surface for the hot
// path, header-only, C++17, no CUDA headers needed by the including translation unit.
//
// The reference (pynbody/genetIC) is a C++ program whose seam is a class API. The classes below
// keep its names, argument meaning and error behaviour (std::runtime_error, tools/parser.hpp:202-207) so that code
// written against the reference reads the same; the data lives in HBM and every operation forwards to a kernel. Paths
// in the comments are relative to genetIC/src/.
//
//   gx::grids::Grid<T>                        simulation/grid/grid.hpp:84-130          (const metadata + flags)
//   gx::fields::Field<T>                      simulation/field/field.hpp:181-892       (device storage, host mirror on demand)
//   gx::filters::*                            simulation/filters/filter.hpp:60-317
//   gx::cosmology::CosmologicalParameters<T>  cosmology/parameters.hpp:20-106
//   gx::cosmology::CAMB<T>                    cosmology/camb.hpp:173-332
//   gx::fields::RandomFieldGenerator<T>       simulation/field/randomfieldgenerator.hpp:25-263
//   gx::particle::ZeldovichParticleGenerator<T>  simulation/particles/zeldovich.hpp:132-304
//   gx::multilevelgrid::MultiLevelGrid<T>, gx::filters::FilterFamily<T>, gx::fields::MultiLevelField<T>   (zoom levels, metric, chi^2)
//   gx::ic::ICGenerator<T>                    ic.hpp (the unigrid command surface: cosmology, camb, seeds, basegrid, done)
//   gx::io::tipsy::save                       io/tipsy.hpp:207-271 (dark-matter-only unigrid)
#pragma once
#include "genetic_b200.h"

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace gx {

inline void check(int rc) {
  if (rc != 0) throw std::runtime_error(gx_last_error());
}
template <typename T> constexpr int dtype_of() { return std::is_same<T, double>::value ? GX_F64 : GX_F32; }

// ---------------------------------------------------------------------------------------------------------------------
namespace grids {
template <typename T> class Grid : public std::enable_shared_from_this<Grid<T>> {
public:
  const T periodicDomainSize, thisGridSize, cellSize;
  const T offsetLower[3];
  const size_t size, size2, size3;
  const size_t simEquivalentSize;
  T kMin;

  // Grid(simsize, n, dx, x0, y0, z0), grid.hpp:95-115
  Grid(T simsize, size_t n, T dx = 0, T x0 = 0, T y0 = 0, T z0 = 0)
      : periodicDomainSize(simsize), thisGridSize((dx == 0 ? simsize / n : dx) * n), cellSize(dx == 0 ? simsize / n : dx),
        offsetLower{x0, y0, z0}, size(n), size2(n * n), size3(n * n * n),
        simEquivalentSize((size_t) std::llround((double) simsize / (double) (dx == 0 ? simsize / n : dx))) {
    kMin = T(2. * M_PI / thisGridSize);  // setKmin, grid.hpp:128-130
  }
  T getFourierKmin() const { return kMin; }
  bool coversFullSimulation() const { return periodicDomainSize == thisGridSize; }
  void flagCells(const std::vector<size_t> &cells) { flags = cells; }
  void getFlaggedCells(std::vector<size_t> &out) const { out.insert(out.end(), flags.begin(), flags.end()); }
  size_t numFlaggedCells() const { return flags.size(); }
  gx_grid_desc desc() const {
    gx_grid_desc d;
    d.n = (int) size; d.cell_size = (double) cellSize; d.domain_size = (double) periodicDomainSize;
    for (int a = 0; a < 3; ++a) d.offset_lower[a] = (double) offsetLower[a];
    return d;
  }
private:
  std::vector<size_t> flags;
};
}  // namespace grids

// ---------------------------------------------------------------------------------------------------------------------
namespace filters {
// A filter is a product of factors the device evaluates per mode (gx_filter_factor). Filter itself is the identity.
template <typename T> class Filter {
public:
  std::vector<gx_filter_factor> factors;
  virtual ~Filter() = default;
  Filter() = default;
  explicit Filter(std::vector<gx_filter_factor> f) : factors(std::move(f)) {}
  Filter<T> operator*(const Filter<T> &o) const {
    std::vector<gx_filter_factor> f(factors);
    f.insert(f.end(), o.factors.begin(), o.factors.end());
    return Filter<T>(f);
  }
  Filter<T> operator*(T c) const {
    std::vector<gx_filter_factor> f(factors);
    f.push_back(gx_filter_factor{GX_FILTER_CONST, (double) c, 0.0, 1});
    return Filter<T>(f);
  }
};
template <typename T> struct LowPassFermiFilter : Filter<T> {  // filter.hpp:209-239, temperature = kcut / 10
  explicit LowPassFermiFilter(T kcut) { this->factors.push_back(gx_filter_factor{GX_FILTER_FERMI_LP, (double) kcut, (double) (kcut * T(0.1)), 1}); }
};
template <typename T> struct ComplementaryCovarianceFilterAdaptor : Filter<T> {  // sqrt(1 - f^2), filter.hpp:246-275
  explicit ComplementaryCovarianceFilterAdaptor(T kcut) { this->factors.push_back(gx_filter_factor{GX_FILTER_FERMI_HP_COV, (double) kcut, (double) (kcut * T(0.1)), 1}); }
};
template <typename T> struct ComplementaryFilterAdaptor : Filter<T> {  // 1 - f, filter.hpp:282-317
  explicit ComplementaryFilterAdaptor(T kcut) { this->factors.push_back(gx_filter_factor{GX_FILTER_FERMI_HP, (double) kcut, (double) (kcut * T(0.1)), 1}); }
};
}  // namespace filters

// ---------------------------------------------------------------------------------------------------------------------
namespace cosmology {
template <typename T> struct CosmologicalParameters {  // parameters.hpp:20-33; defaults of ICGenerator, ic.hpp:201-207
  T OmegaM0 = 0.3, OmegaLambda0 = 0, OmegaBaryons0 = 0, hubble = 0.701, redshift = 0, scalefactor = 1, sigma8 = 0, ns = 0.96,
    TCMB = 2.725;
};
// parameters.hpp:83-106: f = 1 (Zel'dovich, Omega_m(z) ~ 1 at the start), velocity = psi * 100 * H(a)/H0 * sqrt(a)
template <typename T> T zeldovichVelocityToOffsetRatio(const CosmologicalParameters<T> &c) {
  const T a = c.scalefactor;
  return T(T(1.0) * T(100.) * std::sqrt(c.OmegaM0 / (a * a * a) + c.OmegaLambda0) * std::sqrt(a));
}

// io::getBuffer + CAMB::readLinesFromCambOutput (io/input.hpp:47-95, camb.hpp:224-268): whitespace-separated numbers,
// a first line of column headers is skipped; 7 (old CAMB) or 13 columns.
inline std::vector<double> readNumericTable(const std::string &filename, size_t &ncol) {
  std::ifstream f(filename);
  if (!f.is_open()) throw std::runtime_error("File " + filename + " not found");
  std::vector<double> vals;
  std::string line;
  size_t lineno = 0;
  ncol = 0;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    std::string tok;
    std::vector<double> row;
    bool bad = false;
    while (!bad && (ss >> tok)) {
      try {
        size_t used = 0;
        const double v = std::stod(tok, &used);
        if (used != tok.size()) bad = true; else row.push_back(v);
      } catch (...) { bad = true; }
    }
    if (bad && lineno > 0) throw std::runtime_error("Error reading file " + filename);
    if (!bad) {
      if (ncol == 0 && !row.empty()) ncol = row.size();
      vals.insert(vals.end(), row.begin(), row.end());
    }
    ++lineno;  // a first line that does not parse is the column-header line
  }
  return vals;
}

enum class species { dm = 1, baryon = 2, all = 6 };  // CAMB columns (camb.hpp:187-190)

template <typename T> class CAMB {
public:
  gx_spectrum *handle = nullptr;
  CAMB(const CosmologicalParameters<T> &cosmo, const std::string &filename, species sp = species::all) {
    size_t ncol = 0;
    std::vector<double> tab = readNumericTable(filename, ncol);
    if (ncol != 7 && ncol != 13) throw std::runtime_error("CAMB transfer file doesn't have a sensible number of columns");
    const size_t nrows = tab.size() / ncol;
    const T norm = (T) tab[1];  // first CDM entry, camb.hpp:254
    std::vector<double> k, Ts, Tn;
    for (size_t r = 0; r < nrows; ++r) {
      if (!(tab[r * ncol] > 0)) continue;
      k.push_back((double) (T) tab[r * ncol]);
      Ts.push_back((double) (T((T) tab[r * ncol + (size_t) sp] / norm)));
      Tn.push_back((double) (T((T) tab[r * ncol + 6] / norm)));
    }
    check(gx_spectrum_create_camb(k.data(), Ts.data(), Tn.data(), (int) k.size(), cosmo.OmegaM0, cosmo.OmegaLambda0,
                                  cosmo.sigma8, cosmo.ns, cosmo.scalefactor, dtype_of<T>(), &handle));
  }
  CAMB(const CAMB &) = delete;
  ~CAMB() { if (handle) gx_spectrum_destroy(handle); }
  T operator()(T k) const { return (T) gx_spectrum_eval(handle, (double) k); }
  double amplitude() const { return gx_spectrum_amplitude(handle); }
};
}  // namespace cosmology

// ---------------------------------------------------------------------------------------------------------------------
namespace fields {
template <typename T> class Field : public std::enable_shared_from_this<Field<T>> {
public:
  using TGrid = grids::Grid<T>;

  Field(const TGrid &grid, bool fourier = true) : pGrid(&grid), fourier_(fourier) { allocate(); check(gx_memset_zero(dev, bytes(), nullptr)); }
  Field(const TGrid &grid, const std::vector<T> &data, bool fourier = true) : pGrid(&grid), fourier_(fourier) {
    allocate();
    host = data;
    upload();
  }
  Field(const Field<T> &o) : pGrid(o.pGrid), fourier_(o.fourier_) {  // deep copy, as the reference's copy constructor
    allocate();
    const_cast<Field<T> &>(o).toDevice();
    check(gx_memcpy_d2d(dev, o.dev, bytes(), nullptr));
  }
  Field<T> &operator=(const Field<T> &o) {  // deep copy into the existing buffer (reallocated if the grids differ in size)
    if (this == &o) return *this;
    const size_t have = bytes();
    pGrid = o.pGrid;
    fourier_ = o.fourier_;
    if (have != bytes()) { if (dev) gx_free(dev); dev = nullptr; allocate(); }
    const_cast<Field<T> &>(o).toDevice();
    check(gx_memcpy_d2d(dev, o.dev, bytes(), nullptr));
    hostValid = false;
    devValid = true;
    return *this;
  }
  ~Field() { if (dev) gx_free(dev); }

  const TGrid &getGrid() const { return *pGrid; }
  void *devicePointer() { toDevice(); hostValid = false; return dev; }
  const void *devicePointer() const { const_cast<Field<T> *>(this)->toDevice(); return dev; }

  // Host access in the reference's layout (real space: n^3 values; Fourier space: 2 n^2 (n/2+1) values). The reference's
  // callers index the vector directly (ic.hpp:795-796, linearmodification.hpp:80-86): the mirror is downloaded on demand
  // and the device copy is refreshed before the next kernel.
  std::vector<T> &getDataVector() { toHost(); devValid = false; return host; }
  const std::vector<T> &getDataVector() const { const_cast<Field<T> *>(this)->toHost(); return host; }

  bool isFourier() const { return fourier_; }
  void setFourier(bool f) { fourier_ = f; }
  void toFourier() { if (fourier_) return; check(gx_fft_forward(devicePointer(), n(), dt(), nullptr)); fourier_ = true; }
  void toReal() { if (!fourier_) return; check(gx_fft_inverse(devicePointer(), n(), dt(), nullptr)); fourier_ = false; }
  template <typename O> void matchFourier(const O &other) { if (other.isFourier()) toFourier(); else toReal(); }
  void ensureFourierModesAreMirrored() { check(gx_mirror(devicePointer(), n(), dt(), nullptr)); }

  std::complex<T> getFourierCoefficient(int kx, int ky, int kz) const {
    double re, im;
    check(gx_synchronize(nullptr));
    check(gx_get_mode(devicePointer(), n(), dt(), kx, ky, kz, &re, &im));
    return std::complex<T>((T) re, (T) im);
  }
  void setFourierCoefficient(int kx, int ky, int kz, std::complex<T> v) {
    check(gx_synchronize(nullptr));
    check(gx_set_mode(devicePointer(), n(), dt(), kx, ky, kz, v.real(), v.imag()));
  }

  void applyFilter(const filters::Filter<T> &f) {  // field.hpp:741-746
    toFourier();
    if (f.factors.empty()) return;
    check(gx_apply_filter(devicePointer(), n(), dt(), (double) pGrid->thisGridSize, f.factors.data(), (int) f.factors.size(), nullptr));
  }
  void applyTransferFunction(const cosmology::CAMB<T> &spectrum, double power) {  // field.hpp:394-412, covariance in-kernel
    check(gx_apply_transfer(devicePointer(), n(), dt(), (double) pGrid->thisGridSize, spectrum.handle, power, nullptr));
  }
  void applyTransferFunction(const Field<T> &covariance, double power) {
    check(gx_apply_transfer_field(devicePointer(), covariance.devicePointer(), n(), dt(), power, nullptr));
  }
  void setZeroInsideWindow(const TGrid &w) { toReal(); auto a = pGrid->desc(), b = w.desc(); check(gx_zero_window(devicePointer(), &a, dt(), &b, 1, nullptr)); }
  void setZeroOutsideWindow(const TGrid &w) { toReal(); auto a = pGrid->desc(), b = w.desc(); check(gx_zero_window(devicePointer(), &a, dt(), &b, 0, nullptr)); }
  void applyFilterInWindow(const filters::Filter<T> &f, const TGrid &window, bool windowFirst) {  // field.hpp:752-772
    Field<T> inWindow(*this);
    if (windowFirst) { inWindow.setZeroOutsideWindow(window); inWindow.toFourier(); inWindow.applyFilter(f); }
    else { inWindow.toFourier(); inWindow.applyFilter(f); inWindow.setZeroOutsideWindow(window); }
    setZeroInsideWindow(window);
    if (inWindow.isFourier()) toFourier(); else toReal();
    *this += inWindow;
  }
  void addFieldFromDifferentGrid(Field<T> &source) {  // field.hpp:804-812,834-840
    source.toReal();
    toReal();
    auto a = pGrid->desc(), b = source.getGrid().desc();
    check(gx_interp_add(devicePointer(), &a, source.devicePointer(), &b, dt(), nullptr));
  }
  void addFieldFromDifferentGridWithFilter(const Field<T> &source, const filters::Filter<T> &f) {  // FILTER_ON_COARSE_GRID, :819-832
    Field<T> tmp(source);
    tmp.applyFilter(f);
    tmp.toReal();
    toReal();
    addFieldFromDifferentGrid(tmp);
  }

  Field<T> &operator+=(const Field<T> &o) { check(gx_axpy(devicePointer(), o.devicePointer(), n(), dt(), 1.0, nullptr)); return *this; }
  Field<T> &operator-=(const Field<T> &o) { check(gx_axpy(devicePointer(), o.devicePointer(), n(), dt(), -1.0, nullptr)); return *this; }
  Field<T> &operator*=(T v) { check(gx_scale(devicePointer(), n(), dt(), (double) v, nullptr)); return *this; }
  Field<T> &operator/=(T v) { check(gx_scale(devicePointer(), n(), dt(), 1.0 / (double) v, nullptr)); return *this; }
  void addScaled(const Field<T> &o, T scale) { check(gx_axpy(devicePointer(), o.devicePointer(), n(), dt(), (double) scale, nullptr)); }
  void addScaledFourier(const Field<T> &o, T scale) { toFourier(); check(gx_axpy_fourier(devicePointer(), o.devicePointer(), n(), dt(), (double) scale, nullptr)); }
  // real space: Field::innerProduct (field.hpp:365-377); Fourier space: the weighted mode sum of MultiLevelField::innerProduct
  T innerProduct(const Field<T> &o) const {
    double out = 0;
    if (fourier_) check(gx_dot_fourier(devicePointer(), o.devicePointer(), n(), dt(), &out, nullptr));
    else check(gx_dot_real(devicePointer(), o.devicePointer(), n(), dt(), &out, nullptr));
    return (T) out;
  }
  T norm() const { return std::sqrt(innerProduct(*this)); }
  void enforceUnitVariance() { toFourier(); check(gx_unit_variance(devicePointer(), n(), dt(), nullptr)); }
  std::shared_ptr<Field<T>> copy() const { return std::make_shared<Field<T>>(*this); }

private:
  const TGrid *pGrid;
  void *dev = nullptr;
  std::vector<T> host;
  bool fourier_;
  bool hostValid = false, devValid = true;
  int n() const { return (int) pGrid->size; }
  static int dt() { return dtype_of<T>(); }
  size_t bytes() const { return gx_field_bytes(n(), dt()); }
  void allocate() { check(gx_malloc(&dev, bytes())); }
  size_t hostCount() const { return fourier_ ? bytes() / sizeof(T) : pGrid->size3; }
  void upload() {
    if (host.size() != hostCount()) throw std::runtime_error("Field: host data has the wrong length for the field's space");
    check(gx_field_upload(dev, host.data(), n(), dt(), fourier_ ? GX_LAYOUT_FOURIER : GX_LAYOUT_REAL, nullptr));
    check(gx_synchronize(nullptr));
    devValid = true;
    hostValid = true;
  }
  void toDevice() { if (!devValid) upload(); }
  void toHost() {
    if (hostValid && host.size() == hostCount()) return;
    toDevice();
    host.resize(hostCount());
    check(gx_field_download(host.data(), dev, n(), dt(), fourier_ ? GX_LAYOUT_FOURIER : GX_LAYOUT_REAL, nullptr));
    check(gx_synchronize(nullptr));
    hostValid = true;
  }
};

// fields::RandomFieldGenerator over the per-level fields of a (multi-level) output field
template <typename T> class RandomFieldGenerator {
public:
  explicit RandomFieldGenerator(std::vector<std::shared_ptr<Field<T>>> levels, unsigned long seed = 0)
      : levels(std::move(levels)), baseSeed(seed) {}
  void setDrawInFourierSpace(bool v) { drawInFourierSpace = v; }
  void setParallel(bool v) { parallel = v; }
  void setReverseRandomDrawOrder(bool v) { reverseRandomDrawOrder = v; }
  void seed(unsigned long s) {
    if (seeded) throw std::runtime_error("The random number generator has already been seeded");
    baseSeed = s;
    seeded = true;
  }
  void draw() {  // randomfieldgenerator.hpp:122-139: one gsl stream across the levels in the serial orders
    if (!seeded) throw std::runtime_error("The random number generator has not been seeded");
    unsigned long long offset = 0;
    const int mode = drawInFourierSpace ? (parallel ? GX_NOISE_FOURIER_PARALLEL : GX_NOISE_FOURIER_SERIAL) : GX_NOISE_REAL;
    for (auto &f : levels) {
      check(gx_noise_draw(f->devicePointer(), (int) f->getGrid().size, dtype_of<T>(), baseSeed, mode, reverseRandomDrawOrder ? 1 : 0,
                          &offset, nullptr));
      f->setFourier(true);
    }
  }
private:
  std::vector<std::shared_ptr<Field<T>>> levels;
  unsigned long baseSeed;
  bool drawInFourierSpace = false, parallel = false, reverseRandomDrawOrder = false, seeded = false;
};
}  // namespace fields

// ---------------------------------------------------------------------------------------------------------------------
// Zoom levels: the multi-level context, the filter family that splits the information between levels and the
// multi-level field with its covector metric (simulation/multilevelgrid/multilevelgrid.hpp:30-330,
// simulation/filters/filterfamily.hpp:86-143, simulation/field/multilevelfield.hpp:18-606). Pure sequencing over the
// single-level Field operations above; the modifications and the splice (genetic_b200/multilevel.py) build on the same calls.
namespace multilevelgrid {
template <typename T> class MultiLevelGrid {
public:
  void addLevel(std::shared_ptr<grids::Grid<T>> g) {  // multilevelgrid.hpp:112-137
    weights.push_back(pGrids.empty() ? T(1) : T(std::pow((double) (g->cellSize / pGrids[0]->cellSize), 3.0)));
    pGrids.push_back(g);
  }
  void addLevel(T size, size_t nside, T x0 = 0, T y0 = 0, T z0 = 0) {  // :97-110
    if (pGrids.empty()) simSize = size;
    addLevel(std::make_shared<grids::Grid<T>>(simSize, nside, size / nside, x0, y0, z0));
  }
  size_t getNumLevels() const { return pGrids.size(); }
  grids::Grid<T> &getGridForLevel(size_t level) { return *pGrids[level]; }
  const grids::Grid<T> &getGridForLevel(size_t level) const { return *pGrids[level]; }
  T getWeightForLevel(size_t level) const { return weights[level]; }
  size_t getNumDof() const {  // :160-172
    size_t dof = 0;
    for (size_t i = 0; i < pGrids.size(); ++i) {
      dof += pGrids[i]->size3;
      if (i > 0) {
        size_t f = (size_t) std::llround((double) (pGrids[i - 1]->cellSize / pGrids[i]->cellSize));
        dof -= pGrids[i]->size3 / (f * f * f);
      }
    }
    return dof;
  }
  std::shared_ptr<grids::Grid<T>> getGridPtr(size_t level) { return pGrids[level]; }
  void replaceGrid(size_t level, std::shared_ptr<grids::Grid<T>> g) { pGrids[level] = std::move(g); }  // DummyICGenerator: shared grids
  // set once the particle generator has folded the coarse levels into the fine ones (multilevelgrid.hpp:82-95)
  bool levelsAreCombined = false;
  void setLevelsAreCombined() { levelsAreCombined = true; }
private:
  std::vector<std::shared_ptr<grids::Grid<T>>> pGrids;
  std::vector<T> weights;
  T simSize = 0;
};
}  // namespace multilevelgrid

namespace filters {
template <typename T> class FilterFamily {  // filterfamily.hpp:86-143; FRACTIONAL_K_SPLIT = 0.3
public:
  explicit FilterFamily(const multilevelgrid::MultiLevelGrid<T> &ctx) {
    if (ctx.levelsAreCombined) {  // the fine fields carry the low frequencies too: null filter everywhere but on the finest level
      const Filter<T> null = Filter<T>() * T(0);
      for (size_t level = 0; level + 1 < ctx.getNumLevels(); ++level) { fs.push_back(null); lp.push_back(null); hp.push_back(null); }
      fs.emplace_back(); lp.emplace_back(); hp.emplace_back();
      return;
    }
    fs.emplace_back(); lp.emplace_back(); hp.emplace_back();
    for (size_t level = 0; level + 1 < ctx.getNumLevels(); ++level) {
      const auto &g = ctx.getGridForLevel(level);
      const T k_cut = T(0.3) * (T(g.size) * g.getFourierKmin());
      Filter<T> filt = fs.back();
      fs.pop_back(); lp.pop_back();
      fs.push_back(filt * LowPassFermiFilter<T>(k_cut));
      lp.push_back(LowPassFermiFilter<T>(k_cut));
      fs.push_back(filt * ComplementaryCovarianceFilterAdaptor<T>(k_cut));
      hp.push_back(ComplementaryCovarianceFilterAdaptor<T>(k_cut));
      lp.emplace_back();
    }
  }
  const Filter<T> &getFilterForLevel(size_t l) const { return fs[l]; }
  const Filter<T> &getHighPassFilterForLevel(size_t l) const { return hp[l]; }
  const Filter<T> &getLowPassFilterForLevel(size_t l) const { return lp[l]; }
private:
  std::vector<Filter<T>> fs, lp, hp;
};
}  // namespace filters

namespace fields {
template <typename T> class MultiLevelField {
public:
  using Context = multilevelgrid::MultiLevelGrid<T>;
  explicit MultiLevelField(Context &ctx, bool covector = false) : context(&ctx), isCovector(covector) {
    for (size_t l = 0; l < ctx.getNumLevels(); ++l) fieldsOnLevels.push_back(std::make_shared<Field<T>>(ctx.getGridForLevel(l), true));
  }
  MultiLevelField(const MultiLevelField<T> &o) : context(o.context), isCovector(o.isCovector) {  // deep copy
    for (auto &f : o.fieldsOnLevels) fieldsOnLevels.push_back(f->copy());
  }
  size_t getNumLevels() const { return fieldsOnLevels.size(); }
  Field<T> &getFieldForLevel(size_t l) { return *fieldsOnLevels[l]; }
  const Field<T> &getFieldForLevel(size_t l) const { return *fieldsOnLevels[l]; }
  std::vector<std::shared_ptr<Field<T>>> &getFields() { return fieldsOnLevels; }
  filters::FilterFamily<T> getFilters() const { return filters::FilterFamily<T>(*context); }
  void toFourier() { for (auto &f : fieldsOnLevels) f->toFourier(); }
  void toReal() { for (auto &f : fieldsOnLevels) f->toReal(); }
  void applyFilters(const filters::FilterFamily<T> &fam) {
    for (size_t l = 0; l < getNumLevels(); ++l) fieldsOnLevels[l]->applyFilter(fam.getFilterForLevel(l));
  }
  void addScaled(const MultiLevelField<T> &other, T scale) {  // multilevelfield.hpp:194-212
    toFourier();
    for (size_t l = 0; l < getNumLevels(); ++l) fieldsOnLevels[l]->addScaledFourier(other.getFieldForLevel(l), scale);
  }
  void operator/=(T ratio) { for (auto &f : fieldsOnLevels) *f /= ratio; }
  void convertToCovector() { if (isCovector) throw std::runtime_error("already a covector"); applyMetric(true); isCovector = true; }
  void convertToVector() { if (!isCovector) throw std::runtime_error("already a vector"); applyMetric(false); isCovector = false; }
  T innerProduct(const MultiLevelField<T> &other) const {  // :248-294
    if (!isCovector) throw std::runtime_error("The inner product can only be taken if one of the fields is a covector");
    if (other.isCovector) {
      MultiLevelField<T> v(other);
      v.convertToVector();
      v.toFourier();
      return innerProduct(v);
    }
    T result = 0;
    for (size_t l = 0; l < getNumLevels(); ++l) result += fieldsOnLevels[l]->innerProduct(other.getFieldForLevel(l));
    return result;
  }
  T getChi2() {  // :461-487 (the field is white noise)
    toFourier();
    MultiLevelField<T> c(*this);
    c.convertToCovector();
    return c.innerProduct(*this);
  }
private:
  Context *context;
  std::vector<std::shared_ptr<Field<T>>> fieldsOnLevels;
  bool isCovector;
  void applyMetric(bool toCovector) {  // :346-391
    auto fam = getFilters();
    toFourier();
    const size_t nl = getNumLevels();
    std::vector<std::shared_ptr<Field<T>>> fresh;
    for (size_t level = 0; level < nl; ++level) {
      auto result = fieldsOnLevels[level]->copy();
      const auto &f = fam.getFilterForLevel(level);
      if (toCovector && level + 1 < nl) {
        const auto &window = context->getGridForLevel(level + 1);
        result->applyFilterInWindow(f, window, true);
        result->applyFilterInWindow(f, window, false);
      } else {
        result->applyFilter(f * f);
      }
      for (size_t src = 0; src < nl; ++src) {
        if (src == level) continue;
        const T ratio = context->getWeightForLevel(level) / context->getWeightForLevel(src);
        result->addFieldFromDifferentGridWithFilter(*fieldsOnLevels[src], fam.getFilterForLevel(level) * fam.getFilterForLevel(src) * T(std::sqrt((double) ratio)));
      }
      result->toFourier();
      fresh.push_back(result);
    }
    fieldsOnLevels = fresh;
  }
};
}  // namespace fields

// ---------------------------------------------------------------------------------------------------------------------
namespace particle {
template <typename T> class ZeldovichParticleGenerator {
public:
  using TField = fields::Field<T>;
  explicit ZeldovichParticleGenerator(TField &linearOverdensityField) : delta(linearOverdensityField) { recalculate(); }
  void recalculate() {  // calculateOffsetFields, zeldovich.hpp:152-181,248-251: operator, mirror and scaling fused into the passes
    delta.toFourier();
    pOff.clear();
    const auto &g = delta.getGrid();
    for (int axis = 0; axis < 3; ++axis) {
      auto psi = std::make_shared<TField>(g, true);
      check(gx_zeldovich_offset(delta.devicePointer(), psi->devicePointer(), (int) g.size, dtype_of<T>(), (double) g.thisGridSize, axis, nullptr));
      psi->setFourier(false);
      pOff.push_back(psi);
    }
  }
  std::vector<std::shared_ptr<TField>> getGeneratedFields() { return pOff; }
private:
  TField &delta;
  std::vector<std::shared_ptr<TField>> pOff;
};
}  // namespace particle

// ---------------------------------------------------------------------------------------------------------------------
namespace io { namespace tipsy {
// TipsyOutput::operator() for a dark-matter-only unigrid (io/tipsy.hpp:207-271): header (32 bytes with padding), then
// {mass, x, y, z, vx, vy, vz, eps, phi} as float per particle in particle-id (= cell index) order.
template <typename T>
void save(const std::string &filename, double boxLength, size_t npart, const T *pos, const T *vel, T mass, T eps,
          const cosmology::CosmologicalParameters<T> &cosmo) {
  T tot_mass = 0;
  for (size_t i = 0; i < npart; ++i) tot_mass += mass;
  const double mass_factor = cosmo.OmegaM0 / tot_mass;
  const double pos_factor = 1. / boxLength;
  const double dKpcUnit = boxLength * 1000 / cosmo.hubble;
  const double dMsolUnit = 1e10 / cosmo.hubble / mass_factor;
  const double dKmsUnit = std::sqrt(4.3022682e-6 * dMsolUnit / dKpcUnit);
  const double vel_factor = std::pow((double) cosmo.scalefactor, -0.5) / dKmsUnit;
  struct { double scalefactor; int n, ndim, ngas, ndark, nstar, pad; } header;
  static_assert(sizeof(header) == 32, "tipsy header is 32 bytes");
  header.scalefactor = cosmo.scalefactor; header.n = (int) npart; header.ndim = 3; header.ngas = 0; header.ndark = (int) npart;
  header.nstar = 0; header.pad = 0;
  std::ofstream f(filename, std::ios::binary);
  if (!f.is_open()) throw std::runtime_error("Could not open " + filename + " for writing");
  f.write(reinterpret_cast<const char *>(&header), sizeof(header));
  std::vector<float> rec(9 * npart);
  for (size_t i = 0; i < npart; ++i) {
    float *p = &rec[9 * i];
    p[0] = (float) (mass * mass_factor);
    for (int a = 0; a < 3; ++a) {
      p[1 + a] = (float) (pos[3 * i + a] * pos_factor - 0.5);
      p[4 + a] = (float) (vel[3 * i + a] * vel_factor);
    }
    p[7] = (float) (eps * pos_factor);
    p[8] = 0.0f;
  }
  f.write(reinterpret_cast<const char *>(rec.data()), rec.size() * sizeof(float));
}
}}  // namespace io::tipsy

// ---------------------------------------------------------------------------------------------------------------------
namespace ic {
// The unigrid slice of ICGenerator (ic.hpp): the commands of a paramfile that drive the hot path on one level.
// Commands (main.cpp:22-172): Om, Ol, Ob, s8, zin, ns, hubble, camb, base_grid|basegrid, random_seed|seed_field_fourier_parallel,
// random_seed_serial|seedfourier, seedfourier_reverse, random_seed_real_space|seed, fix_power, outname, outdir, outformat, done.
// Anything else throws, as the reference's dispatcher does for an unknown command (tools/parser.hpp:150-160).
template <typename T> class ICGenerator {
public:
  cosmology::CosmologicalParameters<T> cosmo;
  std::string cambFile, outputFolder = ".", outputFilename = "IC", outputFormat;
  T boxSize = 0;
  size_t n = 0;
  unsigned long seed = 0;
  int noiseMode = -1;
  bool reverse = false, fixPower = false, haveOutput = false;
  std::vector<T> pos, vel;  // [n^3][3] after done()
  T mass = 0, eps = 0;

  void setRedshift(T z) { cosmo.redshift = z; cosmo.scalefactor = T(1.) / (z + T(1.)); }  // setZ0, ic.hpp:339-342

  void runParamFile(const std::string &filename) {
    std::ifstream f(filename);
    if (!f.is_open()) throw std::runtime_error("Paramfile " + filename + " not found");
    const std::string dir = filename.find('/') == std::string::npos ? "." : filename.substr(0, filename.rfind('/'));
    std::string line;
    while (std::getline(f, line)) {
      const auto hash = line.find('#');
      if (hash != std::string::npos) line = line.substr(0, hash);
      std::istringstream ss(line);
      std::string cmd;
      if (!(ss >> cmd)) continue;
      for (auto &c : cmd) c = (char) std::tolower(c);
      auto num = [&]() { double v; if (!(ss >> v)) throw std::runtime_error("Insufficient arguments for command " + cmd); return v; };
      auto str = [&]() { std::string v; if (!(ss >> v)) throw std::runtime_error("Insufficient arguments for command " + cmd); return v; };
      if (cmd == "om") cosmo.OmegaM0 = (T) num();
      else if (cmd == "ol") cosmo.OmegaLambda0 = (T) num();
      else if (cmd == "ob") cosmo.OmegaBaryons0 = (T) num();
      else if (cmd == "s8") cosmo.sigma8 = (T) num();
      else if (cmd == "ns") cosmo.ns = (T) num();
      else if (cmd == "hubble") cosmo.hubble = (T) num();
      else if (cmd == "zin" || cmd == "z") setRedshift((T) num());
      else if (cmd == "camb") { cambFile = str(); if (cambFile[0] != '/') cambFile = dir + "/" + cambFile; }
      else if (cmd == "basegrid" || cmd == "base_grid") { boxSize = (T) num(); n = (size_t) num(); }
      else if (cmd == "random_seed" || cmd == "seed_field_fourier_parallel") { seed = (unsigned long) num(); noiseMode = GX_NOISE_FOURIER_PARALLEL; }
      else if (cmd == "random_seed_serial" || cmd == "seedfourier" || cmd == "seed_field_fourier") { seed = (unsigned long) num(); noiseMode = GX_NOISE_FOURIER_SERIAL; }
      else if (cmd == "seedfourier_reverse" || cmd == "seed_field_fourier_reverse") { seed = (unsigned long) num(); noiseMode = GX_NOISE_FOURIER_SERIAL; reverse = true; }
      else if (cmd == "random_seed_real_space" || cmd == "seed" || cmd == "seed_field") { seed = (unsigned long) num(); noiseMode = GX_NOISE_REAL; }
      else if (cmd == "fix_power") fixPower = true;
      else if (cmd == "outname") outputFilename = str();
      else if (cmd == "outdir") { outputFolder = str(); if (outputFolder[0] != '/') outputFolder = dir + "/" + outputFolder; }
      else if (cmd == "outformat") outputFormat = str();
      else if (cmd == "done") done();
      else throw std::runtime_error("Unknown command " + cmd + " (this host layer drives the single-level hot path only)");
    }
  }

  // ICGenerator::done -> write (ic.hpp:1633-1648, 1275-1331) on one level
  void done() {
    if (outputFormat.empty()) throw std::runtime_error("No output format specified!");
    if (n == 0) throw std::runtime_error("Trying to apply zeldovich approximation, but no grids have been created");
    if (noiseMode < 0) throw std::runtime_error("The random number generator has not been seeded");
    if (outputFormat != "tipsy" && outputFormat != "none") throw std::runtime_error("Unknown output format");
    grids::Grid<T> grid(boxSize, n);
    auto field = std::make_shared<fields::Field<T>>(grid, true);
    fields::RandomFieldGenerator<T> rfg({field});
    rfg.seed(seed);
    rfg.setDrawInFourierSpace(noiseMode != GX_NOISE_REAL);
    rfg.setParallel(noiseMode == GX_NOISE_FOURIER_PARALLEL);
    rfg.setReverseRandomDrawOrder(reverse);
    rfg.draw();
    if (fixPower) field->enforceUnitVariance();
    cosmology::CAMB<T> spectrum(cosmo, cambFile, cosmology::species::all);
    field->applyTransferFunction(spectrum, 0.5);
    particle::ZeldovichParticleGenerator<T> gen(*field);
    // ZeldovichParticleEvaluator::getParticle for every cell (zeldovich.hpp:87-119), particle id = cell index
    const size_t n3 = grid.size3;
    void *dpos = nullptr, *dvel = nullptr;
    check(gx_malloc(&dpos, 3 * n3 * sizeof(T)));
    check(gx_malloc(&dvel, 3 * n3 * sizeof(T)));
    const T ratio = cosmology::zeldovichVelocityToOffsetRatio(cosmo);
    auto psi = gen.getGeneratedFields();
    for (int axis = 0; axis < 3; ++axis)
      check(gx_emit_component(psi[axis]->devicePointer(), (int) n, dtype_of<T>(), axis, (double) grid.cellSize, 0.0,
                              (double) grid.periodicDomainSize, (double) ratio, (char *) dpos + axis * sizeof(T),
                              (char *) dvel + axis * sizeof(T), dtype_of<T>(), 3, 0, (long long) n3, nullptr));
    pos.resize(3 * n3);
    vel.resize(3 * n3);
    check(gx_memcpy_d2h(pos.data(), dpos, 3 * n3 * sizeof(T), nullptr));
    check(gx_memcpy_d2h(vel.data(), dvel, 3 * n3 * sizeof(T), nullptr));
    check(gx_synchronize(nullptr));
    gx_free(dpos);
    gx_free(dvel);
    const T boxMass = T(27.744948) * cosmo.OmegaM0 * T(std::pow((double) grid.periodicDomainSize, 3.0));  // zeldovich.hpp:61
    mass = boxMass * T(std::pow((double) (grid.cellSize / grid.periodicDomainSize), 3.0));               // cellMassFrac, grid.hpp:112
    eps = grid.cellSize * T(1.0) * T(0.01075);                                                          // zeldovich.hpp:112-119
    if (outputFormat == "tipsy")
      io::tipsy::save<T>(outputFolder + "/" + outputFilename + ".tipsy", (double) boxSize, n3, pos.data(), vel.data(), mass, eps, cosmo);
    haveOutput = true;
  }
};
}  // namespace ic

}  // namespace gx
