// C++ host layer, part 3: genetic modifications and splicing over the C ABI (include/genetic_b200.h).
//
// The reference's classes, restated with device-resident fields (paths relative to genetIC/src/simulation/):
//   modifications::LinearModification, OverdensityModification, VelocityModification, PotentialModification
//                                             modifications/linearmodification.hpp:24-285
//   modifications::FilteredVarianceModification        modifications/quadraticmodification.hpp:82-221
//   modifications::ModificationManager                  modifications/modificationmanager.hpp:20-385
//   MultiLevelGridBase::generateMultilevelCovectorFromHiresCovector   multilevelgrid/multilevelgrid.hpp:245-321
//   modifications::spliceOneLevel, tools::numerics::conjugateGradient  modifications/splice.hpp:38-118, tools/numerics/cg.hpp:14-58
//   ICGenerator::modify / calculate / applyModifications / dumpGrid     ic.hpp:1598-1631, 822-851
// Everything numerical is a call into the C ABI: the classes only sequence device operations (dot products, axpy, filters,
// transfers, the cross-grid kernels, masked reductions); fields never leave HBM except the scalars.
// `genetic_b200/multilevel.py` is the same flow in Python; both are pinned on the reference's goldens test_02c / 02f / 02g / 02i / 24.
#ifndef GENETIC_B200_MODIF_HPP
#define GENETIC_B200_MODIF_HPP
#include "genetic_b200_zoom.hpp"
#include "genetic_b200_grafic.hpp"
#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>

namespace gx {
namespace modifications {

// where the warnings and the "calculated value" lines go (the reference's logging::entry())
inline std::function<void(const std::string &)> &logger() {
  static std::function<void(const std::string &)> f = [](const std::string &s) { logLine(s); };
  return f;
}

enum class TransferType { whitenoise, dm, all };  // particle::species (species.hpp) as far as the dark-matter path needs them

// the multi-level context with its covariances (MultiLevelGrid::getCovariance)
template <typename T> struct Context {
  multilevelgrid::MultiLevelGrid<T> *grids = nullptr;
  const cosmology::CAMB<T> *dm = nullptr, *all = nullptr;
  const cosmology::CAMB<T> &covariance(TransferType t) const { return t == TransferType::dm ? *dm : *all; }
  size_t getNumLevels() const { return grids->getNumLevels(); }
};

// fields::MultiLevelField with its transfer type (multilevelfield.hpp:18-60, 434-454)
template <typename T> class TypedField : public fields::MultiLevelField<T> {
public:
  using Base = fields::MultiLevelField<T>;
  TypedField(Context<T> &c, TransferType t, bool covector = false) : Base(*c.grids, covector), ctx(&c), transferType(t) {}
  TypedField(const TypedField<T> &o) = default;  // Base's copy constructor deep-copies the fields
  TransferType getTransferType() const { return transferType; }
  void applyTransferRatio(TransferType from, TransferType to) {  // applyTransferRatioOneLevel, :434-454
    if (from == to) return;
    this->toFourier();
    for (size_t l = 0; l < this->getNumLevels(); ++l) {
      auto &f = this->getFieldForLevel(l);
      if (from != TransferType::whitenoise) f.applyTransferFunction(ctx->covariance(from), -0.5);
      if (to != TransferType::whitenoise) f.applyTransferFunction(ctx->covariance(to), 0.5);
    }
  }
  void applyPowerSpectrumFor(TransferType species) { applyTransferRatio(transferType, species); transferType = species; }
  Context<T> *ctx;
  TransferType transferType;
};

// device byte mask of a grid's flagged cells (n^3 bytes, 1 = flagged; complement: 0 = flagged)
template <typename T> class FlagMask {
public:
  FlagMask(const grids::Grid<T> &g, bool complement = false) {
    std::vector<size_t> flags;
    g.getFlaggedCells(flags);
    std::vector<unsigned char> m(g.size3, complement ? 1 : 0);
    for (size_t i : flags) m[i] = complement ? 0 : 1;
    count = flags.size();
    check(gx_malloc(&dev, m.size()));
    check(gx_memcpy_h2d(dev, m.data(), m.size(), nullptr));
    check(gx_synchronize(nullptr));
  }
  ~FlagMask() { if (dev) gx_free(dev); }
  FlagMask(const FlagMask &) = delete;
  void *dev = nullptr;
  size_t count = 0;
};

template <typename T> void mulMask(fields::Field<T> &f, const FlagMask<T> &m) {  // Field::operator*=(Field<char>), field.hpp:340-362
  f.toReal();
  check(gx_mul_mask(f.devicePointer(), (const unsigned char *) m.dev, (int) f.getGrid().size, dtype_of<T>(), nullptr));
}

// multilevelgrid.hpp:245-321: the localisation covector of the finest grid carried to every level
template <typename T>
TypedField<T> generateMultilevelCovectorFromHiresCovector(Context<T> &ctx, std::shared_ptr<fields::Field<T>> hires, TransferType type) {
  const size_t nl = ctx.getNumLevels();
  TypedField<T> ret(ctx, type, true);
  auto &fs = ret.getFields();
  fs[nl - 1] = hires;
  for (size_t l = 0; l + 1 < nl; ++l) fs[l] = std::make_shared<fields::Field<T>>(ctx.grids->getGridForLevel(l), false);
  if (nl > 1) {
    fs[nl - 1]->toReal();
    for (size_t l = nl - 1; l-- > 0;) {
      auto &hi = *fs[l + 1];
      auto &lo = *fs[l];
      const long r = std::lround((double) lo.getGrid().cellSize / (double) hi.getGrid().cellSize);
      auto dl = lo.getGrid().desc(), dh = hi.getGrid().desc();
      check(gx_deinterp_add(lo.devicePointer(), &dl, hi.devicePointer(), &dh, dtype_of<T>(), 1.0 / (double) (r * r * r), nullptr));
    }
  }
  for (size_t l = 0; l < nl; ++l) *fs[l] *= ctx.grids->getWeightForLevel(l) / ctx.grids->getWeightForLevel(nl - 1);
  ret.applyFilters(filters::FilterFamily<T>(*ctx.grids));
  ret.toReal();
  return ret;
}

// Modification::Modification (modification.hpp:36-47)
template <typename T> void warnIfWholeZoomGridIsFlagged(const Context<T> &ctx) {
  for (size_t l = 1; l < ctx.getNumLevels(); ++l) {
    const auto &g = ctx.grids->getGridForLevel(l);
    std::vector<size_t> flags;
    g.getFlaggedCells(flags);
    if (flags.size() == g.size3) {
      logger()("WARNING: Region selected for modification is the entire zoom grid. This is likely because the cell selection extends "
               "beyond the zoom boundaries.");
      logger()("By design, modifications are meant to be defined inside a zoom region. Increase the size of your zoom grid or "
               "decrease your selection to avoid nasty surprises.");
    }
  }
}

// Grid::getFlaggedCellsPhysicalSize (grid.hpp:318-334): the largest side of the periodic window holding the flags
template <typename T> T flaggedCellsPhysicalSize(const grids::Grid<T> &g) {
  std::vector<size_t> flags;
  g.getFlaggedCells(flags);
  if (flags.empty()) return T(0);
  const long n = (long) g.size;
  auto coord = [n](size_t id) { return zoom::Int3{(long) (id / (size_t) (n * n)), (long) ((id / (size_t) n) % (size_t) n), (long) (id % (size_t) n)}; };
  zoom::IntWindow w((long) g.simEquivalentSize, coord(flags[0]));
  for (size_t id : flags) w.expandToInclude(coord(id));
  const zoom::Int3 s = w.sizes();
  return T(std::max(s[0], std::max(s[1], s[2]))) * g.cellSize;
}

// ---- linear modifications ---------------------------------------------------------------------------------------------
template <typename T> class LinearModification {
public:
  LinearModification(Context<T> &c, const cosmology::CosmologicalParameters<T> &cos) : ctx(&c), cosmo(cos) {
    ctx->grids->getGridForLevel(ctx->getNumLevels() - 1).getFlaggedCells(flaggedCellsFinestGrid);
    if (flaggedCellsFinestGrid.empty())
      throw std::runtime_error("Linear modifications use exclusively information from the highest resolution grid but no cells "
                               "have been flagged on this grid.\nChange flagged cells selection to use linear modifications.");
    warnIfWholeZoomGridIsFlagged(c);
  }
  virtual ~LinearModification() = default;
  int order() const { return 1; }
  void setTarget(T t) { target = t; }
  T getTarget() const { return target; }
  TypedField<T> getCovector(TransferType forSpecies) {  // linearmodification.hpp:60-76
    auto &g = ctx->grids->getGridForLevel(ctx->getNumLevels() - 1);
    TypedField<T> r = generateMultilevelCovectorFromHiresCovector<T>(*ctx, calculateLocalisationCovector(g), forSpecies);
    turnLocalisationCovectorIntoModificationCovector(r);
    r.toFourier();
    return r;
  }
  T calculateCurrentValue(const TypedField<T> &field) { return getCovector(field.getTransferType()).innerProduct(field); }

protected:
  virtual std::shared_ptr<fields::Field<T>> calculateLocalisationCovector(const grids::Grid<T> &g) {  // :78-90
    std::vector<T> data(g.size3, T(0));
    const T w = T(1.0) / T(flaggedCellsFinestGrid.size());
    for (size_t i : flaggedCellsFinestGrid) data[i] += w;
    auto f = std::make_shared<fields::Field<T>>(g, data, false);
    f->toFourier();
    return f;
  }
  virtual void turnLocalisationCovectorIntoModificationCovector(TypedField<T> &field) = 0;
  Context<T> *ctx;
  cosmology::CosmologicalParameters<T> cosmo;
  std::vector<size_t> flaggedCellsFinestGrid;
  T target = 0;
};

template <typename T> class OverdensityModification : public LinearModification<T> {
public:
  using LinearModification<T>::LinearModification;
protected:
  void turnLocalisationCovectorIntoModificationCovector(TypedField<T> &field) override {  // :122-125: the type is left unchanged
    field.applyTransferRatio(field.getTransferType(), TransferType::dm);
  }
};

template <typename T> class PotentialModification : public OverdensityModification<T> {  // :128-146, parameters.hpp:208-225
public:
  using OverdensityModification<T>::OverdensityModification;
protected:
  void turnLocalisationCovectorIntoModificationCovector(TypedField<T> &field) override {
    OverdensityModification<T>::turnLocalisationCovectorIntoModificationCovector(field);
    const auto &c = this->cosmo;
    const double prefac = 3. / 2. * (double) c.OmegaM0 / (double) c.scalefactor * 100. * 100. / 299792. / 299792.;
    for (size_t l = 0; l < field.getNumLevels(); ++l) {
      auto &f = field.getFieldForLevel(l);
      f.toFourier();
      check(gx_scale_inv_k2(f.devicePointer(), (int) f.getGrid().size, dtype_of<T>(), (double) f.getGrid().thisGridSize, -prefac, 1, nullptr));
    }
  }
};

// linearmodification.hpp:148-285, the real-space 4th-order derivative of the default build
template <typename T> class VelocityModification : public OverdensityModification<T> {
public:
  VelocityModification(Context<T> &c, const cosmology::CosmologicalParameters<T> &cos, int direction_)
      : OverdensityModification<T>(c, cos), direction(direction_) {}
protected:
  std::shared_ptr<fields::Field<T>> calculateLocalisationCovector(const grids::Grid<T> &g) override {
    const long n = (long) g.size, S = (long) g.simEquivalentSize;
    const T w = T(1.0) / T(this->flaggedCellsFinestGrid.size());
    const T a = T(-w / T(12.) / g.cellSize), b = T(w * T(2.) / T(3.) / g.cellSize);
    std::vector<T> data(g.size3, T(0));
    auto step = [&](std::array<long, 3> c, long s) {  // getIndexFromIndexAndStep wraps with the simulation period (grid.hpp:555-574)
      long d = c[direction] + s;
      if (d > S - 1) d -= S;
      if (d < 0) d += S;
      c[direction] = d;
      return c;
    };
    auto add = [&](const std::array<long, 3> &c, T v) { data[(size_t) ((c[0] * n + c[1]) * n + c[2])] += v; };
    for (size_t id : this->flaggedCellsFinestGrid) {
      const std::array<long, 3> c{(long) (id / (size_t) (n * n)), (long) ((id / (size_t) n) % (size_t) n), (long) (id % (size_t) n)};
      const auto m1 = step(c, -1), p1 = step(c, 1), m2 = step(m1, -1), p2 = step(p1, 1);
      add(m2, a); add(m1, b); add(p1, -b); add(p2, -a);
    }
    auto f = std::make_shared<fields::Field<T>>(g, data, false);
    f->toFourier();
    return f;
  }
  void turnLocalisationCovectorIntoModificationCovector(TypedField<T> &field) override {
    if (this->ctx->grids->levelsAreCombined)
      throw std::runtime_error("Cannot create velocity covector for recombined fields (i.e. after output generation is complete)");
    OverdensityModification<T>::turnLocalisationCovectorIntoModificationCovector(field);
    const double scale = (double) cosmology::zeldovichVelocityToOffsetRatio<T>(this->cosmo);
    for (size_t l = 0; l < field.getNumLevels(); ++l) {
      auto &f = field.getFieldForLevel(l);
      f.toFourier();
      check(gx_scale_inv_k2(f.devicePointer(), (int) f.getGrid().size, dtype_of<T>(), (double) f.getGrid().thisGridSize, -scale, 0, nullptr));
    }
  }
  int direction;
};

// linearmodification.hpp:288-411: specific angular momentum of the flagged Lagrangian patch about its centre,
// -(q - q_c) x grad(Phi) with the 4th-order real-space stencil
template <typename T> class AngMomentumModification : public OverdensityModification<T> {
public:
  AngMomentumModification(Context<T> &c, const cosmology::CosmologicalParameters<T> &cos, int direction_)
      : OverdensityModification<T>(c, cos), direction(direction_) {
    if (direction < 0 || direction > 2) throw std::runtime_error("Angular momentum direction must be 0 (x), 1 (y) or 2 (z)");
  }
protected:
  std::shared_ptr<fields::Field<T>> calculateLocalisationCovector(const grids::Grid<T> &g) override {
    const long n = (long) g.size, S = (long) g.simEquivalentSize;
    const auto &flags = this->flaggedCellsFinestGrid;
    const double dx = (double) g.cellSize, Lbox = (double) g.periodicDomainSize;
    auto coordOf = [n](size_t id) { return std::array<long, 3>{(long) (id / (size_t) (n * n)), (long) ((id / (size_t) n) % (size_t) n), (long) (id % (size_t) n)}; };
    auto pos = [&](long c, int a) { return ((double) c * dx + (double) g.offsetLower[a]) + dx / 2; };
    auto wrapped = [Lbox](double d) { if (d > Lbox / 2) d -= Lbox; if (d < -Lbox / 2) d += Lbox; return d; };
    double qc[3];  // Grid::getCentreWrapped (grid.hpp:466-494)
    {
      const auto c0 = coordOf(flags[0]);
      for (int a = 0; a < 3; ++a) {
        const double p0 = pos(c0[a], a);
        double run = 0;
        for (size_t i = 1; i < flags.size(); ++i) run += wrapped(pos(coordOf(flags[i])[a], a) - p0);
        run /= (double) flags.size();
        double x = std::fmod(run + p0, Lbox);
        qc[a] = x < 0 ? x + Lbox : x;
      }
    }
    const double norm = (double) flags.size() * dx, ca = 1. / 12. / norm, cb = -2. / 3. / norm;
    const int d = direction, d1 = (direction + 1) % 3, d2 = (direction + 2) % 3;
    std::vector<T> data(g.size3, T(0));
    auto step = [S](std::array<long, 3> c, int ax, long s) {
      long v = c[ax] + s;
      if (v > S - 1) v -= S;
      if (v < 0) v += S;
      c[ax] = v;
      return c;
    };
    auto add = [&](const std::array<long, 3> &c, double v) { data[(size_t) ((c[0] * n + c[1]) * n + c[2])] += (T) v; };
    for (int ax = 0; ax < 3; ++ax) {
      if (ax == d) continue;
      for (size_t id : flags) {
        const auto c = coordOf(id);
        double dq[3];
        for (int a = 0; a < 3; ++a) dq[a] = wrapped(qc[a] - pos(c[a], a));
        const double coeff = ax == d2 ? dq[d1] : -dq[d2];
        const auto m1 = step(c, ax, -1), p1 = step(c, ax, 1), m2 = step(m1, ax, -1), p2 = step(p1, ax, 1);
        add(m2, coeff * ca); add(m1, coeff * cb); add(p1, -coeff * cb); add(p2, -coeff * ca);
      }
    }
    auto f = std::make_shared<fields::Field<T>>(g, data, false);
    f->toFourier();
    return f;
  }
  void turnLocalisationCovectorIntoModificationCovector(TypedField<T> &field) override {
    OverdensityModification<T>::turnLocalisationCovectorIntoModificationCovector(field);
    const double scale = (double) cosmology::zeldovichVelocityToOffsetRatio<T>(this->cosmo);
    for (size_t l = 0; l < field.getNumLevels(); ++l) {
      auto &f = field.getFieldForLevel(l);
      f.toFourier();
      check(gx_scale_inv_k2(f.devicePointer(), (int) f.getGrid().size, dtype_of<T>(), (double) f.getGrid().thisGridSize, -scale, 0, nullptr));
    }
  }
  int direction;
};

// ---- quadratic modification: variance of the high-pass filtered field inside the flagged region -------------------------
template <typename T> class FilteredVarianceModification {
public:
  FilteredVarianceModification(Context<T> &c, const cosmology::CosmologicalParameters<T> &cos, T filterscale, int initSteps = 10,
                               T precision = T(0.001))
      : initNumberSteps(initSteps), targetPrecision(precision), ctx(&c), cosmo(cos), scale(filterscale) {
    if (filterscale < T(0))
      throw std::runtime_error("Trying to calculate filtered variance without initialising variance filtering scale. "
                               "Use filtering_scale command to do this.");
    const auto &fine = c.grids->getGridForLevel(c.getNumLevels() - 1);
    if (filterscale < fine.cellSize) throw std::runtime_error("Variance high-pass filtering scale is smaller than the smallest pixel size.");
    warnIfWholeZoomGridIsFlagged(c);
    const T window = flaggedCellsPhysicalSize(fine);
    if (filterscale > window) {  // checkFilterScale, quadraticmodification.hpp:128-149
      char buf[256];
      snprintf(buf, sizeof buf, "WARNING: High-pass filtering scale: %g h**-1 Mpc is greater than the rough window size: %g h**-1 Mpc used for modifications.",
               (double) filterscale, (double) window);
      logger()(buf);
      logger()("This is prone to numerical errors when modifying the field. Decrease filtering scale to avoid it.");
    }
    for (size_t l = 0; l < c.getNumLevels(); ++l) masks.emplace_back(new FlagMask<T>(c.grids->getGridForLevel(l)));
  }
  int order() const { return 2; }
  void setTarget(T t) { target = t; }
  T getTarget() const { return target; }
  T calculateCurrentValue(const TypedField<T> &field) {  // quadraticmodification.hpp:104-110
    TypedField<T> pushed = pushMultiLevelFieldThroughMatrix(field);
    pushed.toFourier();
    return pushed.innerProduct(field);
  }
  TypedField<T> pushMultiLevelFieldThroughMatrix(const TypedField<T> &field) {  // :112-126
    TypedField<T> result(field);
    result.applyTransferRatio(field.getTransferType(), TransferType::dm);
    for (size_t l = 0; l < result.getNumLevels(); ++l) pushOneLevel(result.getFieldForLevel(l), l);
    result.convertToCovector();
    result.applyTransferRatio(field.getTransferType(), TransferType::dm);
    return result;
  }
  int initNumberSteps;
  T targetPrecision;

private:
  void pushOneLevel(fields::Field<T> &f, size_t level) {  // pushOneLevelFieldThroughMatrix, :152-221
    const FlagMask<T> &mask = *masks[level];
    const filters::ComplementaryFilterAdaptor<T> hp(T(2. * M_PI) / scale);
    const int n = (int) f.getGrid().size;
    mulMask(f, mask);   // windowOperator
    f.toFourier();
    f.applyFilter(hp);  // filterOperator
    f.toReal();
    if (mask.count) {   // varianceOperator: window, subtract the mean over the flagged cells, divide by their number
      double s = 0;
      check(gx_masked_sum(f.devicePointer(), (const unsigned char *) mask.dev, n, dtype_of<T>(), &s, nullptr));
      check(gx_masked_shift_scale(f.devicePointer(), (const unsigned char *) mask.dev, n, dtype_of<T>(), s / (double) mask.count, 1.0 / (double) mask.count, nullptr));
    } else {
      mulMask(f, mask);
    }
    f.toFourier();
    f.applyFilter(hp);
    f.toReal();
    mulMask(f, mask);
  }
  Context<T> *ctx;
  cosmology::CosmologicalParameters<T> cosmo;
  T scale;
  T target = 0;
  std::vector<std::unique_ptr<FlagMask<T>>> masks;
};

// ---- ModificationManager (modificationmanager.hpp:20-385) ----------------------------------------------------------------
template <typename T> class ModificationManager {
public:
  ModificationManager(Context<T> &c, const cosmology::CosmologicalParameters<T> &cos, TypedField<T> &out)
      : ctx(&c), cosmo(cos), outputField(&out) {}
  T variance_filterscale = T(-1);
  int initial_number_steps = 10;
  T precision = T(0.001);

  T calculateCurrentValueByName(const std::string &name) {
    outputField->toFourier();
    std::string nm = lower(name);
    if (nm == "variance") return makeVariance()->calculateCurrentValue(*outputField);
    return makeLinear(nm)->calculateCurrentValue(*outputField);
  }
  void addModificationToList(const std::string &name, const std::string &type, T target) {  // :127-161
    const std::string nm = lower(name), ty = lower(type);
    if (ty != "relative" && ty != "absolute") throw std::runtime_error("Modification type must be either 'relative' or 'absolute'");
    if (nm == "variance") {
      auto m = makeVariance();
      if (ty == "relative") target = target * m->calculateCurrentValue(*outputField);
      m->setTarget(target);
      quadratic.push_back(std::move(m));
    } else {
      auto m = makeLinear(nm);
      if (ty == "relative") target = target * m->calculateCurrentValue(*outputField);
      m->setTarget(target);
      linear.push_back(std::move(m));
    }
  }
  bool hasModifications() const { return !linear.empty() || !quadratic.empty(); }
  void clearModifications() { linear.clear(); quadratic.clear(); }

  void applyModifications() {  // :90-125, 199-217
    std::vector<TypedField<T>> alphas;
    std::vector<T> targets;
    for (auto &m : linear) { alphas.push_back(m->getCovector(outputField->getTransferType())); targets.push_back(m->getTarget()); }
    orthonormalise(alphas, targets);
    applyLinear(alphas, targets);
    for (auto &q : quadratic) {
      const int n0 = q->initNumberSteps;
      TypedField<T> test(*outputField);
      performIterations(test, alphas, *q, n0);
      const T achieved = std::abs(q->calculateCurrentValue(test) - q->getTarget());
      const int nSteps = n0 * (int) std::ceil(std::sqrt((double) (achieved / (q->getTarget() * q->targetPrecision))));
      performIterations(*outputField, alphas, *q, nSteps > n0 ? nSteps : n0);
    }
  }

private:
  static std::string lower(std::string s) { for (auto &c : s) c = (char) std::tolower(c); return s; }
  std::unique_ptr<LinearModification<T>> makeLinear(const std::string &nm) {
    if (nm == "overdensity") return std::unique_ptr<LinearModification<T>>(new OverdensityModification<T>(*ctx, cosmo));
    if (nm == "potential") return std::unique_ptr<LinearModification<T>>(new PotentialModification<T>(*ctx, cosmo));
    if (nm == "vx" || nm == "vy" || nm == "vz")
      return std::unique_ptr<LinearModification<T>>(new VelocityModification<T>(*ctx, cosmo, nm[1] - 'x'));
    if (nm == "lx" || nm == "ly" || nm == "lz")
      return std::unique_ptr<LinearModification<T>>(new AngMomentumModification<T>(*ctx, cosmo, nm[1] - 'x'));
    throw std::runtime_error(nm + " is an unknown modification name");
  }
  std::unique_ptr<FilteredVarianceModification<T>> makeVariance() {
    return std::unique_ptr<FilteredVarianceModification<T>>(
        new FilteredVarianceModification<T>(*ctx, cosmo, variance_filterscale, initial_number_steps, precision));
  }
  static void orthonormalise(std::vector<TypedField<T>> &alphas, std::vector<T> &targets) {  // :249-277
    for (size_t i = 0; i < alphas.size(); ++i) {
      for (size_t j = 0; j < i; ++j) {
        const T r = alphas[i].innerProduct(alphas[j]);
        alphas[i].addScaled(alphas[j], -r);
        targets[i] -= r * targets[j];
      }
      const T norm = std::sqrt(alphas[i].innerProduct(alphas[i]));
      alphas[i] /= norm;
      targets[i] /= norm;
    }
  }
  void applyLinear(std::vector<TypedField<T>> &alphas, const std::vector<T> &targets) {  // :279-304
    std::vector<T> existing;
    for (auto &a : alphas) existing.push_back(a.innerProduct(*outputField));
    for (size_t i = 0; i < alphas.size(); ++i) {
      alphas[i].convertToVector();
      alphas[i].toFourier();
      outputField->addScaled(alphas[i], targets[i] - existing[i]);
    }
  }
  static void performIterations(TypedField<T> &field, std::vector<TypedField<T>> &alphas, FilteredVarianceModification<T> &q,
                                int nSteps) {  // :306-331
    // The reference evaluates the current value (a push through the matrix + an inner product) and then pushes the same
    // field again (:309-323); the second push repeats the first bit for bit, so one is kept.
    TypedField<T> pushed = q.pushMultiLevelFieldThroughMatrix(field);
    pushed.toFourier();
    T cur = pushed.innerProduct(field);
    const T start = cur, end = q.getTarget();
    std::vector<T> targets;  // tools::linspace (util_functions.hpp:78-92)
    if (nSteps == 0 || nSteps == 1 || start == end) targets.push_back(end);
    else {
      const T step = (end - start) / T(nSteps);
      for (int c = 0; c <= nSteps; ++c) targets.push_back(start + T(c) * step);
    }
    for (int i = 0; i < nSteps; ++i) {
      if (i > 0) {
        pushed = q.pushMultiLevelFieldThroughMatrix(field);
        pushed.toFourier();
        cur = pushed.innerProduct(field);
      }
      const T norm = std::sqrt(pushed.innerProduct(pushed));
      for (auto &a : alphas) {  // addToOrthonormalFamily: the alphas are vectors by now
        const T r = pushed.innerProduct(a);
        pushed.addScaled(a, -r);
      }
      pushed /= std::sqrt(pushed.innerProduct(pushed));
      const T mult = T(0.5) * (targets[(size_t) i + 1] - cur) / norm;
      pushed.convertToVector();
      field.addScaled(pushed, mult);
    }
  }
  Context<T> *ctx;
  cosmology::CosmologicalParameters<T> cosmo;
  TypedField<T> *outputField;
  std::vector<std::unique_ptr<LinearModification<T>>> linear;
  std::vector<std::unique_ptr<FilteredVarianceModification<T>>> quadratic;
};

// ---- splicing (splice.hpp:38-118) with tools::numerics::conjugateGradient (cg.hpp:14-58) -----------------------------------
template <typename T>
fields::Field<T> conjugateGradient(const std::function<fields::Field<T>(const fields::Field<T> &)> &Q, const fields::Field<T> &b,
                                   double rtol = 1e-6, double atol = 1e-12) {
  fields::Field<T> residual(b), direction(b), x(b.getGrid(), false);
  direction *= T(-1);
  const double scale = (double) b.norm();
  if (scale == 0.0) return x;
  const size_t dimension = b.getGrid().size3;
  for (size_t i = 0; i < dimension + 1; ++i) {
    fields::Field<T> Qd = Q(direction);
    const T alpha = -residual.innerProduct(direction) / direction.innerProduct(Qd);
    x.addScaled(direction, alpha);
    residual = Q(x);
    residual -= b;
    const double norm = (double) residual.norm();
    if (norm < rtol * scale || norm < atol) break;
    const T beta = residual.innerProduct(Qd) / direction.innerProduct(Qd);
    direction *= beta;
    direction -= residual;
  }
  return x;
}

// a, b: white-noise fields on one grid, cov: the covariance field; the grid's flagged cells are the mask. Returns the spliced
// field in real space.
template <typename T> fields::Field<T> spliceOneLevel(fields::Field<T> &a, fields::Field<T> &b, const fields::Field<T> &cov) {
  const auto &grid = a.getGrid();
  FlagMask<T> mask(grid, false), maskCompl(grid, true);
  a.toFourier();
  b.toFourier();
  fields::Field<T> pre(cov);
  pre.setFourierCoefficient(0, 0, 0, std::complex<T>(1, 0));  // splice.hpp:57-58
  auto Tf = [&](fields::Field<T> &f, double power) { f.toFourier(); f.applyTransferFunction(pre, power); f.toReal(); };
  fields::Field<T> delta_diff(b);
  delta_diff -= a;
  Tf(delta_diff, 0.5);
  fields::Field<T> z(delta_diff);
  mulMask(z, mask);
  Tf(z, -1.0);
  mulMask(z, maskCompl);
  Tf(z, 0.5);
  std::function<fields::Field<T>(const fields::Field<T> &)> X = [&](const fields::Field<T> &inp) {
    fields::Field<T> v(inp);
    Tf(v, 0.5);
    mulMask(v, maskCompl);
    Tf(v, -1.0);
    mulMask(v, maskCompl);
    Tf(v, 0.5);
    return v;
  };
  fields::Field<T> alpha = conjugateGradient<T>(X, z);
  Tf(alpha, 0.5);
  fields::Field<T> bInDeltaBasis(b);
  Tf(bInDeltaBasis, 0.5);
  mulMask(alpha, maskCompl);
  alpha += bInDeltaBasis;
  mulMask(delta_diff, mask);
  alpha -= delta_diff;
  Tf(alpha, -0.5);
  return alpha;
}

}  // namespace modifications

// ---------------------------------------------------------------------------------------------------------------------------
namespace ic {
// ZoomICGenerator plus the modification commands of a paramfile: filtering_scale, calculate, modify, clear_modifications,
// apply_modifications, chi2, splice, dump_grid, done (ic.hpp:1598-1648, 822-851, 1700-1762).
template <typename T> class ModifICGenerator : public ZoomICGenerator<T> {
public:
  using Base = ZoomICGenerator<T>;
  std::unique_ptr<cosmology::CAMB<T>> specDM, specAll;
  modifications::Context<T> mctx;
  std::unique_ptr<modifications::TypedField<T>> outputField;
  std::unique_ptr<modifications::ModificationManager<T>> manager;
  T variance_filterscale = T(-1);
  // mapper_relative_to (ic.hpp:1077-1112, dummyic.hpp): the ids of later id files refer to the setup of another parameter file,
  // whose geometry is rebuilt by a field-less run of its commands on shared grid objects
  std::unique_ptr<ModifICGenerator<T>> pInput;
  typename particle::mapper::ParticleMapper<T>::Ptr pInputMapper;
  size_t subsample = 1, supersample = 1;      // grafic output only in this layer (virtual levels of the centred stack)
  T pvarValue = T(1);                         // ic.hpp:148
  bool centerOnTargetRegion = false;          // ic.hpp:126, `center_output`
  static void log(const std::string &s) { modifications::logger()(s); }

  std::array<T, 3> graficCentre() {           // ic.hpp:1313-1319
    if (centerOnTargetRegion) return this->centre;
    const T half = this->context.getGridForLevel(0).thisGridSize / T(2);
    return {half, half, half};
  }
  // ParticleMapper::extendParticleListToUnreferencedGrids (mapper.hpp:183-196): every grid of `ctx` the mapper does not use
  // receives the flags of the mapper's finest grid, re-expressed on it
  static void extendParticleListToUnreferencedGrids(particle::mapper::ParticleMapper<T> &m, multilevelgrid::MultiLevelGrid<T> &ctx) {
    const auto layers = m.layers();
    for (size_t level = 0; level < ctx.getNumLevels(); ++level) {
      auto &g = ctx.getGridForLevel(level);
      bool referenced = false;
      for (auto &kv : layers) referenced = referenced || kv.second->grid == &g;
      if (referenced) continue;
      std::vector<size_t> add = zoom::proxyFlags<T>(m.getFinestGrid(), g), have;
      g.getFlaggedCells(have);
      have.insert(have.end(), add.begin(), add.end());
      std::sort(have.begin(), have.end());
      have.erase(std::unique(have.begin(), have.end()), have.end());
      g.flagCells(have);
    }
  }

  // ICGenerator::getCentre (ic.hpp:1372-1386) + Grid::getCentreWrapped (grid.hpp:466-494): after an id file has been loaded
  // the current point is the centre of the flagged cells of the deepest level that has any
  void updateCentreFromFlags() {
    size_t level = 0;
    std::vector<size_t> flags;
    for (size_t l = this->context.getNumLevels(); l-- > 0;) {
      flags.clear();
      this->context.getGridForLevel(l).getFlaggedCells(flags);
      if (!flags.empty()) { level = l; break; }
    }
    if (flags.empty()) throw std::runtime_error("Cannot calculate the center of an empty region");
    std::sort(flags.begin(), flags.end());
    const auto &g = this->context.getGridForLevel(level);
    const size_t n = g.size;
    const double Lbox = (double) g.periodicDomainSize, dx = (double) g.cellSize;
    for (int a = 0; a < 3; ++a) {
      auto pos = [&](size_t id) {
        const size_t c = a == 0 ? id / (n * n) : (a == 1 ? (id / n) % n : id % n);
        return ((double) c * dx + (double) g.offsetLower[a]) + dx / 2;
      };
      const double p0 = pos(flags[0]);
      double run = 0;
      for (size_t i = 1; i < flags.size(); ++i) {
        double d = pos(flags[i]) - p0;
        if (d > Lbox / 2) d -= Lbox;
        if (d < -Lbox / 2) d += Lbox;
        run += d;
      }
      run /= (double) flags.size();
      double x = std::fmod(run + p0, Lbox);
      if (x < 0) x += Lbox;
      this->centre[a] = (T) x;
    }
    char buf[160];
    snprintf(buf, sizeof buf, "Centre of region is %.12g %.12g %.12g", (double) this->centre[0], (double) this->centre[1], (double) this->centre[2]);
    log(buf);
  }

  void initialiseRandomComponentIfUninitialised() {  // ic.hpp:1238-1259
    if (outputField) return;
    if (this->noiseMode < 0) throw std::runtime_error("The random number generator has not been seeded");
    specDM.reset(new cosmology::CAMB<T>(this->cosmo, this->cambFile, cosmology::species::dm));
    specAll.reset(new cosmology::CAMB<T>(this->cosmo, this->cambFile, cosmology::species::all));
    mctx.grids = &this->context; mctx.dm = specDM.get(); mctx.all = specAll.get();
    outputField.reset(new modifications::TypedField<T>(mctx, modifications::TransferType::whitenoise));
    fields::RandomFieldGenerator<T> rfg(outputField->getFields());
    rfg.seed(this->seed);
    rfg.setDrawInFourierSpace(this->noiseMode != GX_NOISE_REAL);
    rfg.setParallel(this->noiseMode == GX_NOISE_FOURIER_PARALLEL);
    rfg.setReverseRandomDrawOrder(this->reverse);
    rfg.draw();
    if (this->fixPower) for (auto &f : outputField->getFields()) f->enforceUnitVariance();
    manager.reset(new modifications::ModificationManager<T>(mctx, this->cosmo, *outputField));
    manager->variance_filterscale = variance_filterscale;
  }

  // returns true if the command was one of ours
  bool runModifCommand(const std::string &cmd, std::istringstream &ss) {
    auto num = [&]() { double v; if (!(ss >> v)) throw std::runtime_error("Insufficient arguments for command " + cmd); return v; };
    auto str = [&]() { std::string v; if (!(ss >> v)) throw std::runtime_error("Insufficient arguments for command " + cmd); return v; };
    char buf[256];
    auto path = [this](const std::string &p) { return p[0] == '/' ? p : this->dir + "/" + p; };
    const bool isGrafic = this->outputFormat == "grafic";
    if (this->pUnderlying) {  // DummyICGenerator (dummyic.hpp:60-120): everything that touches fields or writes output is a no-op
      static const char *noop[] = {"zerolevel", "zero_level", "dump_grid", "dump_ps", "dump_mask", "dump_idfile", "dump_id_file", "modify",
                                   "done", "apply_modifications", "calculate", "chi2", "reverse", "reverse_small_k", "import_level",
                                   "dump_tipsy", "dump_grid_fourier", "splice", "dump_vx"};
      for (const char *c : noop) if (cmd == c) return true;
    }
    if (cmd == "mapper_relative_to") {
      const std::string given = str(), fname = path(given);
      if (this->context.getNumLevels() == 0) throw std::runtime_error("Mapper relative command cannot be used before a basegrid has been initialised.");
      { std::ifstream probe(fname); if (!probe.is_open()) throw std::runtime_error("Cannot open IC paramfile for relative_to command"); }
      log("");
      log("+ Input mapper: computing geometry from " + given);
      pInput.reset(new ModifICGenerator<T>());
      pInput->pUnderlying = this;
      {
        const std::string saved = logIndent();
        logIndent() = saved + "| ";
        log("");
        try { pInput->runParamFile(fname); } catch (...) { logIndent() = saved; throw; }
        logIndent() = saved;
      }
      log("");
      if (pInput->outputFormat == "grafic") throw std::runtime_error("mapper_relative_to a grafic setup is not implemented in the C++ layer");
      if (pInput->subsample != 1 || pInput->supersample != 1)
        throw std::runtime_error("mapper_relative_to a sub/supersampled setup is not implemented in the C++ layer");
      pInputMapper = pInput->mapper();
      return true;
    }
    if (cmd == "outformat") {
      std::string f = str();
      for (auto &c : f) c = (char) std::tolower(c);
      if (f == "grafic" || f == "5") this->outputFormat = "grafic";
      else { std::istringstream again(f); Base::runCommand(cmd, again); }
    }
    else if (cmd == "zerolevel" || cmd == "zero_level") {  // ic.hpp:1650-1670, for tests
      const size_t level = (size_t) num();
      snprintf(buf, sizeof buf, "*** Warning: your script calls zeroLevel(%zu). This is intended for testing purposes only!", level);
      log(buf);
      initialiseRandomComponentIfUninitialised();
      if (level + 1 > outputField->getNumLevels()) throw std::runtime_error("Specified level does not exist.");
      auto &f = outputField->getFieldForLevel(level);
      check(gx_memset_zero(f.devicePointer(), gx_field_bytes((int) f.getGrid().size, dtype_of<T>()), nullptr));
    }
    else if (cmd == "reverse") {  // MultiLevelField::reverse, multilevelfield.hpp:182-191
      initialiseRandomComponentIfUninitialised();
      for (size_t l = 0; l < outputField->getNumLevels(); ++l) outputField->getFieldForLevel(l) *= T(-1);
    }
    else if (cmd == "reverse_small_k") {  // ic.hpp:1690-1708
      const double kmax = num();
      initialiseRandomComponentIfUninitialised();
      for (size_t l = 0; l < outputField->getNumLevels(); ++l) {
        auto &f = outputField->getFieldForLevel(l);
        f.toFourier();
        check(gx_reverse_small_k(f.devicePointer(), (int) f.getGrid().size, dtype_of<T>(), (double) f.getGrid().thisGridSize, kmax, nullptr));
      }
    }
    else if (cmd == "velocity_offset") {  // ic.hpp:285-289: km/s -> km a^1/2 s^-1
      const T conversion = (T) std::pow((double) this->cosmo.scalefactor, -0.5);
      for (int a = 0; a < 3; ++a) this->velOffset[a] = (T) num() * conversion;
    }
    else if (cmd == "import_level") {  // ic.hpp:809-821 + Field::loadGridData
      const size_t level = (size_t) num();
      const std::string fn = path(str());
      initialiseRandomComponentIfUninitialised();
      log("Importing random field on level " + std::to_string(level) + " from " + fn);
      auto &f = outputField->getFieldForLevel(level);
      std::vector<T> data = grafic::loadNpy<T>(fn, f.getGrid().size);
      f.setFourier(false);
      f.getDataVector() = data;
      f.toFourier();
      f.applyTransferFunction(*specDM, -0.5);  // back to white noise
      log("... success!");
    }
    else if (cmd == "dump_tipsy") dumpTipsy(str());
    else if (cmd == "adapt_mask") adaptMask((size_t) num());
    else if (cmd == "subsample") subsample = (size_t) num();
    else if (cmd == "supersample") supersample = (size_t) num();
    else if (cmd == "pvar") pvarValue = (T) num();
    else if (cmd == "center_output" || cmd == "centre_output" || cmd == "center_grafic_output") centerOnTargetRegion = true;
    else if (cmd == "dump_mask")
      grafic::dumpMask<T>(this->context, graficCentre(), subsample, supersample, this->zoomParticleArray, this->outputFolder);
    else if (cmd == "idfile" || cmd == "id_file" || cmd == "append_idfile" || cmd == "merge_id_file") {
      const bool replace = cmd == "idfile" || cmd == "id_file";
      const std::vector<size_t> ids = Base::readIds(path(str()));
      if (isGrafic) {  // ensureParticleGeneratorInitialised (ic.hpp:1140-1147): ids are the cells of the grafic levels in turn
        grafic::GraficMapper<T> m(this->context, this->centre, subsample, supersample);
        if (replace) m.unflagAllParticles();
        m.flagParticles(ids);
      } else if (subsample != 1 || supersample != 1) {
        throw std::runtime_error("subsample / supersample are implemented for grafic output only in the C++ layer");
      } else if (pInputMapper) {
        // flagCellsCorrespondingToParticles (ic.hpp:1247-1261): the ids refer to the other setup; flag there, then carry the
        // flags onto the grids only this setup has, and back
        auto mine = this->mapper();
        if (replace) { mine->unflagAllParticles(); pInputMapper->unflagAllParticles(); }
        pInputMapper->flagParticles(ids);
        extendParticleListToUnreferencedGrids(*pInputMapper, this->context);
        extendParticleListToUnreferencedGrids(*mine, pInput->context);
      } else {
        auto m = this->mapper();
        if (replace) m->unflagAllParticles();
        m->flagParticles(ids);
      }
      updateCentreFromFlags();
    }
    else if ((cmd == "dump_idfile" || cmd == "dump_id_file") && isGrafic) {
      std::vector<size_t> ids;
      grafic::GraficMapper<T>(this->context, this->centre, subsample, supersample).getFlaggedParticles(ids);
      std::ofstream o(path(str()));
      for (size_t i : ids) o << i << std::endl;
    }
    else if (cmd == "filtering_scale") { variance_filterscale = (T) num(); if (manager) manager->variance_filterscale = variance_filterscale; }
    else if (cmd == "calculate") {
      const std::string name = str();
      initialiseRandomComponentIfUninitialised();
      snprintf(buf, sizeof buf, "%s: calculated value = %g", name.c_str(), (double) manager->calculateCurrentValueByName(name));
      log(buf);
    }
    else if (cmd == "modify") {
      const std::string name = str(), type = str();
      const float target = (float) num();  // ICGenerator::modify(string, string, float target), ic.hpp:1606
      initialiseRandomComponentIfUninitialised();
      manager->addModificationToList(name, type, (T) target);
    }
    else if (cmd == "clear_modifications") { log("Clearing modification list"); if (manager) manager->clearModifications(); }
    else if (cmd == "apply_modifications") { initialiseRandomComponentIfUninitialised(); applyModificationsLogged(); }
    else if (cmd == "chi2") {
      initialiseRandomComponentIfUninitialised();
      snprintf(buf, sizeof buf, "Calculated chi^2 = %.10g (dof = %zu)", (double) outputField->getChi2(), this->context.getNumDof());
      log(buf);
    }
    else if (cmd == "dump_grid") { dumpGrid((size_t) num()); }
    else if (cmd == "dump_ps") { dumpPS((size_t) num()); }
    else if (cmd == "dump_grid_fourier") { dumpGridFourier((size_t) num()); }
    else if (cmd == "splice") { splice((unsigned long) num()); }
    else if (cmd == "done") { doneWithModifications(); }
    else return false;
    return true;
  }

  void runParamFile(const std::string &filename) {
    std::ifstream f(filename);
    if (!f.is_open()) throw std::runtime_error("Paramfile " + filename + " not found");
    this->dir = filename.find('/') == std::string::npos ? "." : filename.substr(0, filename.rfind('/'));
    std::string line;
    int lineno = 0;
    while (std::getline(f, line)) {
      ++lineno;
      std::istringstream ss(line);
      std::string cmd;
      if (!(ss >> cmd) || cmd[0] == '#' || cmd[0] == '%') continue;
      for (auto &c : cmd) c = (char) std::tolower(c);
      try {
        if (!runModifCommand(cmd, ss)) Base::runCommand(cmd, ss);
      } catch (std::exception &e) {
        throw std::runtime_error("Error \"" + std::string(e.what()) + "\" on line " + std::to_string(lineno) + " (\"" + line + "\")");
      }
    }
  }

  // ModificationManager::applyModifications with its chi^2 bookkeeping (modificationmanager.hpp:93-125)
  void applyModificationsLogged() {
    const T pre = outputField->getChi2();
    log("");
    log("Applying modifications");
    manager->applyModifications();
    const T post = outputField->getChi2();
    char buf[160];
    snprintf(buf, sizeof buf, "   Post-modification chi^2 = %g", (double) post); log(buf);
    snprintf(buf, sizeof buf, "  Modification Delta chi^2 = %g", (double) (post - pre)); log(buf);
    snprintf(buf, sizeof buf, "           d.o.f. in field = %zu", this->context.getNumDof()); log(buf);
    log("");
  }

  // ICGenerator::adaptMask (ic.hpp:1545-1575): working up from the deepest zoom, the cells each level replaces become the
  // footprint of the level below's zoom cells on it, padded by nPadCells (Grid::expandFlaggedRegion, grid.hpp:277-306)
  void adaptMask(size_t nPadCells) {
    const size_t nl = this->context.getNumLevels();
    if (nl < 3) throw std::runtime_error("Adapting the mask requires there to be more than one zoom level");
    auto &z = this->zoomParticleArray;
    for (size_t level = nl - 2; level >= 1; --level) {
      const auto &cur = this->context.getGridForLevel(level);
      const auto &above = this->context.getGridForLevel(level - 1);
      const long na = (long) above.size;
      // the parents of this level's zoom cells on the level above (makeProxyGridToMatch: periodic section + subsample)
      grids::Grid<T> src(cur.periodicDomainSize, cur.size, cur.cellSize, cur.offsetLower[0], cur.offsetLower[1], cur.offsetLower[2]);
      src.flagCells(z[level]);
      std::vector<size_t> cells = zoom::proxyFlags<T>(src, above);
      auto uniq = [](std::vector<size_t> &v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); };
      uniq(cells);
      for (size_t it = 0; it < nPadCells; ++it)
        for (int axis = 0; axis < 3; ++axis) {
          const size_t have = cells.size();
          for (size_t k = 0; k < have; ++k) {
            long c[3] = {(long) (cells[k] / (size_t) (na * na)), (long) ((cells[k] / (size_t) na) % (size_t) na), (long) (cells[k] % (size_t) na)};
            for (long stepv : {1L, -1L}) {
              long d[3] = {c[0], c[1], c[2]};
              d[axis] = ((d[axis] + stepv) % na + na) % na;
              cells.push_back((size_t) ((d[0] * na + d[1]) * na + d[2]));
            }
          }
          uniq(cells);
        }
      z[level - 1] = cells;
      if (level == 1) break;
    }
  }

  // ic.hpp:853-862 + io::tipsy::saveFieldTipsyArray (tipsy.hpp:38-54): the overdensity of every particle's cell in mapper order
  void dumpTipsy(const std::string &filename) {
    initialiseRandomComponentIfUninitialised();
    auto m = this->mapper();
    const particle::mapper::Layout &lay = m->layout();
    std::vector<float> out(lay.size());
    for (auto &kv : m->layers()) {
      auto &f = outputField->getFieldForLevel(kv.second->level);
      f.toReal();
      const std::vector<T> &data = f.getDataVector();
      for (size_t i = 0; i < lay.size(); ++i)
        if (lay.layer[i] == kv.first) out[i] = (float) data[(size_t) lay.cell[i]];
    }
    std::ofstream o(this->outputFolder + "/" + filename, std::ios::binary);
    const unsigned n = (unsigned) out.size();
    o.write((const char *) &n, 4);
    o.write((const char *) out.data(), (std::streamsize) (out.size() * sizeof(float)));
  }

  void doneWithModifications() {  // ICGenerator::done, ic.hpp:1633-1648
    if (this->outputFormat.empty()) throw std::runtime_error("No output format specified!");
    initialiseRandomComponentIfUninitialised();
    if (manager->hasModifications()) applyModificationsLogged();
    if (outputField->getTransferType() == modifications::TransferType::whitenoise)
      outputField->applyPowerSpectrumFor(modifications::TransferType::all);
    const std::string base0 = this->outputFolder + "/" + this->outputFilename;
    if (this->outputFormat == "grafic") {  // ic.hpp:1312-1324: the grids in turn, no particle mapper
      particle::MultiLevelParticleGenerator<T> ggen(this->context, outputField->getFields());
      outputField->toReal();
      grafic::save<T>(base0 + ".grafic", this->context, ggen, outputField->getFields(), this->cosmo, pvarValue, graficCentre(),
                      subsample, supersample, this->zoomParticleArray);
      return;
    }
    if (subsample != 1 || supersample != 1)
      throw std::runtime_error("subsample / supersample are implemented for grafic output only in the C++ layer");
    this->pMapper = this->mapper();
    particle::MultiLevelParticleGenerator<T> gen(this->context, outputField->getFields());
    const double boxlen = (double) this->context.getGridForLevel(0).periodicDomainSize;
    const std::string base = this->outputFolder + "/" + this->outputFilename;
    const bool offset = this->velOffset[0] != T(0) || this->velOffset[1] != T(0) || this->velOffset[2] != T(0);
    if (this->outputFormat == "tipsy") {
      particle::MappedParticles<T, T> P(*this->pMapper, &gen, this->cosmo);
      if (offset)  // OffsetEvaluator (offsetgenerator.hpp:38-43)
        for (size_t i = 0; i < P.vel.size(); ++i) P.vel[i] += this->velOffset[i % 3];
      io::tipsy::saveMapped<T>(base + ".tipsy", boxlen, P, this->cosmo);
    } else {
      if (offset) throw std::runtime_error("velocity_offset is implemented for tipsy output only in the C++ layer");
      particle::MappedParticles<T, float> P(*this->pMapper, &gen, this->cosmo);
      io::gadget::save<T>(base + ".gadget", boxlen, P, this->cosmo, this->outputFormat == "gadget2" ? 2 : 3, this->nGadgetFiles);
    }
  }

  void splice(unsigned long newSeed) {  // ICGenerator::splice, ic.hpp:1700-1762 (one level)
    initialiseRandomComponentIfUninitialised();
    if (this->context.getNumLevels() != 1) throw std::runtime_error("Splicing is implemented for a single level here");
    if (outputField->getTransferType() != modifications::TransferType::whitenoise)
      throw std::runtime_error("It is too late in the IC generation process to perform splicing; try moving the splice command earlier");
    auto &grid = this->context.getGridForLevel(0);
    fields::Field<T> other(grid, true);
    {
      std::vector<std::shared_ptr<fields::Field<T>>> lv{std::shared_ptr<fields::Field<T>>(&other, [](fields::Field<T> *) {})};
      fields::RandomFieldGenerator<T> rfg(lv);
      rfg.seed(newSeed);
      rfg.setDrawInFourierSpace(this->noiseMode != GX_NOISE_REAL);
      rfg.setParallel(this->noiseMode == GX_NOISE_FOURIER_PARALLEL);
      rfg.setReverseRandomDrawOrder(this->reverse);
      rfg.draw();
    }
    fields::Field<T> cov(grid, true);
    check(gx_covariance_field(cov.devicePointer(), (int) grid.size, dtype_of<T>(), (double) grid.thisGridSize, specDM->handle, nullptr));
    auto &orig = outputField->getFieldForLevel(0);
    fields::Field<T> spliced = modifications::spliceOneLevel<T>(other, orig, cov);
    spliced.toFourier();
    outputField->getFields()[0] = std::make_shared<fields::Field<T>>(spliced);
  }

  // ic.hpp:879-885 + getComplexFourierField (fourier.hpp:685-716): the full complex n^3 Fourier cube as .npy ('<c16' / '<c8')
  void dumpGridFourier(size_t level) {
    initialiseRandomComponentIfUninitialised();
    auto &f = outputField->getFieldForLevel(level);
    f.toFourier();
    const auto &g = f.getGrid();
    const size_t n = g.size, nc = n / 2 + 1;
    const std::vector<T> &half = f.getDataVector();  // [n][n][nc] complex, interleaved
    std::vector<T> full(2 * n * n * n);
    for (size_t x = 0; x < n; ++x)
      for (size_t y = 0; y < n; ++y)
        for (size_t z = 0; z < n; ++z) {
          const bool neg = z >= nc;  // kz < 0: the conjugate of the mode (-kx, -ky, -kz)
          const size_t sx = neg ? (n - x) % n : x, sy = neg ? (n - y) % n : y, sz = neg ? n - z : z;
          const size_t s = 2 * ((sx * n + sy) * nc + sz), d = 2 * ((x * n + y) * n + z);
          full[d] = half[s];
          full[d + 1] = neg ? -half[s + 1] : half[s + 1];
        }
    const std::string base = this->outputFolder + "/grid-" + std::to_string(level);
    std::ofstream o(base + ".npy", std::ios::binary);
    char dict[128];
    snprintf(dict, sizeof dict, "{'descr': '<c%d', 'fortran_order': False, 'shape': (%zu, %zu, %zu), }", (int) (2 * sizeof(T)), n, n, n);
    std::string h(dict);
    const size_t total = 10 + h.size() + 1;
    h.append((64 - total % 64) % 64, ' ');
    h.push_back('\n');
    const unsigned short hl = (unsigned short) h.size();
    o.write("\x93NUMPY\x01\x00", 8);
    o.write(reinterpret_cast<const char *>(&hl), 2);
    o.write(h.data(), (std::streamsize) h.size());
    o.write(reinterpret_cast<const char *>(full.data()), (std::streamsize) (full.size() * sizeof(T)));
    std::ofstream t(this->outputFolder + "/grid-info-" + std::to_string(level) + "-fourier.txt");
    char buf[256];
    snprintf(buf, sizeof buf, "%g %g %g %g\n", (double) g.offsetLower[0], (double) g.offsetLower[1], (double) g.offsetLower[2], (double) g.thisGridSize);
    t << buf << "The line above contains information about grid level " << level << "\n"
      << "It gives the x-offset, y-offset and z-offset of the low-left corner and also the box length\n";
  }

  void dumpPS(size_t level, int nBins = 100) {  // ic.hpp:943-962 + cosmology::dumpPowerSpectrum, parameters.hpp:172-205
    initialiseRandomComponentIfUninitialised();
    auto &f = outputField->getFieldForLevel(level);
    f.toFourier();
    const auto &g = f.getGrid();
    const int n = (int) g.size;
    const double L = (double) g.thisGridSize;
    fields::Field<T> cov(g, true);
    check(gx_covariance_field(cov.devicePointer(), n, dtype_of<T>(), L, specAll->handle, nullptr));
    std::vector<double> bins(4 * (size_t) nBins, 0.0);  // count, sum k, sum |delta_k|^2, sum C(k): the reduction runs on the device
    check(gx_power_spectrum_bins(f.devicePointer(), cov.devicePointer(), n, dtype_of<T>(), L, nBins, bins.data(), nullptr));
    const double kmax = M_PI / L * n, kmin = 2.0 * M_PI / L, dklog = std::log10(kmax / kmin) / nBins, kw = 2.0 * M_PI / L;
    const double psnorm = 1.0 / ((kw * kw * kw / (double) std::pow(2.0f * (float) M_PI, 3.0f)) * (double) n * n * n);
    std::FILE *o = std::fopen((this->outputFolder + "/" + this->outputFilename + "_" + std::to_string(level) + ".ps").c_str(), "w");
    if (!o) throw std::runtime_error("Could not open the power spectrum file for writing");
    for (int i = 0; i < nBins; ++i) {
      const double c = bins[i];
      if (c <= 0) continue;
      const double col[5] = {std::pow(10.0, std::log10(kmin) + dklog * (i + 0.5)), bins[nBins + i] / c, bins[3 * nBins + i] / c * psnorm,
                             bins[2 * nBins + i] / c * psnorm, c};
      for (double v : col) { char num[64]; snprintf(num, sizeof num, "%g", v); std::fprintf(o, "%16s", num); }
      std::fprintf(o, "\n");
    }
    std::fclose(o);
  }

  void dumpGrid(size_t level) {  // ic.hpp:822-851 + io/numpy.hpp: the n^3 array as .npy and the info text file
    initialiseRandomComponentIfUninitialised();
    auto &f = outputField->getFieldForLevel(level);
    f.toReal();
    const auto &g = f.getGrid();
    const std::vector<T> &data = f.getDataVector();
    const std::string base = this->outputFolder + "/grid-" + std::to_string(level);
    {
      std::ofstream o(base + ".npy", std::ios::binary);
      char dict[128];
      snprintf(dict, sizeof dict, "{'descr': '<f%d', 'fortran_order': False, 'shape': (%zu, %zu, %zu), }", (int) sizeof(T), g.size, g.size, g.size);
      std::string h(dict);
      const size_t total = 10 + h.size() + 1;
      h.append((64 - total % 64) % 64, ' ');
      h.push_back('\n');
      const unsigned short hl = (unsigned short) h.size();
      o.write("\x93NUMPY\x01\x00", 8);
      o.write(reinterpret_cast<const char *>(&hl), 2);
      o.write(h.data(), (std::streamsize) h.size());
      o.write(reinterpret_cast<const char *>(data.data()), (std::streamsize) (g.size3 * sizeof(T)));
    }
    std::ofstream t(this->outputFolder + "/grid-info-" + std::to_string(level) + ".txt");
    char buf[256];
    snprintf(buf, sizeof buf, "%g %g %g %g\n", (double) g.offsetLower[0], (double) g.offsetLower[1], (double) g.offsetLower[2], (double) g.thisGridSize);
    t << buf << "The line above contains information about grid level " << level << "\n"
      << "It gives the x-offset, y-offset and z-offset of the low-left corner and also the box length\n";
  }
};
}  // namespace ic
}  // namespace gx
#endif
