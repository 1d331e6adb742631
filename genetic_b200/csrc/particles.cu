// Particle emit for a unigrid: ZeldovichParticleEvaluator::getParticle (zeldovich.hpp:87-119,
// generator.hpp:49-53), Grid::getCentroidFromCoordinate / wrapPoint (grid.hpp:519-531,656-669).
#include "../../include/genetic_b200.h"
#include "spectral.cuh"

namespace gx {

template <typename T, typename O>
__global__ void __launch_bounds__(256)
emit_component_kernel(const T *__restrict__ psi, int n, int axis, T dx, T lower, T domain, T ratio,
                      O *__restrict__ out_pos, O *__restrict__ out_vel, long stride, long long id0, long long count,
                      const long long *__restrict__ cells) {
  const long long n2 = (long long) n * n;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long) gridDim.x * blockDim.x) {
    const long long id = cells ? cells[i] : id0 + i;   // a mapper's cell list, or a contiguous id range
    const int x = (int) (id / n2);
    const long long rem = id - (long long) x * n2;
    const int y = (int) (rem / n);
    const int z = (int) (rem - (long long) y * n);
    const T p = psi[((size_t) x * n + y) * (size_t) (n + 2) + z];
    const int c = axis == 0 ? x : (axis == 1 ? y : z);
    // result = coord*cellSize; += offsetLower; += cellSize/2   (three separately rounded operations)
    T cen = mul_rn((T) c, dx);
    cen = add_rn(cen, lower);
    cen = add_rn(cen, dx / (T) 2);
    const T q = wrap_coordinate<T>(add_rn(p, cen), domain);
    if (out_pos) out_pos[i * stride] = (O) q;
    if (out_vel) out_vel[i * stride] = (O) mul_rn(p, ratio);
  }
}

template <typename T, typename O>
static int emit_t(const void *psi, int n, int axis, double dx, double lower, double domain, double ratio, void *pos,
                  void *vel, long stride, long long id0, long long count, cudaStream_t st,
                  const long long *cells = nullptr) {
  if (count <= 0) return 0;
  long long blocks = (count + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  GX_TIMED("emit", st);
  emit_component_kernel<T, O><<<(unsigned) blocks, 256, 0, st>>>((const T *) psi, n, axis, (T) dx, (T) lower, (T) domain,
                                                                (T) ratio, (O *) pos, (O *) vel, stride, id0, count, cells);
  count_launch();
  GX_LAUNCH_CHECK();
  return 0;
}

}  // namespace gx

extern "C" int gx_emit_component(const void *psi, int n, int dtype, int axis, double cell_size, double offset_lower,
                                 double domain_size, double vel_ratio, void *out_pos, void *out_vel, int out_dtype,
                                 long stride_elems, long long id0, long long count, void *stream) {
  using namespace gx;
  if (axis < 0 || axis > 2) return fail("axis must be 0, 1 or 2");
  const long long n3 = (long long) n * n * n;
  if (id0 < 0 || id0 + count > n3) return fail("gx_emit_component: id range outside the grid");
  cudaStream_t st = (cudaStream_t) stream;
  if (dtype == GX_F64 && out_dtype == GX_F64)
    return emit_t<double, double>(psi, n, axis, cell_size, offset_lower, domain_size, vel_ratio, out_pos, out_vel, stride_elems, id0, count, st);
  if (dtype == GX_F64 && out_dtype == GX_F32)
    return emit_t<double, float>(psi, n, axis, cell_size, offset_lower, domain_size, vel_ratio, out_pos, out_vel, stride_elems, id0, count, st);
  if (dtype == GX_F32 && out_dtype == GX_F32)
    return emit_t<float, float>(psi, n, axis, cell_size, offset_lower, domain_size, vel_ratio, out_pos, out_vel, stride_elems, id0, count, st);
  if (dtype == GX_F32 && out_dtype == GX_F64)
    return emit_t<float, double>(psi, n, axis, cell_size, offset_lower, domain_size, vel_ratio, out_pos, out_vel, stride_elems, id0, count, st);
  return fail("bad dtype");
}

// The gather behind the particle mappers (TwoLevelParticleMapper::dereferenceIterator -> ParticleEvaluator::getParticle,
// twolevelmapper.hpp:846-852, generator.hpp:49-53): the same evaluation at a device list of cell indices.
extern "C" int gx_emit_cells(const void *psi, int n, int dtype, int axis, double cell_size, double offset_lower,
                             double domain_size, double vel_ratio, const long long *cells, long long count,
                             void *out_pos, void *out_vel, int out_dtype, long stride_elems, void *stream) {
  using namespace gx;
  if (axis < 0 || axis > 2) return fail("axis must be 0, 1 or 2");
  if (count > 0 && !cells) return fail("gx_emit_cells: null cell list");
  cudaStream_t st = (cudaStream_t) stream;
  if (dtype == GX_F64 && out_dtype == GX_F64)
    return emit_t<double, double>(psi, n, axis, cell_size, offset_lower, domain_size, vel_ratio, out_pos, out_vel, stride_elems, 0, count, st, cells);
  if (dtype == GX_F64 && out_dtype == GX_F32)
    return emit_t<double, float>(psi, n, axis, cell_size, offset_lower, domain_size, vel_ratio, out_pos, out_vel, stride_elems, 0, count, st, cells);
  if (dtype == GX_F32 && out_dtype == GX_F32)
    return emit_t<float, float>(psi, n, axis, cell_size, offset_lower, domain_size, vel_ratio, out_pos, out_vel, stride_elems, 0, count, st, cells);
  if (dtype == GX_F32 && out_dtype == GX_F64)
    return emit_t<float, double>(psi, n, axis, cell_size, offset_lower, domain_size, vel_ratio, out_pos, out_vel, stride_elems, 0, count, st, cells);
  return fail("bad dtype");
}
