// Whole unigrid path behind one C-ABI call: ICGenerator::done -> write (ic.hpp:1633-1648, 1275-1331)
//   draw white noise -> [fix_power] -> delta_k = sqrt(C(k)) * noise -> psi_j = c2r(i k_j delta_k / k^2), j = x,y,z
//   -> particles (pos = wrap(centroid + psi), vel = psi * ratio), streamed in x-slabs through a bounded staging
//   buffer, optionally copied to HOST arrays (component-major: out[axis][particle id]).
// The reference pulls particles one at a time through virtual evaluators into an mmapped file
// (mapperiterator.hpp:151-156, tipsy.hpp:120-146); here a slab of a component is emitted per kernel and the
// device->host copy of slab s overlaps the emit of slab s+1.
#include "../../include/genetic_b200.h"
#include "gx_common.cuh"
#include "zeldovich.cuh"
#include <algorithm>
#include <cstdlib>

extern "C" int gx_unigrid_ic(const gx_ic_params *p, const gx_spectrum *spectrum, void *delta_field, void *psi_field,
                             void *stage, size_t stage_bytes, void *host_pos, void *host_vel,
                             size_t host_capacity_particles, void *stream) {
  using namespace gx;
  if (!p || !spectrum || !delta_field || !psi_field || !stage) return fail("gx_unigrid_ic: null argument");
  const int n = p->n, dt = p->dtype, odt = p->out_dtype;
  cudaStream_t st = (cudaStream_t) stream;
  int rc;
  unsigned long long off = 0;
  static int fuse_transfer = -1;
  if (fuse_transfer < 0) { const char *e = getenv("GX_FUSE_TRANSFER"); fuse_transfer = (e && e[0] == '0') ? 0 : 1; }
  if (fuse_transfer && !p->fix_power && p->noise_mode != GX_NOISE_REAL) {
    // delta_k / k^2 = noise * sqrt(C(k)) / k^2 formed while the noise is delivered: the factor comes from the |k|^2 table
    // (L2 resident), the product is the one Field::applyTransferFunction would form, and the transfer sweep (2 B) is gone
    const void *table = nullptr;
    if ((rc = build_transfer_table(n, dt, p->grid_length, spectrum, 0.5, 1, st, &table))) return rc;
    // (the work field is idle until the first transform: it holds the staged draws of the noise generator)
    if ((rc = noise_draw_scaled(delta_field, n, dt, p->seed, p->noise_mode, p->reverse_order, &off, st, table, psi_field,
                                gx_field_bytes(n, dt))))
      return rc;
  } else {
    if ((rc = gx_noise_draw(delta_field, n, dt, p->seed, p->noise_mode, p->reverse_order, &off, stream))) return rc;
    if (p->fix_power && (rc = gx_unit_variance(delta_field, n, dt, stream))) return rc;
    // delta_k = sqrt(C(k)) * noise, left divided by k^2: the quotient all three displacement components start from
    if ((rc = apply_transfer_ex(delta_field, n, dt, p->grid_length, spectrum, 0.5, 1, st))) return rc;
  }

  const size_t osz = odt == GX_F64 ? 8 : 4;
  const long long n3 = (long long) n * n * n, n2 = (long long) n * n;
  // staging: two halves, each holding pos+vel of `slab` particles (whole x-planes)
  long long planes = (long long) (stage_bytes / 2 / (2 * osz)) / n2;
  if (planes < 1) return fail("gx_unigrid_ic: staging buffer smaller than one x-plane of pos+vel");
  if (planes > n) planes = n;
  const long long slab = planes * n2;
  const bool to_host = host_pos != nullptr || host_vel != nullptr;
  if (to_host && host_capacity_particles == 0) return fail("gx_unigrid_ic: host_capacity_particles == 0");
  static cudaStream_t copy_stream = nullptr;
  static cudaEvent_t ev_emit[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
  static int copy_device = -1;
  int cur_device = 0;
  GX_CUDA_CHECK(cudaGetDevice(&cur_device));
  if (to_host && (!copy_stream || copy_device != cur_device)) {
    copy_device = cur_device;
    GX_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      GX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_emit[i], cudaEventDisableTiming));
      GX_CUDA_CHECK(cudaEventCreateWithFlags(&ev_copy[i], cudaEventDisableTiming));
    }
  }
  bool copy_pending[2] = {false, false};
  int buf = 0;
  // z-pass (c2r) of one component, slab of x-planes by slab, with the particle evaluation fused into its store
  auto emit_component = [&](int axis, const void *field) -> int {
    EmitGeo eg;
    eg.cell_size = p->cell_size; eg.offset_lower = p->offset_lower[axis]; eg.domain_size = p->domain_size;
    eg.vel_ratio = p->vel_ratio;
    for (long long id0 = 0; id0 < n3; id0 += slab) {
      const long long cnt = std::min(slab, n3 - id0);
      char *half = (char *) stage + (size_t) buf * (stage_bytes / 2);
      void *dpos = half, *dvel = half + (size_t) slab * osz;
      if (to_host && copy_pending[buf]) GX_CUDA_CHECK(cudaStreamWaitEvent(st, ev_copy[buf], 0));
      if (int rc2 = zeldovich_z_emit(field, n, dt, axis, eg, (long) (id0 / n), (long) (cnt / n), dpos, dvel, odt, st))
        return rc2;
      if (to_host) {
        GX_CUDA_CHECK(cudaEventRecord(ev_emit[buf], st));
        GX_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, ev_emit[buf], 0));
        // component-major host arrays; a capacity below n^3 makes the host side a ring (the consumer is a writer thread)
        const size_t cap = host_capacity_particles;
        const size_t hoff = (size_t) (((unsigned long long) axis * (unsigned long long) n3 + (unsigned long long) id0) % (3ull * cap));
        const size_t room = 3 * cap - hoff;
        const size_t c1 = std::min<size_t>((size_t) cnt, room);
        if (host_pos) {
          GX_CUDA_CHECK(cudaMemcpyAsync((char *) host_pos + hoff * osz, dpos, c1 * osz, cudaMemcpyDeviceToHost, copy_stream));
          if (c1 < (size_t) cnt)
            GX_CUDA_CHECK(cudaMemcpyAsync(host_pos, (char *) dpos + c1 * osz, ((size_t) cnt - c1) * osz, cudaMemcpyDeviceToHost, copy_stream));
        }
        if (host_vel) {
          GX_CUDA_CHECK(cudaMemcpyAsync((char *) host_vel + hoff * osz, dvel, c1 * osz, cudaMemcpyDeviceToHost, copy_stream));
          if (c1 < (size_t) cnt)
            GX_CUDA_CHECK(cudaMemcpyAsync(host_vel, (char *) dvel + c1 * osz, ((size_t) cnt - c1) * osz, cudaMemcpyDeviceToHost, copy_stream));
        }
        GX_CUDA_CHECK(cudaEventRecord(ev_copy[buf], copy_stream));
        copy_pending[buf] = true;
      }
      buf ^= 1;
    }
    return 0;
  };
  static int shared = -1;
  if (shared < 0) { const char *e = getenv("GX_ZELDOVICH_SHARED"); shared = (e && e[0] == '0') ? 0 : 1; }
  if (shared) {
    // 5 column passes instead of 6: i k_y and i k_z do not depend on kx, so psi_y and psi_z share one x-transform of
    // G = delta_k / k^2. psi_x goes first (it needs G itself); then G <- x-pass(mirror(G)) in place; psi_y = y-pass(i k_y G)
    // into the work field; psi_z = y-pass(G) in place with i k_z applied as it is stored. (The first attempt shared the
    // y-transform instead and needed an out-of-place x-pass: 105 ms against 38 in place, profiles/r1s14_*.)
    if ((rc = zeldovich_xy(delta_field, psi_field, n, dt, p->grid_length, 0, 1, st))) return rc;
    if ((rc = emit_component(0, psi_field))) return rc;
    if ((rc = zeldovich_shared_x(delta_field, n, dt, st))) return rc;
    if ((rc = zeldovich_shared_y(delta_field, psi_field, n, dt, p->grid_length, 1, st))) return rc;
    if ((rc = emit_component(1, psi_field))) return rc;
    if ((rc = zeldovich_shared_y(delta_field, psi_field, n, dt, p->grid_length, 2, st))) return rc;
    if ((rc = emit_component(2, delta_field))) return rc;
  } else {
    for (int axis = 0; axis < 3; ++axis) {
      // x- and y-pass with the ik/k^2 operator and the Hermitian mirror fused into the first load
      if ((rc = zeldovich_xy(delta_field, psi_field, n, dt, p->grid_length, axis, 1, st))) return rc;
      if ((rc = emit_component(axis, psi_field))) return rc;
    }
  }
  if (to_host) {
    GX_CUDA_CHECK(cudaStreamSynchronize(copy_stream));
  }
  return 0;
}
