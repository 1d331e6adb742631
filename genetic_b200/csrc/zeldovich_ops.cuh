// Diagonal operators of the shared-transform Zeldovich schedule, used by the single-GPU path (zeldovich.cu) and by the
// slab path (slab.cu): psi_y = c2r(i k_y G), psi_z = c2r(i k_z G) with G = delta_k / k^2 (zeldovich.hpp:152-181); neither
// factor depends on kx, so one x-transform of G serves both components.
#pragma once
#include "spectral.cuh"

namespace gx {

template <typename T> struct OpIKy {  // (value, ix, iy, iz): times i k_y
  int n;
  T kmin;
  __device__ __forceinline__ cplx<T> operator()(cplx<T> v, int, int iy, int) const {
    if (iy == (n >> 1)) return mk<T>(0, 0);  // zeldovich.hpp:170-176: zero on the Nyquist plane of the component's axis
    const T ka = mul_rn((T) signed_mode(iy, n), kmin);
    return mk<T>(mul_rn(-v.y, ka), mul_rn(v.x, ka));
  }
};
template <typename T> struct OpIKz {  // times i k_z
  int n;
  T kmin;
  __device__ __forceinline__ cplx<T> operator()(cplx<T> v, int, int, int iz) const {
    if (iz == (n >> 1)) return mk<T>(0, 0);
    const T ka = mul_rn((T) iz, kmin);
    return mk<T>(mul_rn(-v.y, ka), mul_rn(v.x, ka));
  }
};

}  // namespace gx
