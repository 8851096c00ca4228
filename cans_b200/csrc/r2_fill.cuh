// The fused source of the forward x transform (r2r2.cuh): fillps (+ updt_rhs_b) evaluated at load time.
// Kept in a header of its own, free of kernel code, so that the CPU suite can compile exactly this arithmetic with g++
// (tests/emu/emu_aux.cpp, op "fill_source") and hold it to the oracle.
#pragma once
#include <cuda_runtime.h>

namespace cb {

// The right-hand side of the pressure Poisson equation is the scaled divergence of the prediction velocity
// (/root/reference/src/fillps.f90:38-50) plus the wall terms of updt_rhs_b (src/bound.f90:514-598), and the first thing
// `solver` does with it is the forward x transform.  With this source the transform's loads evaluate
//     p(i,j,k) = (w(i,j,k)-w(i,j,k-1))*dti*dzfi(k) + (v(i,j,k)-v(i,j-1,k))*dtidyi + (u(i,j,k)-u(i-1,j,k))*dtidxi
// (same expression, same operation order as the stand-alone fillps kernel) straight from u, v, w: p is never written and
// re-read (-16 B/point of HBM traffic on the two steps).  u, v, w are haloed like p and point at element (1,1,1).
template <class T> struct R2Fill {
  const T *u, *v, *w;
  const T* dzfi;        // dzfi(0:n3+1) of the local slab
  T dti, dtidxi, dtidyi;
  long long sj, sk;     // strides of j and k in the haloed arrays
  int k0;               // 0-based interior plane of group 0 of this launch (launches on z chunks)
  int any_rhsb;         // updt_rhs_b: any wall term at all?
  int idx[3][2];        // 1-based interior index of the plane that takes the wall term of (direction, side), 0 = none
  T val[3][2];          // rhsb * norm
};
// one line (j, k) of the fused source
template <class T> struct R2FillLine {
  const T *u, *v, *vm, *w, *wm;   // x = 0 of u(:,j,k), v(:,j,k), v(:,j-1,k), w(:,j,k), w(:,j,k-1)
  T dz;                           // dzfi(k)
  bool y0, y1, z0, z1;            // the line lies in the plane that takes the wall term of (y | z, lower | upper)
  __device__ __forceinline__ R2FillLine(const R2Fill<T>& F, long long off, int g, int j) {
    u = F.u + off; v = F.v + off; w = F.w + off;
    vm = v - F.sj; wm = w - F.sk;
    const int k = F.k0 + g;       // 0-based interior plane
    dz = __ldg(F.dzfi + k + 1);
    y0 = j + 1 == F.idx[1][0]; y1 = j + 1 == F.idx[1][1];
    z0 = k + 1 == F.idx[2][0]; z1 = k + 1 == F.idx[2][1];
  }
  // sample i (0-based) of the line
  __device__ __forceinline__ T at(const R2Fill<T>& F, int i) const {
    T r = (w[i] - wm[i]) * F.dti * dz + (v[i] - vm[i]) * F.dtidyi + (u[i] - u[i - 1]) * F.dtidxi;
    if (F.any_rhsb) {   // uniform branch; the six terms are added one after the other as the reference's six loops do
      if (i + 1 == F.idx[0][0]) r += F.val[0][0];
      if (i + 1 == F.idx[0][1]) r += F.val[0][1];
      if (y0) r += F.val[1][0];
      if (y1) r += F.val[1][1];
      if (z0) r += F.val[2][0];
      if (z1) r += F.val[2][1];
    }
    return r;
  }
};

}  // namespace cb
