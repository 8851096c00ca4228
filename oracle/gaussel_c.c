/* C restatement of `gaussel` (/root/reference/src/solver.f90:114-307), FP64.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/cans_oracle.py header).
 * Same loop structure as the reference: k outer, (j,i) inner with an OpenMP
 * parallel-for over the plane, work array d(nx,ny,nn) allocated per call,
 * identical expression order; built with -ffp-contract=off so that no FMA is
 * formed (gfortran -O3 on baseline x86-64 forms none either).
 * p is [k][j][i] contiguous, lambdaxy is [j][i]. */
#include <float.h>
#include <math.h>
#include <stdlib.h>

int gaussel_c(long nx, long ny, long n, const double* a, const double* b, const double* c, int is_periodic,
              double norm, double* p, const double* lam) {
  const long np = nx * ny;
  const long nn = is_periodic ? n - 1 : n;
  if (nn < 1) return -1;
  double* d = (double*)malloc(sizeof(double) * np * nn);
  if (!d) return -2;
  long k, e;
#pragma omp parallel for
  for (e = 0; e < np; ++e) {
    double z = 1.0 / (b[0] + lam[e]);
    d[e] = c[0] * z;
    p[e] = p[e] * norm * z;
  }
  for (k = 1; k < nn; ++k) {
    double* pk = p + k * np; const double* pm = p + (k - 1) * np;
    double* dk = d + k * np; const double* dm = d + (k - 1) * np;
    const double ak = a[k], bk = b[k], ck = c[k];
    const int last = (k == nn - 1);
#pragma omp parallel for
    for (e = 0; e < np; ++e) {
      const double bl = bk + lam[e];
      const double ad = ak * dm[e];
      const double den = bl - ad;
      const double tol = DBL_EPSILON * fmax(fabs(bl), fabs(ad));
      if (last && fabs(den) <= tol) { dk[e] = 0.0; pk[e] = 0.0; }
      else { const double z = 1.0 / den; dk[e] = ck * z; pk[e] = (pk[e] * norm - ak * pm[e]) * z; }
    }
  }
  for (k = nn - 2; k >= 0; --k) {
    double* pk = p + k * np; const double* pp = p + (k + 1) * np; const double* dk = d + k * np;
#pragma omp parallel for
    for (e = 0; e < np; ++e) pk[e] = pk[e] - dk[e] * pp[e];
  }
  if (is_periodic) {
    double* p2 = (double*)calloc(np * nn, sizeof(double));
    if (!p2) { free(d); return -2; }
#pragma omp parallel for
    for (e = 0; e < np; ++e) { p2[e] = -a[0]; p2[(nn - 1) * np + e] = p2[(nn - 1) * np + e] - c[nn - 1]; }
#pragma omp parallel for
    for (e = 0; e < np; ++e) { double z = 1.0 / (b[0] + lam[e]); d[e] = c[0] * z; p2[e] = p2[e] * z; }
    for (k = 1; k < nn; ++k) {
      double* qk = p2 + k * np; const double* qm = p2 + (k - 1) * np;
      double* dk = d + k * np; const double* dm = d + (k - 1) * np;
#pragma omp parallel for
      for (e = 0; e < np; ++e) {
        const double z = 1.0 / (b[k] + lam[e] - a[k] * dm[e]);
        dk[e] = c[k] * z; qk[e] = (qk[e] - a[k] * qm[e]) * z;
      }
    }
    for (k = nn - 2; k >= 0; --k) {
      double* qk = p2 + k * np; const double* qp = p2 + (k + 1) * np; const double* dk = d + k * np;
#pragma omp parallel for
      for (e = 0; e < np; ++e) qk[e] = qk[e] - dk[e] * qp[e];
    }
    double* pl = p + nn * np;
#pragma omp parallel for
    for (e = 0; e < np; ++e) {
      const double q1 = p2[e], qn = p2[(nn - 1) * np + e];
      const double den = b[nn] + lam[e] + c[nn] * q1 + a[nn] * qn;
      const double tol = DBL_EPSILON * fmax(fabs(b[nn] + lam[e]), fabs(c[nn] * q1 + a[nn] * qn));
      if (fabs(den) <= tol) pl[e] = 0.0;
      else pl[e] = (pl[e] * norm - c[nn] * p[e] - a[nn] * p[(nn - 1) * np + e]) / den;
    }
    for (k = 0; k < nn; ++k) {
      double* pk = p + k * np; const double* qk = p2 + k * np;
#pragma omp parallel for
      for (e = 0; e < np; ++e) pk[e] = pk[e] + qk[e] * pl[e];
    }
    free(p2);
  }
  free(d);
  return 0;
}
