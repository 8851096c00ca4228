// Host-side planning for the r2r engine: radix schedule, twiddle tables and
// the digit-reversal table.  Plays the role of fftw_plan_guru_r2r in
// /root/reference/src/fft.f90:96-97,161-162 (plan creation) -- pure C++, no
// CUDA, shared by the library and the CPU emulator.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include "fft_engine.cuh"

namespace cb {

struct HostFftPlan {
  int n = 0, M = 0, kind = 0;
  bool fast = false;  // false -> direct O(n^2) kernel
  std::vector<int> radix;
  std::vector<long double> tw_re, tw_im;    // M
  std::vector<long double> twp_re, twp_im;  // M/2+1
  std::vector<long double> mak_re, mak_im;  // M+1
  std::vector<uint16_t> rev;                // M
};

inline bool plan_radices(int M, std::vector<int>& radix) {
  radix.clear();
  if (M < 1) return false;
  int e = 0, m = M;
  while (m % 2 == 0) { m /= 2; ++e; }
  std::vector<int> odd;
  const int primes[] = {3, 5, 7, 11, 13};
  for (int p : primes)
    while (m % p == 0) { m /= p; odd.push_back(p); }
  if (m != 1) return false;
  if (e > 0) {
    int ns = (e + 3) / 4;
    int base = e / ns, extra = e % ns;
    for (int s = 0; s < ns; ++s) radix.push_back(1 << (base + (s < extra ? 1 : 0)));
  }
  for (int p : odd) radix.push_back(p);
  return (int)radix.size() <= CB_MAX_STAGES;
}

inline HostFftPlan make_host_plan(int n, int kind) {
  HostFftPlan P;
  P.n = n;
  P.kind = kind;
  P.M = n / 2;
  const long double pi = 3.14159265358979323846264338327950288L;
  P.fast = kind_is_fast(kind) && n >= 2 && (n % 2 == 0) && P.M <= 65536 && plan_radices(P.M, P.radix);
  if (!P.fast) return P;
  const int M = P.M;
  P.tw_re.resize(M); P.tw_im.resize(M);
  for (int t = 0; t < M; ++t) {
    long double a = -2.0L * pi * t / M;
    P.tw_re[t] = cosl(a); P.tw_im[t] = sinl(a);
  }
  P.twp_re.resize(M / 2 + 1); P.twp_im.resize(M / 2 + 1);
  for (int k = 0; k <= M / 2; ++k) {
    long double a = -2.0L * pi * k / n;
    P.twp_re[k] = cosl(a); P.twp_im[k] = sinl(a);
  }
  P.mak_re.resize(M + 1); P.mak_im.resize(M + 1);
  for (int k = 0; k <= M; ++k) {
    long double a = pi * k / (2.0L * n);
    P.mak_re[k] = cosl(a); P.mak_im[k] = sinl(a);
  }
  P.rev.resize(M);
  for (int k = 0; k < M; ++k) {
    int kk = k, pos = 0, blk = M;
    for (int r : P.radix) {
      blk /= r;
      pos += (kk % r) * blk;
      kk /= r;
    }
    P.rev[k] = (uint16_t)pos;
  }
  return P;
}

// tables for the direct O(n^2) kernel: (cos, sin)(pi m / Q), m = 0..2Q-1
inline int slow_Q(int n, int kind) {
  switch (kind) {
    case K_R2HC: case K_HC2R: return n;
    case K_REDFT00: return n > 1 ? n - 1 : 1;
    case K_RODFT00: return n + 1;
    case K_REDFT11: case K_RODFT11: return 4 * n;
    default: return 2 * n;
  }
}

}  // namespace cb
