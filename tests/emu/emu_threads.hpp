// Host-thread execution model shared by the kernel emulators (TEST INFRASTRUCTURE): every CUDA thread of a CTA is a host
// thread, __syncthreads / __syncwarp / bar.sync are real barriers between those threads (a thread that returns early drops
// out of them, as on the GPU), warp shuffles go through one slot per lane between two warp barriers, atomics take the
// emulator's lock, shared memory is one buffer per CTA, CTAs run one after the other.
// Include it BEFORE the kernel headers.
#pragma once
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

// ---- the execution context of the emulated thread -------------------------------------------------------------------
static thread_local uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
namespace cb { unsigned char cb_smem_raw[232 * 1024] __attribute__((aligned(128))); }   // `extern __shared__` of the kernels

struct Bar {
  int expected = 0, arrived = 0;
  unsigned gen = 0;
  std::condition_variable cv;   // one per barrier: a release wakes the threads of this group only
  void reset(int exp) { expected = exp; arrived = 0; gen = 0; }
};
static std::mutex g_m;
static Bar g_cta, g_warp[32], g_named[16];
static void bar_wait(Bar& b, int expected) {   // expected <= 0: every live thread of the group (read under the lock)
  std::unique_lock<std::mutex> lk(g_m);
  if (expected <= 0) expected = b.expected;
  const unsigned gen = b.gen;
  if (++b.arrived >= expected) { b.arrived = 0; ++b.gen; b.cv.notify_all(); }
  else b.cv.wait(lk, [&] { return b.gen != gen; });
}
static void bar_drop(Bar& b) {   // a thread of the group has returned from the kernel
  --b.expected;
  if (b.arrived > 0 && b.arrived >= b.expected) { b.arrived = 0; ++b.gen; b.cv.notify_all(); }
}
static inline void __syncthreads() { bar_wait(g_cta, 0); }
static inline void __syncwarp(unsigned = 0xffffffffu) { bar_wait(g_warp[threadIdx.x / 32], 0); }
static inline void cb_emu_named_bar(int id, int cnt) { bar_wait(g_named[id], cnt); }
// warp shuffles: every lane publishes its value, the warp meets, every lane reads its partner's
static unsigned long long g_shfl[32][32];
template <class T> static T __shfl_xor_sync(unsigned, T v, int o) {
  static_assert(sizeof(T) <= 8, "shuffle slot");
  const int w = threadIdx.x / 32, l = threadIdx.x % 32;
  memcpy(&g_shfl[w][l], &v, sizeof(T));
  bar_wait(g_warp[w], 0);
  T r;
  memcpy(&r, &g_shfl[w][l ^ o], sizeof(T));
  bar_wait(g_warp[w], 0);
  return r;
}
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { std::lock_guard<std::mutex> lk(g_m); const unsigned long long o = *a; *a += v; return o; }
static inline double atomicAdd(double* a, double v) { std::lock_guard<std::mutex> lk(g_m); const double o = *a; *a += v; return o; }
static inline unsigned long long atomicCAS(unsigned long long* a, unsigned long long c, unsigned long long v) {
  std::lock_guard<std::mutex> lk(g_m);
  const unsigned long long o = *a;
  if (o == c) *a = v;
  return o;
}
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline unsigned __float_as_uint(float d) { unsigned r; memcpy(&r, &d, 4); return r; }
template <class T> static T __ldg(const T* p) { return *p; }
template <class T, class V> static void __stcs(T* p, V v) { *p = v; }
#define __launch_bounds__(...)


// run `kernel(args...)` for every thread of every CTA of a 1-D launch
template <class K, class... ARGS> static void launch(unsigned grid, unsigned block, K kernel, const ARGS&... A) {
  gridDim = dim3(grid); blockDim = dim3(block);
  for (unsigned bx = 0; bx < grid; ++bx) {
    g_cta.reset((int)block);
    for (unsigned w = 0; w < 32; ++w) { const int lo = (int)w * 32; g_warp[w].reset((int)block > lo ? std::min(32, (int)block - lo) : 0); }
    for (auto& b : g_named) b.reset(0);
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned tx = 0; tx < block; ++tx)
      th.emplace_back([&, tx] {
        blockIdx = uint3{bx, 0, 0};
        threadIdx = uint3{tx, 0, 0};
        kernel(A...);
        std::lock_guard<std::mutex> lk(g_m);
        bar_drop(g_cta);
        bar_drop(g_warp[tx / 32]);
      });
    for (auto& t : th) t.join();
  }
}

template <class T> static std::vector<T> rd(const std::string& dir, const char* name, size_t n) {
  std::vector<T> v(n);
  FILE* f = fopen((dir + "/" + name + ".bin").c_str(), "rb");
  if (!f || fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "emulator: cannot read %s\n", name); exit(2); }
  fclose(f);
  return v;
}
template <class T> static void wr(const std::string& dir, const std::string& name, const std::vector<T>& v) {
  FILE* f = fopen((dir + "/" + name + ".bin").c_str(), "wb");
  if (!f || fwrite(v.data(), sizeof(T), v.size(), f) != v.size()) { fprintf(stderr, "emulator: cannot write %s\n", name.c_str()); exit(2); }
  fclose(f);
}
