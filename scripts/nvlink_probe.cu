// NVLink store-pattern probe (2 GPUs, one process, peer access): which write pattern of a producer kernel
// reaches which fraction of the peer-copy bandwidth?  Decides the layout of the exchange buffers of
// solve_dist (capi.cu): 128-byte rows 8 KB apart (what the y / Thomas kernels did in round 1) against
// dense blocks written with vector stores or TMA bulk stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nvlink_probe scripts/nvlink_probe.cu && ./nvlink_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// 16 bytes per thread, fully contiguous
__global__ void k_contig16(double2* dst, const double2* src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// 32 bytes per thread (st.global.v4.f64), contiguous
__global__ void k_contig32(double4* dst, const double4* src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double4 v = src[i];
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
  }
}
// rows of ROWB bytes whose consecutive rows (as a CTA walks them) are `stride` bytes apart at the destination;
// source read contiguously.  8 lanes x 16 B = 128-byte row pieces, like the y kernels.
__global__ void k_rows(char* dst, const char* src, size_t nrows, int rowb, size_t stride, size_t wrap) {
  const int ppr = rowb / 16;   // 16-byte pieces per row
  const size_t total = nrows * ppr;
  for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
    const size_t row = q / ppr, pc = q - row * ppr;
    // `wrap` consecutive rows land `stride` apart, then the next column block of the same region (stride * wrap bytes)
    const size_t blk = row / wrap, cpb = stride / rowb;
    const size_t d = (blk / cpb) * (stride * wrap) + (blk % cpb) * (size_t)rowb + (row % wrap) * stride + pc * 16;
    *reinterpret_cast<double2*>(dst + d) = *reinterpret_cast<const double2*>(src + q * 16);
  }
}
// TMA bulk store of `chunk` contiguous bytes per operation from shared memory
__global__ void k_bulk(char* dst, const char* src, size_t nchunks, int chunk) {
  extern __shared__ __align__(128) char sm[];
  for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    for (int o = threadIdx.x * 16; o < chunk; o += blockDim.x * 16)
      *reinterpret_cast<double2*>(sm + o) = *reinterpret_cast<const double2*>(src + ch * chunk + o);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + ch * chunk),
                   "r"((unsigned)__cvta_generic_to_shared(sm)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncthreads();
  }
}
// TMA bulk stores of 128-byte rows `stride` apart (one elected thread issues `rows` operations per tile)
__global__ void k_bulk_rows(char* dst, const char* src, size_t ntiles, int rows, size_t stride) {
  extern __shared__ __align__(128) char sm[];
  const int chunk = rows * 128;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    for (int o = threadIdx.x * 16; o < chunk; o += blockDim.x * 16)
      *reinterpret_cast<double2*>(sm + o) = *reinterpret_cast<const double2*>(src + t * chunk + o);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
      for (int r = threadIdx.x; r < rows; r += 32)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(dst + (t / (stride / 128)) * (stride * rows) + (t % (stride / 128)) * 128 + r * stride),
                     "r"((unsigned)__cvta_generic_to_shared(sm + r * 128)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncthreads();
  }
}
// remote LOADS: pull contiguous data from the peer
__global__ void k_pull16(double2* dst_local, const double2* src_remote, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst_local[i] = src_remote[i];
}

int main(int argc, char** argv) {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  const size_t bytes = (size_t)256 << 20;
  char *loc[2], *rem[2];   // loc[d]: source on device d; rem[d]: destination on device d (written by the other one)
  cudaStream_t st[2];
  cudaEvent_t e0[2], e1[2];
  for (int d = 0; d < 2; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    CK(cudaMalloc(&loc[d], bytes));
    CK(cudaMalloc(&rem[d], bytes + (64 << 20)));
    CK(cudaMemset(loc[d], 1, bytes));
    CK(cudaStreamCreate(&st[d]));
    CK(cudaEventCreate(&e0[d]));
    CK(cudaEventCreate(&e1[d]));
  }
  auto run = [&](const char* name, int both, auto launch) {
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      for (int d = 0; d <= both; ++d) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d])); launch(d); CK(cudaEventRecord(e1[d], st[d])); }
      float worst = 0.f;
      for (int d = 0; d <= both; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d])); if (ms > worst) worst = ms; }
      if (it >= 2 && worst < best) best = worst;
    }
    printf("%-44s %s  %7.3f ms  %7.1f GB/s per direction\n", name, both ? "bidir" : "unidir", best, bytes / best / 1e6);
    fflush(stdout);
  };
  const int G = 148 * 8;
  for (int both = 0; both <= 1; ++both) {
    run("cudaMemcpyPeerAsync", both, [&](int d) { CK(cudaMemcpyPeerAsync(rem[1 - d], 1 - d, loc[d], d, bytes, st[d])); });
    run("contiguous 16 B/thread", both, [&](int d) { k_contig16<<<G, 256, 0, st[d]>>>((double2*)rem[1 - d], (const double2*)loc[d], bytes / 16); });
    run("contiguous 32 B/thread (st.v4.f64)", both, [&](int d) { k_contig32<<<G, 256, 0, st[d]>>>((double4*)rem[1 - d], (const double4*)loc[d], bytes / 32); });
    run("128 B rows, 8 KB apart", both, [&](int d) { k_rows<<<G, 256, 0, st[d]>>>(rem[1 - d], loc[d], bytes / 128, 128, 8192, 64); });
    run("256 B rows, 8 KB apart", both, [&](int d) { k_rows<<<G, 256, 0, st[d]>>>(rem[1 - d], loc[d], bytes / 256, 256, 8192, 32); });
    run("512 B rows, 8 KB apart", both, [&](int d) { k_rows<<<G, 256, 0, st[d]>>>(rem[1 - d], loc[d], bytes / 512, 512, 8192, 16); });
    run("128 B rows, 2 MB apart", both, [&](int d) { k_rows<<<G, 256, 0, st[d]>>>(rem[1 - d], loc[d], bytes / 128, 128, 2 << 20, 128); });
    run("TMA bulk store, 8 KB chunks", both, [&](int d) { k_bulk<<<G, 256, 8192, st[d]>>>(rem[1 - d], loc[d], bytes / 8192, 8192); });
    run("TMA bulk store, 2 KB chunks", both, [&](int d) { k_bulk<<<G, 256, 2048, st[d]>>>(rem[1 - d], loc[d], bytes / 2048, 2048); });
    run("TMA bulk store, 128 B rows 8 KB apart", both, [&](int d) { k_bulk_rows<<<G, 256, 8192, st[d]>>>(rem[1 - d], loc[d], bytes / 8192, 64, 8192); });
    run("remote loads (pull) 16 B/thread", both, [&](int d) { k_pull16<<<G, 256, 0, st[d]>>>((double2*)loc[d], (const double2*)rem[1 - d], bytes / 16); });
    // copy engines: strided 2-D copies (one x window of the way-back exchange: 2 KB rows, 8 KB pitch) and chunked copies
    run("cudaMemcpy2DAsync 2 KB rows, pitch 8 KB", both, [&](int d) {
      for (int w = 0; w < 4; ++w) CK(cudaMemcpy2DAsync(rem[1 - d] + w * 2048, 8192, loc[d] + w * 2048, 8192, 2048, bytes / 8192, cudaMemcpyDeviceToDevice, st[d]));
    });
    run("cudaMemcpy2DAsync 512 B rows, pitch 8 KB", both, [&](int d) {
      for (int w = 0; w < 16; ++w) CK(cudaMemcpy2DAsync(rem[1 - d] + w * 512, 8192, loc[d] + w * 512, 8192, 512, bytes / 8192, cudaMemcpyDeviceToDevice, st[d]));
    });
    run("cudaMemcpyAsync 32 chunks of 8 MB", both, [&](int d) {
      for (int c = 0; c < 32; ++c) CK(cudaMemcpyAsync(rem[1 - d] + (size_t)c * (8 << 20), loc[d] + (size_t)c * (8 << 20), 8 << 20, cudaMemcpyDeviceToDevice, st[d]));
    });
    run("cudaMemcpyAsync 128 chunks of 2 MB", both, [&](int d) {
      for (int c = 0; c < 128; ++c) CK(cudaMemcpyAsync(rem[1 - d] + (size_t)c * (2 << 20), loc[d] + (size_t)c * (2 << 20), 2 << 20, cudaMemcpyDeviceToDevice, st[d]));
    });
  }
  // how many CTAs (of 256 threads) does a store kernel need to fill the link?  (can a producer kernel leave SMs to others?)
  for (int ctas : {18, 37, 74, 148, 296, 592}) {
    char name[64];
    snprintf(name, sizeof(name), "128 B rows 8 KB apart, %d CTAs", ctas);
    run(name, 1, [&](int d) { k_rows<<<ctas, 256, 0, st[d]>>>(rem[1 - d], loc[d], bytes / 128, 128, 8192, 64); });
  }
  // a copy-engine transfer next to an HBM-bound kernel on the same GPU: does either slow the other down?
  {
    char *big0, *big1;
    const size_t nb = (size_t)1 << 30;
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&big0, nb));
    CK(cudaMalloc(&big1, nb));
    cudaStream_t s2;
    CK(cudaStreamCreate(&s2));
    cudaEvent_t a0, a1, b0, b1;
    CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&a1)); CK(cudaEventCreate(&b0)); CK(cudaEventCreate(&b1));
    for (int mode = 0; mode < 3; ++mode) {   // 0: kernel alone, 1: copies alone (both directions), 2: together
      float tk = 0, tc = 0;
      for (int it = 0; it < 4; ++it) {
        CK(cudaDeviceSynchronize());
        CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize()); CK(cudaSetDevice(0));
        if (mode != 1) { CK(cudaEventRecord(a0, s2)); for (int r = 0; r < 2; ++r) k_contig16<<<G, 256, 0, s2>>>((double2*)big1, (const double2*)big0, nb / 16); CK(cudaEventRecord(a1, s2)); }
        if (mode != 0) {
          CK(cudaEventRecord(b0, st[0]));
          CK(cudaMemcpyAsync(rem[1], loc[0], bytes, cudaMemcpyDeviceToDevice, st[0]));
          CK(cudaEventRecord(b1, st[0]));
          CK(cudaSetDevice(1)); CK(cudaMemcpyAsync(rem[0], loc[1], bytes, cudaMemcpyDeviceToDevice, st[1])); CK(cudaSetDevice(0));
        }
        CK(cudaDeviceSynchronize());
        CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize()); CK(cudaSetDevice(0));
        if (mode != 1) CK(cudaEventElapsedTime(&tk, a0, a1));
        if (mode != 0) CK(cudaEventElapsedTime(&tc, b0, b1));
      }
      printf("overlap mode %d: local copy kernel (2 x 2 GiB traffic) %.3f ms = %.0f GB/s ; peer DMA (256 MiB out, 256 MiB in) %.3f ms = %.0f GB/s\n",
             mode, tk, tk > 0 ? 4.0 * nb / tk / 1e6 : 0.0, tc, tc > 0 ? bytes / tc / 1e6 : 0.0);
    }
  }
  return 0;
}
