// umma_probe: microbenchmarks and known-answer checks of tcgen05 behaviour the conv kernels depend on.
// Not part of the product library; built by `make -C tools` and run on the GPU box by hand:
//   tools/build/umma_probe rate          cycles per tcgen05.mma (1-CTA and 2-CTA, N = 64/128/256), operands
//                                        resident in shared memory, with and without concurrent bulk-copy fills
//   tools/build/umma_probe check         (a) cta_group::2 GEMM against a host reference,
//                                        (b) descriptors whose start address / SBO are not 1024-byte aligned
//   tools/build/umma_probe checknarrow   (d) K-major 32 / 64-byte swizzle with shifted starts, (e) MN-major 32 / 64-byte
//                                        swizzle with pixel-shifted M atoms and row-shifted N atoms (conv_halo.cu)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../unet-medical-image-contour-segmentation_b200/csrc/tc_common.cuh"

using namespace ub;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

// ---------------------------------------------------------------- cta_group::2 wrappers
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* g, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem)), "l"(g), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- rate
// One CTA (or CTA pair) per SM.  The issuing thread loops over `iters` steps of 8 MMAs (4 K-slices x 2
// accumulators, like one (tap, channel-chunk) step of tc2_fprop), rotating over 4 operand stages and
// committing each step to a ring of 4 mbarriers it waits on 4 steps later (a bounded pipeline).  A
// second warp optionally streams `fill_bytes` per step from global memory into scratch shared memory
// with bulk copies, paced by the same ring.
constexpr int kRing = 4;
template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int fill_bytes, const uint8_t* gsrc,
                                                      long long* out, int mn) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;                    // 4 x 16 KB
  uint8_t* b_ring = smem + 4 * 16384;        // 4 x 16 KB (32 KB for N = 256 at CG = 1 -> 2 stages alias; values unused)
  uint8_t* scratch = smem + 8 * 16384;           // 64 KB fill target
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + 65536);
  uint64_t* done = bars;                     // kRing
  uint64_t* fbar = bars + kRing;             // kRing
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRing);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < (8 * 16384 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kRing; ++i) { mbar_init(&done[i], 1); mbar_init(&fbar[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) { if (CG == 2) tmem_alloc2(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && rank == 0) {
    // mn = 1: both operands MN-major (the wgrad form: K = pixels), K step = 16 rows of 128 B
    const uint32_t idesc = make_idesc(false, mn != 0, mn != 0, 128 * CG, N);
    const uint64_t a_t = make_desc(smem_u32(a_ring), mn ? 8192 : 16, 1024), b_t = make_desc(smem_u32(b_ring), mn ? 8192 : 16, 1024);
    const uint32_t kstep = mn ? 128u : 2u;
    const uint32_t dcol2 = N <= 256 ? (uint32_t)N : 0u;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % kRing;
      if (it >= kRing) mbar_wait(&done[s], ((it / kRing) - 1) & 1);
      if (fill_bytes > 0) mbar_wait(&fbar[s], (it / kRing) & 1);      // like a consumer waiting for its stage
      tc_fence_after();
      const uint64_t a0 = a_t + s * (16384 >> 4), b0 = b_t + (s & (N > 128 && CG == 1 ? 1 : 3)) * ((N > 128 && CG == 1 ? 32768 : 16384) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (CG == 2) {
            umma2(tmem_base, a0 + kstep * (kk & (mn ? 1 : 3)), b0 + kstep * (kk & (mn ? 1 : 3)), idesc, 1);
            umma2(tmem_base + dcol2, a0 + 512 + kstep * (kk & (mn ? 1 : 3)), b0 + kstep * (kk & (mn ? 1 : 3)), idesc, 1);
          } else {
            umma<false>(tmem_base, a0 + kstep * (kk & (mn ? 1 : 3)), b0 + kstep * (kk & (mn ? 1 : 3)), idesc, 1);
            umma<false>(tmem_base + dcol2, a0 + 512 + kstep * (kk & (mn ? 1 : 3)), b0 + kstep * (kk & (mn ? 1 : 3)), idesc, 1);
          }
        }
        if (CG == 2) umma_commit2(&done[s], 3); else umma_commit(&done[s]);
      }
      __syncwarp();
    }
    for (int it = iters > kRing ? iters - kRing : 0; it < iters; ++it) mbar_wait(&done[it % kRing], (it / kRing) & 1);
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp == 0) {
    // follower CTA of a pair: its `done` barriers receive the multicast commits; just wait for the end
    for (int it = 0; it < iters; ++it) mbar_wait(&done[it % kRing], (it / kRing) & 1);
    if (lane == 0) out[blockIdx.x] = 0;
  } else if (warp == 2 && fill_bytes > 0) {
    const uint8_t* src = gsrc + (size_t)blockIdx.x * (1 << 20);
    for (int it = 0; it < iters; ++it) {
      const int s = it % kRing;
      if (it >= kRing) {
        mbar_wait(&done[s], ((it / kRing) - 1) & 1);      // MMAs of step it-4 are done
        mbar_wait(&fbar[s], ((it / kRing) - 1) & 1);      // and so is the fill that used this slot
      }
      if (elect_one()) {
        mbar_expect_tx(&fbar[s], fill_bytes);
        bulk_g2s(scratch + s * 16384, src + ((it * 16384) & ((1 << 20) - 1)), fill_bytes, &fbar[s]);
      }
      __syncwarp();
    }
    for (int it = iters > kRing ? iters - kRing : 0; it < iters; ++it) mbar_wait(&fbar[it % kRing], (it / kRing) & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) { if (CG == 2) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

static void run_rate() {
  const int smem = 8 * 16384 + 65536 + 1024 + 256;
  CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  uint8_t* gsrc;
  CK(cudaMalloc(&gsrc, 148u << 20));
  CK(cudaMemset(gsrc, 0, 148u << 20));
  long long* out;
  CK(cudaMalloc(&out, 148 * sizeof(long long)));
  const int iters = 4000;
  for (int mn = 0; mn <= 1; ++mn)
  for (int cg = 1; cg <= 2; ++cg)
    for (int N : {64, 128, 256})
      for (int fill : {0, 8192}) {
        if (mn && N == 256) continue;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        for (int rep = 0; rep < 2; ++rep) {
          if (cg == 1) CK(cudaLaunchKernelEx(&cfg, rate_kernel<1>, N, iters, fill, (const uint8_t*)gsrc, out, mn));
          else CK(cudaLaunchKernelEx(&cfg, rate_kernel<2>, N, iters, fill, (const uint8_t*)gsrc, out, mn));
          CK(cudaDeviceSynchronize());
        }
        std::vector<long long> h(148);
        CK(cudaMemcpy(h.data(), out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
        double sum = 0; int n = 0; long long mx = 0;
        for (auto v : h) if (v > 0) { sum += v; ++n; if (v > mx) mx = v; }
        const double per = sum / n / (iters * 8.0);
        const double floor = 128.0 * N / 256.0;            // cycles per 128 x N x 16 MMA per SM at 8192 FLOP/cycle/SM
        printf("%s cta_group::%d N=%3d fill=%5d B/step: %.1f cycles/MMA (max CTA %.1f), floor %.0f -> %.0f%% of tensor peak\n", mn ? "MN-major" : "K-major ", cg, N,
               fill, per, mx / (iters * 8.0), floor, 100.0 * floor / per);
      }
}


// ---------------------------------------------------------------- rate2: MN-major shapes of the N-stacked wgrad
// cta_group::1 (or ::2), both operands MN-major (K = pixels).  B = `N / 64` sub-tiles of 64 channels `lbo_b` bytes apart
// (128 B = the same halo box shifted by one pixel, i.e. overlapping reads; 8192 = separate tiles), 8-pixel groups
// `sbo_b` bytes apart.  M = 64 / 128 per CTA.  Reports cycles per MMA against the tensor floor M*N*16*2 / 8192.
template <int CG>
__global__ void __launch_bounds__(128, 1) rate2_kernel(int M, int N, int lbo_b, int sbo_b, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;                    // 4 x 16 KB
  uint8_t* b_ring = smem + 4 * 16384;        // 4 x 32 KB
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 4 * 16384 + 4 * 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + kRing);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kRing; ++i) mbar_init(&done[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) { if (CG == 2) tmem_alloc2(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc(false, true, true, M * CG, N);
    const uint64_t a_t = make_desc(smem_u32(a_ring), 8192, 1024), b_t = make_desc(smem_u32(b_ring), lbo_b, sbo_b);
    const uint32_t dcol2 = 2 * N <= 512 ? (uint32_t)N : 0u;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % kRing;
      if (it >= kRing) mbar_wait(&done[s], ((it / kRing) - 1) & 1);
      tc_fence_after();
      const uint64_t a0 = a_t + s * (16384 >> 4), b0 = b_t + s * (32768 >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t ka = (uint32_t)(kk * 2 * 1024) >> 4, kb = (uint32_t)(kk * 2 * sbo_b) >> 4;
          if (CG == 2) {
            umma2(tmem_base, a0 + ka, b0 + kb, idesc, 1);
            umma2(tmem_base + dcol2, a0 + 64 + ka, b0 + kb, idesc, 1);
          } else {
            umma<false>(tmem_base, a0 + ka, b0 + kb, idesc, 1);
            umma<false>(tmem_base + dcol2, a0 + 64 + ka, b0 + kb, idesc, 1);
          }
        }
        if (CG == 2) umma_commit2(&done[s], 3); else umma_commit(&done[s]);
      }
      __syncwarp();
    }
    for (int it = iters > kRing ? iters - kRing : 0; it < iters; ++it) mbar_wait(&done[it % kRing], (it / kRing) & 1);
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp == 0) {
    for (int it = 0; it < iters; ++it) mbar_wait(&done[it % kRing], (it / kRing) & 1);
    if (lane == 0) out[blockIdx.x] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) { if (CG == 2) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

static void run_rate2() {
  const int smem = 4 * 16384 + 4 * 32768 + 1024 + 256;
  CK(cudaFuncSetAttribute(rate2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(rate2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* out;
  CK(cudaMalloc(&out, 148 * sizeof(long long)));
  const int iters = 4000;
  struct Case { int cg, M, N, lbo, sbo; };
  const Case cases[] = {{1, 128, 64, 8192, 1024},  {1, 128, 128, 8192, 1024}, {1, 128, 192, 8192, 1024}, {1, 128, 256, 8192, 1024},
                        {1, 128, 128, 128, 1280},  {1, 128, 192, 128, 1280},  {1, 128, 256, 128, 1280},
                        {1, 64, 64, 8192, 1024},   {1, 64, 128, 128, 1280},   {1, 64, 192, 128, 1280},   {1, 64, 256, 128, 1280},
                        {2, 128, 128, 8192, 1024}, {2, 128, 192, 8192, 1024}, {2, 128, 256, 8192, 1024}, {2, 128, 256, 128, 1280},
                        {2, 64, 128, 8192, 1024},  {2, 64, 256, 8192, 1024}};
  for (const Case& c : cases) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c.cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
      if (c.cg == 1) CK(cudaLaunchKernelEx(&cfg, rate2_kernel<1>, c.M, c.N, c.lbo, c.sbo, iters, out));
      else CK(cudaLaunchKernelEx(&cfg, rate2_kernel<2>, c.M, c.N, c.lbo, c.sbo, iters, out));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("rate2 cg=%d M=%d N=%d: CUDA error %s\n", c.cg, c.M, c.N, cudaGetErrorString(e)); exit(1); }
    }
    std::vector<long long> h(148);
    CK(cudaMemcpy(h.data(), out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
    double sum = 0; int n = 0;
    for (auto v : h) if (v > 0) { sum += v; ++n; }
    const double per = sum / n / (iters * 8.0);
    const double floor = (double)c.M * c.N / 256.0;      // per SM: M x N x 16 x 2 FLOP at 8192 FLOP/cycle
    printf("rate2 MN-major cta_group::%d M/CTA=%3d N=%3d lboB=%4d sboB=%4d: %.1f cycles/MMA, floor %.0f -> %.0f%% of tensor peak\n",
           c.cg, c.M, c.N, c.lbo, c.sbo, per, floor, 100.0 * floor / per);
  }
}

// ---------------------------------------------------------------- check (a): cta_group::2 GEMM
// D[256][N] = A[256][64] * B[N][64]^T, bf16 in, fp32 out.  CTA r of the pair holds A rows [128r, 128r+128) and
// B rows [N/2 r, N/2 r + N/2), both K-major SWIZZLE_128B at the same shared-memory offsets.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_gemm_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int N) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;            // 128 rows x 128 B
  uint8_t* sb = smem + 16384;    // N/2 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 16384);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // swizzled store: row r, 16-byte chunk c -> r * 128 + ((c ^ (r & 7)) << 4)
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sa + r * 128 + ((c ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(A + (size_t)(rank * 128 + r) * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < (N / 2) * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sb + r * 128 + ((c ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(B + (size_t)(rank * (N / 2) + r) * 64 + c * 8);
  }
  if (warp == 0 && lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc2(tmem_slot, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc(false, false, false, 256, N);
    const uint64_t a0 = make_desc(smem_u32(sa), 16, 1024), b0 = make_desc(smem_u32(sb), 16, 1024);
    if (elect_one()) {
      for (int kk = 0; kk < 4; ++kk) umma2(tmem_base, a0 + 2 * kk, b0 + 2 * kk, idesc, kk > 0);
      umma_commit2(bar, 3);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int ch = 0; ch < N / 32; ++ch) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
    for (int e = 0; e < 32; ++e) D[(size_t)(rank * 128 + warp * 32 + lane) * N + ch * 32 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem_base, 256);
}

// ---------------------------------------------------------------- check (b): unaligned descriptor starts
// Shared memory holds `rows` pixel rows of 128 B written with the swizzle TMA applies (16-byte chunk index XOR
// bits [7,10) of the absolute shared address).  A = 128 rows addressed as start + (m / 8) * sbo + (m % 8) * 128
// with start = base + start_row * 128; B = 64 x 64 identity-like known matrix; D is compared on the host.
__global__ void __launch_bounds__(128, 1) shift_kernel(const __nv_bfloat16* X, const __nv_bfloat16* B, float* D, int rows,
                                                        int start_row, int sbo, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sx = smem;                 // rows x 128 B (up to 48 KB)
  uint8_t* sb = smem + 49152;         // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const uint32_t addr = smem_u32(sx) + r * 128;
    *reinterpret_cast<uint4*>(sx + r * 128 + ((c ^ ((addr >> 7) & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(X + (size_t)r * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sb + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + (size_t)r * 64 + c * 8);
  }
  if (warp == 0 && lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(false, false, false, 128, 64);
    const uint32_t start = smem_u32(sx) + start_row * 128;
    uint64_t a0 = make_desc(start, 16, sbo);
    if (use_base_offset) a0 |= (uint64_t)((start >> 7) & 7) << 49;
    const uint64_t b0 = make_desc(smem_u32(sb), 16, 1024);
    if (elect_one()) {
      for (int kk = 0; kk < 4; ++kk) umma<false>(tmem_base, a0 + 2 * kk, b0 + 2 * kk, idesc, kk > 0);
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
    for (int e = 0; e < 32; ++e) D[(size_t)(warp * 32 + lane) * 64 + ch * 32 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

// ---------------------------------------------------------------- check (c): MN-major operands (wgrad form)
// D[h*64 + c][n] = sum over an 8x8 pixel tile of X[(i + a) * box_w + j + dx0 + h][c] * Y[i * 8 + j][n]
// X = box of box_w-pixel rows (128 B per pixel, absolute-address swizzle), A is MN-major with its two 64-row halves
// `lbo` bytes apart (lbo = 128: the same box shifted by one pixel) and 8-pixel groups `box_w * 128` bytes apart.
__global__ void __launch_bounds__(128, 1) mn_kernel(const __nv_bfloat16* X, const __nv_bfloat16* Y, float* D, int xrows,
                                                     int box_w, int a, int dx0, int lbo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sx = smem;                 // xrows x 128 B (<= 48 KB)
  uint8_t* sy = smem + 49152;         // 64 pixels x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < xrows * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const uint32_t addr = smem_u32(sx) + r * 128;
    *reinterpret_cast<uint4*>(sx + r * 128 + ((c ^ ((addr >> 7) & 7)) << 4)) = *reinterpret_cast<const uint4*>(X + (size_t)r * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sy + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(Y + (size_t)r * 64 + c * 8);
  }
  if (warp == 0 && lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(false, true, true, 128, 64);
    if (elect_one()) {
      for (int k = 0; k < 4; ++k) {
        const uint32_t astart = smem_u32(sx) + ((2 * k + a) * box_w + dx0) * 128;
        const uint64_t da = make_desc(astart, lbo, box_w * 128, kLayoutSW128);
        const uint64_t db = make_desc(smem_u32(sy) + k * 2048, 8192, 1024, kLayoutSW128);
        umma<false>(tmem_base, da, db, idesc, k > 0);
      }
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
    for (int e = 0; e < 32; ++e) D[(size_t)(warp * 32 + lane) * 64 + ch * 32 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}


// ---------------------------------------------------------------- check (d): K-major 32 / 64-byte swizzle, shifted starts
// Pixel rows of P bytes (P / 2 bf16 channels) written with the P-byte swizzle as a function of the absolute
// shared-memory address (16-byte chunk index XOR address bits [7, 7 + log2(P / 16))).  A = 128 rows addressed as
// start + (m / 8) * sbo + (m % 8) * P, start = base + start_row * P (not atom aligned); B = 16 x (P / 2), K-major.
constexpr uint32_t kLayoutSW64 = 4, kLayoutSW32 = 6;
template <int P>
__global__ void __launch_bounds__(128, 1) shiftn_kernel(const __nv_bfloat16* X, const __nv_bfloat16* B, float* D, int rows,
                                                         int start_row, int sbo) {
  constexpr int CPR = P / 16, MASK = CPR - 1, CH = P / 2;
  constexpr uint32_t LAYOUT = P == 64 ? kLayoutSW64 : kLayoutSW32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sx = smem;
  uint8_t* sb = smem + 49152;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < rows * CPR; i += blockDim.x) {
    const int r = i / CPR, c = i % CPR;
    const uint32_t addr = smem_u32(sx) + r * P;
    *reinterpret_cast<uint4*>(sx + r * P + ((c ^ ((addr >> 7) & MASK)) << 4)) =
        *reinterpret_cast<const uint4*>(X + (size_t)r * CH + c * 8);
  }
  for (int i = threadIdx.x; i < 16 * CPR; i += blockDim.x) {
    const int r = i / CPR, c = i % CPR;
    const uint32_t addr = smem_u32(sb) + r * P;
    *reinterpret_cast<uint4*>(sb + r * P + ((c ^ ((addr >> 7) & MASK)) << 4)) = *reinterpret_cast<const uint4*>(B + (size_t)r * CH + c * 8);
  }
  if (warp == 0 && lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(tmem_slot, 32);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(false, false, false, 128, 16);
    const uint64_t a0 = make_desc(smem_u32(sx) + start_row * P, 16, sbo, LAYOUT);
    const uint64_t b0 = make_desc(smem_u32(sb), 16, 8 * P, LAYOUT);
    if (elect_one()) {
      for (int kk = 0; kk < P / 32; ++kk) umma<false>(tmem_base, a0 + 2 * kk, b0 + 2 * kk, idesc, kk > 0);
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
    for (int e = 0; e < 16; ++e) D[(size_t)(warp * 32 + lane) * 16 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 32);
}

// ---------------------------------------------------------------- check (e): MN-major 32 / 64-byte swizzle, stacked shifts
// wgrad form for narrow layers.  X: box of box_w-pixel rows, PX bytes per pixel; G: box of gbox_w-pixel rows, PG bytes
// per pixel.  A (MN-major): M = 128 = (256 / PX) atoms of PX / 2 channels, LBO = PX: atom h is the box shifted by h
// pixels; B (MN-major): N = 3 atoms of PG / 2 channels, LBO = gbox_w * PG: atom s is the box shifted by s rows.
// K = 16 pixels = 2 groups of 8 pixels, `rowk` ? one image row apart : consecutive.
template <int PX, int PG>
__global__ void __launch_bounds__(128, 1) mnn_kernel(const __nv_bfloat16* X, const __nv_bfloat16* G, float* D, int xrows,
                                                      int grows, int box_w, int gbox_w, int xstart, int gstart, int rowk) {
  constexpr int CX = PX / 16, CG = PG / 16;
  constexpr uint32_t LX = PX == 128 ? kLayoutSW128 : (PX == 64 ? kLayoutSW64 : kLayoutSW32);
  constexpr uint32_t LG = PG == 128 ? kLayoutSW128 : (PG == 64 ? kLayoutSW64 : kLayoutSW32);
  constexpr int NN = 3 * (PG / 2);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sx = smem;
  uint8_t* sg = smem + 49152;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152 + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < xrows * CX; i += blockDim.x) {
    const int r = i / CX, c = i % CX;
    const uint32_t addr = smem_u32(sx) + r * PX;
    *reinterpret_cast<uint4*>(sx + r * PX + ((c ^ ((addr >> 7) & (CX - 1))) << 4)) = *reinterpret_cast<const uint4*>(X + (size_t)r * (PX / 2) + c * 8);
  }
  for (int i = threadIdx.x; i < grows * CG; i += blockDim.x) {
    const int r = i / CG, c = i % CG;
    const uint32_t addr = smem_u32(sg) + r * PG;
    *reinterpret_cast<uint4*>(sg + r * PG + ((c ^ ((addr >> 7) & (CG - 1))) << 4)) = *reinterpret_cast<const uint4*>(G + (size_t)r * (PG / 2) + c * 8);
  }
  if (warp == 0 && lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(false, true, true, 128, NN);
    if (elect_one()) {
      const uint64_t da = make_desc(smem_u32(sx) + xstart * PX, PX, rowk ? box_w * PX : 8 * PX, LX);
      const uint64_t db = make_desc(smem_u32(sg) + gstart * PG, gbox_w * PG, rowk ? gbox_w * PG : 8 * PG, LG);
      umma<false>(tmem_base, da, db, idesc, 0);
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < NN; c0 += 16) {
    uint32_t v[32];
    // 16 columns at a time (x32 load of a 32-column aligned window would overrun NN = 48)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(tmem_base + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 16; ++e) D[(size_t)(warp * 32 + lane) * NN + c0 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int P>
static void run_check_narrow_k() {
  const int rows = 640, CH = P / 2;
  std::vector<__nv_bfloat16> hX(rows * CH), hB(16 * CH);
  std::vector<float> fX(rows * CH), fB(16 * CH);
  srand(5);
  for (size_t i = 0; i < hX.size(); ++i) { fX[i] = bf((rand() % 31 - 15) / 8.f); hX[i] = __float2bfloat16(fX[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = bf((rand() % 9 - 4) / 4.f); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dX, *dB; float* dD;
  CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * 16 * 4));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const int smem = 49152 + 8192 + 1024 + 256;
  CK(cudaFuncSetAttribute(shiftn_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int sbo_rows : {8, 10, 18, 34}) {
    for (int start_row : {0, 1, 3, 8, 19}) {
      CK(cudaMemset(dD, 0xff, 128 * 16 * 4));
      shiftn_kernel<P><<<1, 128, smem>>>(dX, dB, dD, rows, start_row, sbo_rows * P);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("check(d) P=%d sbo_rows=%d start=%d: CUDA error %s\n", P, sbo_rows, start_row, cudaGetErrorString(e)); exit(1); }
      std::vector<float> hD(128 * 16);
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int m = 0; m < 128; ++m) {
        const int r = start_row + (m / 8) * sbo_rows + (m % 8);
        for (int n = 0; n < 16; ++n) {
          double ref = 0;
          for (int k = 0; k < CH; ++k) ref += (double)fX[r * CH + k] * fB[n * CH + k];
          double er = fabs(ref - hD[m * 16 + n]);
          if (!(er <= maxerr)) maxerr = er;
        }
      }
      printf("check(d) K-major SW%d sbo=%2d rows start_row=%2d: max abs err %.3g %s\n", P, sbo_rows, start_row, maxerr,
             maxerr < 1e-3 ? "OK" : "MISMATCH");
    }
  }
}

template <int PX, int PG>
static void run_check_narrow_mn() {
  const int xrows = 24 * 16, grows = 24 * 16, CX = PX / 2, CGc = PG / 2, NN = 3 * CGc;
  std::vector<__nv_bfloat16> hX(xrows * CX), hG(grows * CGc);
  std::vector<float> fX(xrows * CX), fG(grows * CGc);
  srand(7);
  for (size_t i = 0; i < hX.size(); ++i) { fX[i] = bf((rand() % 31 - 15) / 8.f); hX[i] = __float2bfloat16(fX[i]); }
  for (size_t i = 0; i < hG.size(); ++i) { fG[i] = bf((rand() % 9 - 4) / 4.f); hG[i] = __float2bfloat16(fG[i]); }
  __nv_bfloat16 *dX, *dG; float* dD;
  CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dG, hG.size() * 2)); CK(cudaMalloc(&dD, 128 * NN * 4));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dG, hG.data(), hG.size() * 2, cudaMemcpyHostToDevice));
  const int smem = 49152 + 32768 + 1024 + 256;
  CK(cudaFuncSetAttribute(mnn_kernel<PX, PG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct Case { int box_w, gbox_w, xstart, gstart, rowk; };
  const Case cases[] = {{24, 16, 0, 0, 0}, {24, 16, 25, 16, 0}, {24, 16, 51, 35, 0}, {18, 16, 0, 0, 1}, {18, 16, 19, 16, 1}, {19, 16, 40, 8, 1}};
  for (const Case& c : cases) {
    CK(cudaMemset(dD, 0xff, 128 * NN * 4));
    mnn_kernel<PX, PG><<<1, 128, smem>>>(dX, dG, dD, xrows, grows, c.box_w, c.gbox_w, c.xstart, c.gstart, c.rowk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("check(e) PX=%d PG=%d: CUDA error %s\n", PX, PG, cudaGetErrorString(e)); exit(1); }
    std::vector<float> hD(128 * NN);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m) {
      const int h = m / CX, ch = m % CX;
      for (int n = 0; n < NN; ++n) {
        const int sft = n / CGc, co = n % CGc;
        double ref = 0;
        for (int k = 0; k < 16; ++k) {
          const int xo = c.rowk ? (k / 8) * c.box_w + (k % 8) : k;
          const int go = c.rowk ? (k / 8) * c.gbox_w + (k % 8) : k;
          ref += (double)fX[(c.xstart + xo + h) * CX + ch] * fG[(c.gstart + go + sft * c.gbox_w) * CGc + co];
        }
        const double er = fabs(ref - hD[m * NN + n]);
        if (!(er <= maxerr)) maxerr = er;
      }
    }
    printf("check(e) MN-major X SW%d G SW%d box_w=%d xstart=%d gstart=%d rowk=%d: max abs err %.3g %s\n", PX, PG, c.box_w,
           c.xstart, c.gstart, c.rowk, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
  }
}

static void run_check_narrow() {
  run_check_narrow_k<32>();
  run_check_narrow_k<64>();
  run_check_narrow_mn<32, 32>();
  run_check_narrow_mn<64, 32>();
  run_check_narrow_mn<32, 64>();
  run_check_narrow_mn<64, 64>();
  run_check_narrow_mn<128, 64>();
}

static void run_check() {
  // (a)
  for (int N : {64, 128, 256}) {
    std::vector<__nv_bfloat16> hA(256 * 64), hB(N * 64);
    std::vector<float> fA(256 * 64), fB(N * 64);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = bf((rand() % 17 - 8) / 8.f); hA[i] = __float2bfloat16(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = bf((rand() % 13 - 6) / 4.f); hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 256 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, 256 * N * 4));
    const int smem = 16384 + 16384 + 1024 + 256;
    CK(cudaFuncSetAttribute(pair_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pair_gemm_kernel<<<2, 128, smem>>>(dA, dB, dD, N);
    CK(cudaDeviceSynchronize());
    std::vector<float> hD(256 * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)fA[m * 64 + k] * fB[n * 64 + k];
        double e = fabs(ref - hD[m * N + n]);
        if (!(e <= maxerr)) maxerr = e;
      }
    printf("check(a) cta_group::2 M=256 N=%d K=64: max abs err %.3g %s\n", N, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  // (b)
  const int rows = 384;
  std::vector<__nv_bfloat16> hX(rows * 64), hB(64 * 64);
  std::vector<float> fX(rows * 64), fB(64 * 64);
  srand(2);
  for (size_t i = 0; i < hX.size(); ++i) { fX[i] = bf((rand() % 31 - 15) / 8.f); hX[i] = __float2bfloat16(fX[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = bf((rand() % 9 - 4) / 4.f); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dX, *dB; float* dD;
  CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * 64 * 4));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const int smem = 49152 + 8192 + 1024 + 256;
  CK(cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int sbo : {1024, 1280, 2304}) {
    for (int start_row : {0, 1, 3, 8, 9, 19}) {
      for (int ubo = 0; ubo < 2; ++ubo) {
        CK(cudaMemset(dD, 0xff, 128 * 64 * 4));
        shift_kernel<<<1, 128, smem>>>(dX, dB, dD, rows, start_row, sbo, ubo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("check(b) sbo=%d start=%d base_offset=%d: CUDA error %s\n", sbo, start_row, ubo, cudaGetErrorString(e)); exit(1); }
        std::vector<float> hD(128 * 64);
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) {
          const int r = start_row + (m / 8) * (sbo / 128) + (m % 8);
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += (double)fX[r * 64 + k] * fB[n * 64 + k];
            double er = fabs(ref - hD[m * 64 + n]);
            if (!(er <= maxerr)) maxerr = er;
          }
        }
        printf("check(b) sbo=%4d start_row=%2d base_offset_field=%d: max abs err %.3g %s\n", sbo, start_row, ubo, maxerr,
               maxerr < 1e-3 ? "OK" : "MISMATCH");
      }
    }
  }
}

static void run_check_mn() {
  const int xrows = 384;
  std::vector<__nv_bfloat16> hX(xrows * 64), hY(64 * 64);
  std::vector<float> fX(xrows * 64), fY(64 * 64);
  srand(3);
  for (size_t i = 0; i < hX.size(); ++i) { fX[i] = bf((rand() % 31 - 15) / 8.f); hX[i] = __float2bfloat16(fX[i]); }
  for (size_t i = 0; i < hY.size(); ++i) { fY[i] = bf((rand() % 9 - 4) / 4.f); hY[i] = __float2bfloat16(fY[i]); }
  __nv_bfloat16 *dX, *dY; float* dD;
  CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dY, hY.size() * 2)); CK(cudaMalloc(&dD, 128 * 64 * 4));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dY, hY.data(), hY.size() * 2, cudaMemcpyHostToDevice));
  const int smem = 49152 + 8192 + 1024 + 256;
  CK(cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct Case { int box_w, a, dx0, lbo_pixels; };   // lbo in pixels (x128 B); 80 = a second box 10 rows x 8 px further
  const Case cases[] = {{8, 0, 0, 80}, {8, 1, 0, 80}, {8, 2, 0, 80}, {9, 0, 0, 1}, {9, 1, 0, 1}, {9, 2, 0, 1}, {10, 1, 1, 1}, {10, 2, 0, 100}};
  for (const Case& c : cases) {
    CK(cudaMemset(dD, 0xff, 128 * 64 * 4));
    mn_kernel<<<1, 128, smem>>>(dX, dY, dD, xrows, c.box_w, c.a, c.dx0, c.lbo_pixels * 128);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("check(c) box_w=%d: CUDA error %s\n", c.box_w, cudaGetErrorString(e)); exit(1); }
    std::vector<float> hD(128 * 64);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m) {
      const int h = m / 64, ch = m % 64;
      for (int n = 0; n < 64; ++n) {
        double ref = 0;
        for (int i = 0; i < 8; ++i)
          for (int j = 0; j < 8; ++j) {
            const int xp = (i + c.a) * c.box_w + j + c.dx0 + h * c.lbo_pixels;
            ref += (double)fX[xp * 64 + ch] * fY[(i * 8 + j) * 64 + n];
          }
        const double er = fabs(ref - hD[m * 64 + n]);
        if (!(er <= maxerr)) maxerr = er;
      }
    }
    printf("check(c) MN-major box_w=%2d dy=%d dx0=%d lbo=%4d px: max abs err %.3g %s\n", c.box_w, c.a, c.dx0, c.lbo_pixels, maxerr,
           maxerr < 1e-3 ? "OK" : "MISMATCH");
  }
}

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "all";
  if (!strcmp(mode, "check") || !strcmp(mode, "all")) run_check();
  if (!strcmp(mode, "checkmn") || !strcmp(mode, "all")) run_check_mn();
  if (!strcmp(mode, "checknarrow") || !strcmp(mode, "all")) run_check_narrow();
  if (!strcmp(mode, "rate") || !strcmp(mode, "all")) run_rate();
  if (!strcmp(mode, "rate2") || !strcmp(mode, "all")) run_rate2();
  return 0;
}
