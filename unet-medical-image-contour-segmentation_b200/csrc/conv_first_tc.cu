// First convolution of the network on the tensor cores (nn.Conv2d(C_in <= 4, 64, 3, padding=1), unet_parts.py:15 inside
// `inc`, unet_model.py:15): K = 9 * C_in <= 36 is far too short for a TMA-staged implicit GEMM (one 128-byte swizzle
// row holds 64 bf16), so the im2col rows are BUILT by threads: a producer thread owns one output pixel, gathers its
// 3 x 3 x C_in neighbourhood from global memory (L1-resident: neighbours overlap) and writes the K-major, 128-byte
// swizzled A row (zero padded to a multiple of 16) straight into shared memory; one tcgen05.mma per 16 values of K
// (M = 128 pixels, N = 64) accumulates in TMEM; the epilogue warps read the accumulator back, apply either nothing
// (+ BatchNorm batch statistics of the rounded tile, training) or the folded eval-mode BatchNorm + ReLU (inference),
// and each lane stores its pixel's 128-byte NHWC row.  The layer is HBM bound on its output (2 B x 64 channels per
// pixel); the CUDA-core kernels it replaces were FMA bound (22 TFLOP/s of fp32: 0.23 ms at C2, 1.30 + 0.39 ms at C5).
//
// A tile = 128 CONSECUTIVE pixels in (b, h, w) order, so any H and W work and a warp's 32 output rows are one
// contiguous 4 KB run.  Warps 0-3 produce (3 A stages), warp 8 issues the MMAs, warps 4-7 drain (2 accumulators).
#include <cstring>

#include "tc_common.cuh"

namespace ub {

struct FirstTcParams {
  const __nv_bfloat16* x;        // [npix][ld_in]
  const __nv_bfloat16* wp;       // packed [64][9 * Cin], tap order of the descriptor
  __nv_bfloat16* y;              // [npix][ld_out]
  const float* affine;           // MODE 1: scale[64] then shift[64]
  float* stats_ws;               // MODE 0: [grid * 4][2][64] or null
  long long ld_in, ld_out;
  long long npix;
  int H, W;
  int tap_dy[9], tap_dx[9];
  int ntiles;
};

constexpr int kFtThreads = 288, kFtStages = 3, kFtAcc = 2;
constexpr uint32_t kFtA = 128 * 128, kFtB = 64 * 128;

template <int CIN, int MODE>
__global__ void __launch_bounds__(kFtThreads, 2) first_tc_kernel(const __grid_constant__ FirstTcParams p) {
  constexpr int K = 9 * CIN, KMMA = (K + 15) / 16, KCH = KMMA * 2;      // 16-byte chunks of an A / B row in use
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;                              // 3 x 16 KB
  uint8_t* b_tile = a_ring + kFtStages * kFtA;         // 8 KB
  uint8_t* stage = b_tile + kFtB;                      // 4 epilogue warps x 4 KB (statistics transposition)
  float* coef = reinterpret_cast<float*>(stage + 4 * 4096);          // 128 floats
  uint64_t* a_full = reinterpret_cast<uint64_t*>(coef + 128);
  uint64_t* a_empty = a_full + kFtStages;
  uint64_t* t_full = a_empty + kFtStages;
  uint64_t* t_empty = t_full + kFtAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + kFtAcc);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kFtStages; ++s) { mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kFtAcc; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 128); }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 128);
  // weights: thread n < 64 writes row n of the K-major swizzled B tile (zero padded), once per CTA
  if (threadIdx.x < 64) {
    const int n = threadIdx.x;
    const __nv_bfloat16* src = p.wp + (long long)n * K;
    uint32_t w32[KCH * 4];
#pragma unroll
    for (int i = 0; i < KCH * 4; ++i) {
      const int k0 = 2 * i, k1 = 2 * i + 1;
      const uint32_t lo = k0 < K ? (uint32_t)__bfloat16_as_ushort(src[k0]) : 0u;
      const uint32_t hi = k1 < K ? (uint32_t)__bfloat16_as_ushort(src[k1]) : 0u;
      w32[i] = lo | (hi << 16);
    }
    const uint32_t row = smem_u32(b_tile) + n * 128;
#pragma unroll
    for (int c = 0; c < KCH; ++c)
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((c ^ (n & 7)) << 4)), "r"(w32[4 * c]),
                   "r"(w32[4 * c + 1]), "r"(w32[4 * c + 2]), "r"(w32[4 * c + 3]) : "memory");
  }
  if (MODE == 1 && threadIdx.x >= 64 && threadIdx.x < 192) coef[threadIdx.x - 64] = p.affine[threadIdx.x - 64];
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------ producers: one pixel (= one A row) per thread
    const int r = threadIdx.x;                          // row of the tile
    // the 9 * C_in raw values of the NEXT tile are requested as soon as this tile's row is published, so their latency
    // overlaps the MMA / epilogue of the tile and the other CTA of the SM
    uint32_t raw[K];
    auto gather = [&](int tile) {
      const long long pix = (long long)tile * 128 + r;
      const bool live = tile < p.ntiles && pix < p.npix;
      const int j = live ? (int)(pix % p.W) : 0;
      const long long rest = live ? pix / p.W : 0;
      const int i = (int)(rest % p.H);
      const long long img = rest - i;                   // b * H
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = i + p.tap_dy[t], xx = j + p.tap_dx[t];
        const bool in = live && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        const __nv_bfloat16* src = p.x + ((img + yy) * p.W + xx) * p.ld_in;
#pragma unroll
        for (int c = 0; c < CIN; ++c) raw[t * CIN + c] = in ? (uint32_t)__bfloat16_as_ushort(__ldg(src + c)) : 0u;
      }
    };
    uint32_t s = 0, ph = 1;
    gather(blockIdx.x);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      uint32_t v32[KCH * 4];
#pragma unroll
      for (int i = 0; i < KCH * 4; ++i) v32[i] = 0u;
#pragma unroll
      for (int k = 0; k < K; ++k) v32[k >> 1] |= raw[k] << ((k & 1) * 16);
      mbar_wait(&a_empty[s], ph);
      const uint32_t row = smem_u32(a_ring) + s * kFtA + r * 128;
#pragma unroll
      for (int c = 0; c < KCH; ++c)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((c ^ (r & 7)) << 4)), "r"(v32[4 * c]),
                     "r"(v32[4 * c + 1]), "r"(v32[4 * c + 2]), "r"(v32[4 * c + 3]) : "memory");
      fence_async_smem();                               // generic-proxy writes -> visible to the tensor core
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_full[s])) : "memory");
      // (requested AFTER the proxy fence: the fence waits for the thread's outstanding loads, a request issued before
      // it would be drained there -- first version, 2x slower)
      gather(tile + gridDim.x);
      if (++s == kFtStages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 8) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc(false, false, false, 128, 64);
    const uint64_t a_t = make_desc(smem_u32(a_ring), 16, 1024), b_t = make_desc(smem_u32(b_tile), 16, 1024);
    uint32_t s = 0, ph = 0, acc = 0, pacc = 1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      mbar_wait(&a_full[s], ph);
      mbar_wait(&t_empty[acc], pacc);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < KMMA; ++kk)
          umma<false>(tmem_base + acc * 64, a_t + s * (kFtA >> 4) + 2 * kk, b_t + 2 * kk, idesc, kk > 0 ? 1u : 0u);
        umma_commit(&a_empty[s]);
        umma_commit(&t_full[acc]);
      }
      __syncwarp();
      if (++s == kFtStages) { s = 0; ph ^= 1; }
      if (++acc == kFtAcc) { acc = 0; pacc ^= 1; }
    }
  } else {
    // ------------------------------------------------ epilogue: warp w drains TMEM lanes [32 (w % 4), +32)
    const int quad = warp & 3;
    const uint32_t stg = smem_u32(stage) + quad * 4096;
    uint32_t acc = 0, pacc = 0;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;       // MODE 0: channels 2 * lane, 2 * lane + 1
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const long long pix = (long long)tile * 128 + quad * 32 + lane;
      mbar_wait(&t_full[acc], pacc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 64;
      const bool live = pix < p.npix;
      // two halves of 32 columns (keeps 32 + 16 values live instead of 64 + 32): TMEM -> [affine + ReLU] -> bf16 ->
      // the warp's staging tile (32 rows x 128 B, swizzled); rows past the end are staged as zeros (statistics)
      __syncwarp();                                                       // the previous tile's readers are done
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld32(taddr + h * 32, v);
        if (h == 1) {
          tc_fence_before();
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&t_empty[acc])) : "memory");   // accumulator free
        }
        uint32_t o[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float x0 = __uint_as_float(v[2 * c]), x1 = __uint_as_float(v[2 * c + 1]);
          if (MODE == 1) {
            const int ch = h * 32 + 2 * c;
            x0 = fmaxf(fmaf(x0, coef[ch], coef[64 + ch]), 0.f);
            x1 = fmaxf(fmaf(x1, coef[ch + 1], coef[64 + ch + 1]), 0.f);
          }
          o[c] = live ? pack_bf16x2(x0, x1) : 0u;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 128 + (((h * 4 + c) ^ (lane & 7)) << 4)),
                       "r"(o[4 * c]), "r"(o[4 * c + 1]), "r"(o[4 * c + 2]), "r"(o[4 * c + 3]) : "memory");
      }
      if (++acc == kFtAcc) { acc = 0; pacc ^= 1; }
      __syncwarp();
      // store transposed: one instruction = 4 rows x 128 B, a contiguous 512-byte run when the rows are packed (a lane
      // writing its own 128-byte row would touch 32 lines per instruction)
      {
        const long long pix0 = (long long)tile * 128 + quad * 32;
        const int ch = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3);
          uint4 q;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                       : "r"(stg + rr * 128 + ((ch ^ (rr & 7)) << 4)) : "memory");
          if (pix0 + rr < p.npix) *reinterpret_cast<uint4*>(p.y + (pix0 + rr) * p.ld_out + ch * 8) = q;
        }
      }
      if (MODE == 0 && p.stats_ws) {
        // BatchNorm statistics of the rounded tile: lane = channel pair sums down the staged rows
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          uint32_t u;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u)
                       : "r"(stg + rr * 128 + ((((lane >> 2) ^ (rr & 7)) << 4) | ((lane & 3) << 2))) : "memory");
          const float a = __uint_as_float(u << 16), b = __uint_as_float(u & 0xffff0000u);
          s0 += a; q0 = fmaf(a, a, q0); s1 += b; q1 = fmaf(b, b, q1);
        }
      }
    }
    if (MODE == 0 && p.stats_ws) {
      float* dst = p.stats_ws + ((long long)blockIdx.x * 4 + quad) * 128;
      dst[2 * lane] = s0; dst[2 * lane + 1] = s1;
      dst[64 + 2 * lane] = q0; dst[64 + 2 * lane + 1] = q1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_base, 128);
  }
}

static bool first_tc_shape_ok(const unetb200_gconv_t* d) {
  static const bool off = getenv("UNETB200_NO_FIRST_TC") != nullptr;
  if (off || d->dtype != UNETB200_BF16 || d->N != 64 || d->Cin < 1 || d->Cin > 4) return false;
  if (d->ntaps != 9 || d->in_scale != 1 || d->out_scale != 1 || d->nquad != 1) return false;
  if (d->in_off_y || d->in_off_x || d->out_off_y || d->out_off_x) return false;
  if (d->Hm != d->Hout || d->Wm != d->Wout || d->Hm != d->Hin || d->Wm != d->Win) return false;
  if (d->ld_out % 8) return false;
  return (long long)d->B * d->Hm * d->Wm < (1LL << 31) - 256;
}

int first_tc_supported(const unetb200_gconv_t* d, const void* y) {
  return first_tc_shape_ok(d) && aligned16(y) ? 1 : 0;
}

static int first_tc_grid(const unetb200_gconv_t* d, int* ntiles) {
  const long long npix = (long long)d->B * d->Hm * d->Wm;
  *ntiles = (int)((npix + 127) / 128);
  const int slots = 2 * sm_count();                     // 73 KB of shared memory, 128 TMEM columns: two CTAs per SM fit
  return *ntiles < slots ? *ntiles : slots;
}

long long first_tc_stats_rows(const unetb200_gconv_t* d) {
  if (!first_tc_shape_ok(d)) return 0;
  int nt;
  return (long long)first_tc_grid(d, &nt) * 4;
}

template <int CIN, int MODE>
static int first_tc_launch(const FirstTcParams& P, int grid, cudaStream_t s) {
  constexpr int smem = kFtStages * kFtA + kFtB + 4 * 4096 + 512 + 128 + 1024;
  if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(&first_tc_kernel<CIN, MODE>), smem, "first_tc smem attribute"))
    return rc;
  first_tc_kernel<CIN, MODE><<<grid, kFtThreads, smem, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "first_tc launch");
  return 0;
}

int first_tc_fprop(const unetb200_gconv_t* d, const void* x, const void* wp, void* y, double* stats, float* stats_ws,
                   const float* affine, cudaStream_t s) {
  if (!first_tc_supported(d, y)) { set_error("first_tc_fprop: unsupported shape"); return UNETB200_E_INVALID; }
  FirstTcParams P;
  memset(&P, 0, sizeof(P));
  P.x = (const __nv_bfloat16*)x; P.wp = (const __nv_bfloat16*)wp; P.y = (__nv_bfloat16*)y;
  P.affine = affine;
  P.stats_ws = (stats && !affine) ? stats_ws : nullptr;
  P.ld_in = d->ld_in; P.ld_out = d->ld_out;
  P.npix = (long long)d->B * d->Hm * d->Wm;
  P.H = d->Hm; P.W = d->Wm;
  for (int t = 0; t < 9; ++t) { P.tap_dy[t] = d->tap_dy[t]; P.tap_dx[t] = d->tap_dx[t]; }
  const int grid = first_tc_grid(d, &P.ntiles);
  int rc;
#define UB_FT(CIN)                                                                                  \
  case CIN: rc = affine ? first_tc_launch<CIN, 1>(P, grid, s) : first_tc_launch<CIN, 0>(P, grid, s); break;
  switch (d->Cin) {
    UB_FT(1) UB_FT(2) UB_FT(3)
    default: rc = affine ? first_tc_launch<4, 1>(P, grid, s) : first_tc_launch<4, 0>(P, grid, s); break;
  }
#undef UB_FT
  if (rc) return rc;
  if (P.stats_ws) return launch_stats_reduce(stats_ws, (long long)grid * 4, 2 * d->N, stats, s);
  return 0;
}

}  // namespace ub
