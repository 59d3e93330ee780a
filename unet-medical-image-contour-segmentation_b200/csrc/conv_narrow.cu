// 3x3 convolutions with narrow channel counts on the tensor cores (fprop and dgrad of nn.Conv2d(k=3, padding=1) with
// C_in <= 64 and C_out <= 64, not both 64): the layers of the reference's light variants UNet_S / UNet_T / UNet_SA
// (unet_model.py:52-189; UNet_S is what train.py:253 builds) and of their decoders.
//
// These shapes do not fit the TMA-staged kernels (a 128-byte swizzle row is 64 bf16 channels), and on the CUDA-core
// implicit GEMM they cost 3 ms per full-resolution layer where the HBM floor is 0.05 ms.  Generalisation of
// conv_first_tc.cu: the im2col rows are BUILT by threads.  A producer thread owns one output pixel, gathers its nine
// taps x C_in channels from global memory (16-byte loads, L1 resident: neighbours overlap) and stores them at
// k = tap * C_in + c of the pixel's K-major row -- ceil(K / 64) swizzled [128 rows x 128 B] chunk tiles per stage,
// the K padding zeroed once; one tcgen05.mma per 16 values of K (M = 128 pixels, N = 16 / 32 / 64) accumulates in
// TMEM; the epilogue warps apply nothing (+ BatchNorm batch statistics of the rounded tile, training) or the folded
// eval-mode BatchNorm + ReLU, stage the tile and store it with 512-byte contiguous runs.
// A tile = 128 consecutive pixels in (b, h, w) order (any H, W).  Warps 0-3 produce, warp 8 issues, warps 4-7 drain.
#include <cstring>

#include "tc_common.cuh"

namespace ub {

struct NarrowParams {
  const __nv_bfloat16* x;        // [npix][ld_in]
  const __nv_bfloat16* wp;       // packed [N][9 * Cin], tap order of the descriptor
  __nv_bfloat16* y;              // [npix][ld_out]
  const float* affine;           // MODE 1: scale[N] then shift[N]
  float* stats_ws;               // MODE 0: [grid * 4][2][N] or null
  long long ld_in, ld_out;
  long long npix;
  int H, W, N;                   // N = real output channels (<= NT)
  int tap_dy[9], tap_dx[9];
  long long tap_off[9];          // (dy * W + dx) * ld_in, elements
  int ntiles;
};

constexpr int kNwThreads = 288;

template <int CIN, int NT>
struct NarrowCfg {
  static constexpr int K = 9 * CIN, NKC = (K + 63) / 64, KMMA = (K + 15) / 16;
  static constexpr int STAGES = NKC <= 2 ? 3 : (NKC <= 5 ? 2 : 1);      // C_in = 16: 2 x 48 KB -> two CTAs per SM gather at once
  static constexpr uint32_t kAStage = NKC * 16384u, kB = NKC * NT * 128u;
  static constexpr int kTmemCols = NT == 64 ? 128 : 64;
  static constexpr int smem = STAGES * (int)kAStage + (int)kB + 4 * 32 * NT * 2 + 2 * 64 * 4 + 256 + 1024;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// C_in <= 16: two CTAs per SM (registers capped at 112, 105 KB of shared memory each): a lone producer warp per
// scheduler is latency bound, a second CTA doubles what is in flight
template <int CIN, int NT, int MODE>
__global__ void __launch_bounds__(kNwThreads, (CIN <= 16 && NT <= 32) ? 2 : 1) narrow_conv_kernel(const __grid_constant__ NarrowParams p) {
  using Cfg = NarrowCfg<CIN, NT>;
  constexpr int K = Cfg::K, KMMA = Cfg::KMMA, STAGES = Cfg::STAGES;
  constexpr uint32_t kAStage = Cfg::kAStage;
  constexpr int RS = NT * 2;                            // staging row bytes
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  uint8_t* b_tile = a_ring + STAGES * kAStage;         // NKC x [NT rows x 128 B]
  uint8_t* stage = b_tile + Cfg::kB;                   // 4 epilogue warps x 32 rows x RS
  float* coef = reinterpret_cast<float*>(stage + 4 * 32 * RS);       // scale[64] shift[64]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(coef + 128);
  uint64_t* a_empty = a_full + STAGES;
  uint64_t* t_full = a_empty + STAGES;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 128); }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  // zero the A ring once (the K padding of the last chunk tile stays zero) and the B tile, then fill B: row n of chunk
  // tile j holds Wp[n][64 j .. 64 j + 63]
  for (uint32_t i = threadIdx.x; i < (STAGES * kAStage + Cfg::kB) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int e = threadIdx.x; e < p.N * K; e += blockDim.x) {
    const int n = e / K, k = e - n * K;
    const int j = k >> 6, kk = k & 63;
    reinterpret_cast<__nv_bfloat16*>(b_tile + j * (NT * 128) + n * 128 + (((kk >> 3) ^ (n & 7)) << 4))[kk & 7] = p.wp[e];
  }
  if (MODE == 1 && threadIdx.x < 128) {
    const int c = threadIdx.x & 63, which = threadIdx.x >> 6;
    coef[threadIdx.x] = c < p.N ? p.affine[which * p.N + c] : 0.f;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------ producers: one pixel (= one A row) per thread
    const int r = threadIdx.x;
    uint32_t s = 0, ph = 1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      // (32-bit index arithmetic: npix < 2^31; a lone producer warp per scheduler is instruction-latency bound)
      const unsigned pix = (unsigned)tile * 128u + (unsigned)r;
      const bool live = pix < (unsigned)p.npix;
      const unsigned rest = live ? pix / (unsigned)p.W : 0u;
      const int j0 = live ? (int)(pix - rest * (unsigned)p.W) : 0;
      const int i0 = (int)(rest % (unsigned)p.H);
      const __nv_bfloat16* pc = p.x + (long long)(live ? pix : 0u) * p.ld_in;      // the pixel itself; taps are offsets from it
      mbar_wait(&a_empty[s], ph);
      const uint32_t row = smem_u32(a_ring) + s * kAStage + r * 128;
      if constexpr (CIN >= 8) {
        constexpr int VPT = CIN / 8;                    // 16-byte vectors per tap
        // taps per batch: every load of a batch is in flight before the first store (the gather is latency bound:
        // three taps per batch cost three global round trips per tile, 0.34 ms per full-resolution 16 -> 16 layer)
        constexpr int TB = CIN <= 8 ? 9 : (CIN <= 32 ? 5 : 3);
#pragma unroll
        for (int t0 = 0; t0 < 9; t0 += TB) {
          uint4 v[TB * VPT];
#pragma unroll
          for (int tt = 0; tt < TB; ++tt) {
            const int t = t0 + tt;
            if (t < 9) {
              const int yy = i0 + p.tap_dy[t], xx = j0 + p.tap_dx[t];
              const bool in = live && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
              const uint4* src = reinterpret_cast<const uint4*>(pc + p.tap_off[t]);
#pragma unroll
              for (int q = 0; q < VPT; ++q) v[tt * VPT + q] = in ? __ldg(src + q) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
#pragma unroll
          for (int tt = 0; tt < TB; ++tt)
#pragma unroll
            for (int q = 0; q < VPT; ++q) {
              if (t0 + tt < 9) {
                const int k = (t0 + tt) * CIN + 8 * q;           // compile-time
                const uint32_t dst = row + (k >> 6) * 16384 + ((((k & 63) >> 3) ^ (r & 7)) << 4);
                const uint4 w = v[tt * VPT + q];
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w)
                             : "memory");
              }
            }
        }
      } else {
        constexpr int KCH = KMMA * 2;
        uint32_t v32[KCH * 4];
#pragma unroll
        for (int i = 0; i < KCH * 4; ++i) v32[i] = 0u;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int yy = i0 + p.tap_dy[t], xx = j0 + p.tap_dx[t];
          const bool in = live && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
          const __nv_bfloat16* src = pc + p.tap_off[t];
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            const int k = t * CIN + c;
            const uint32_t val = in ? (uint32_t)__bfloat16_as_ushort(__ldg(src + c)) : 0u;
            v32[k >> 1] |= val << ((k & 1) * 16);
          }
        }
#pragma unroll
        for (int c = 0; c < KCH; ++c)
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((c ^ (r & 7)) << 4)), "r"(v32[4 * c]),
                       "r"(v32[4 * c + 1]), "r"(v32[4 * c + 2]), "r"(v32[4 * c + 3]) : "memory");
      }
      fence_async_smem();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_full[s])) : "memory");
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 8) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc(false, false, false, 128, NT);
    const uint64_t a_t = make_desc(smem_u32(a_ring), 16, 1024), b_t = make_desc(smem_u32(b_tile), 16, 1024);
    uint32_t s = 0, ph = 0, acc = 0, pacc = 1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      mbar_wait(&a_full[s], ph);
      mbar_wait(&t_empty[acc], pacc);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < KMMA; ++kk) {
          const int j = kk >> 2, q = kk & 3;            // chunk tile, 32-byte slice inside it
          umma<false>(tmem_base + acc * NT, a_t + ((s * kAStage + j * 16384) >> 4) + 2 * q,
                      b_t + ((j * NT * 128) >> 4) + 2 * q, idesc, kk > 0 ? 1u : 0u);
        }
        umma_commit(&a_empty[s]);
        umma_commit(&t_full[acc]);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
  } else {
    // ------------------------------------------------ epilogue: warp w drains TMEM lanes [32 (w % 4), +32)
    const int quad = warp & 3;
    const uint32_t stg = smem_u32(stage) + quad * 32 * RS;
    auto swz = [](int row) { return NT == 64 ? (row & 7) : (NT == 32 ? ((row >> 1) & 3) : 0); };
    constexpr int PP = NT / 2, G = 32 / PP;             // channel pairs per row; row groups for the statistics
    const int pair = lane % PP, grp = lane / PP;
    const int cpr = p.N / 8;                            // 16-byte chunks per STORED row (N may be 8 with NT = 16)
    uint32_t acc = 0, pacc = 0;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const long long pix0 = (long long)tile * 128 + quad * 32;
      const bool live = pix0 + lane < p.npix;
      mbar_wait(&t_full[acc], pacc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * NT;
      __syncwarp();
#pragma unroll
      for (int h = 0; h < (NT + 31) / 32; ++h) {
        constexpr int W = NT < 32 ? NT : 32;
        uint32_t v[W];
        if constexpr (NT < 32) tmem_ld16(taddr, v);
        else tmem_ld32(taddr + h * 32, v);
        if (h == (NT + 31) / 32 - 1) {
          tc_fence_before();
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&t_empty[acc])) : "memory");
        }
        uint32_t o[W / 2];
#pragma unroll
        for (int c = 0; c < W / 2; ++c) {
          float x0 = __uint_as_float(v[2 * c]), x1 = __uint_as_float(v[2 * c + 1]);
          if (MODE == 1) {
            const int ch = h * 32 + 2 * c;
            x0 = fmaxf(fmaf(x0, coef[ch], coef[64 + ch]), 0.f);
            x1 = fmaxf(fmaf(x1, coef[ch + 1], coef[64 + ch + 1]), 0.f);
          }
          o[c] = live ? pack_bf16x2(x0, x1) : 0u;
        }
#pragma unroll
        for (int c = 0; c < W / 8; ++c)
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * RS + (((h * 4 + c) ^ swz(lane)) << 4)),
                       "r"(o[4 * c]), "r"(o[4 * c + 1]), "r"(o[4 * c + 2]), "r"(o[4 * c + 3]) : "memory");
      }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
      __syncwarp();
      // transposed store: one instruction = 32 / cpr rows x (16 cpr) bytes, contiguous when the rows are packed
      {
        const int rpi = 32 / cpr, ch = lane % cpr;
        for (int i = 0; i < cpr; ++i) {
          const int rr = i * rpi + lane / cpr;
          uint4 q;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                       : "r"(stg + rr * RS + ((ch ^ swz(rr)) << 4)) : "memory");
          if (pix0 + rr < p.npix) *reinterpret_cast<uint4*>(p.y + (pix0 + rr) * p.ld_out + ch * 8) = q;
        }
      }
      if (MODE == 0 && p.stats_ws) {
#pragma unroll
        for (int rr = grp; rr < 32; rr += G) {
          uint32_t u;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u)
                       : "r"(stg + rr * RS + ((((pair >> 2) ^ swz(rr)) << 4) | ((pair & 3) << 2))) : "memory");
          const float a = __uint_as_float(u << 16), b = __uint_as_float(u & 0xffff0000u);
          s0 += a; q0 = fmaf(a, a, q0); s1 += b; q1 = fmaf(b, b, q1);
        }
      }
    }
    if (MODE == 0 && p.stats_ws) {
#pragma unroll
      for (int o = PP; o < 32; o <<= 1) {               // combine the row groups (fixed order)
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
      }
      float* dst = p.stats_ws + ((long long)blockIdx.x * 4 + quad) * 2 * p.N;
      if (grp == 0 && 2 * pair < p.N) {
        dst[2 * pair] = s0; dst[2 * pair + 1] = s1;
        dst[p.N + 2 * pair] = q0; dst[p.N + 2 * pair + 1] = q1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------ host side
static bool narrow_cin_ok(int c) { return c == 1 || c == 2 || c == 3 || c == 4 || c == 8 || c == 16 || c == 32 || c == 64; }

static bool narrow_shape_ok(const unetb200_gconv_t* d) {
  static const bool off = getenv("UNETB200_NO_NARROW_TC") != nullptr;
  if (off || d->dtype != UNETB200_BF16) return false;
  if (d->ntaps != 9 || d->in_scale != 1 || d->out_scale != 1 || d->nquad != 1) return false;
  if (d->in_off_y || d->in_off_x || d->out_off_y || d->out_off_x) return false;
  if (d->Hm != d->Hout || d->Wm != d->Wout || d->Hm != d->Hin || d->Wm != d->Win) return false;
  if (!narrow_cin_ok(d->Cin) || d->N < 8 || d->N > 64 || (d->N & (d->N - 1))) return false;      // N = 8, 16, 32, 64
  if (d->Cin % 64 == 0 && d->N % 64 == 0) return false;            // the TMA-staged pair kernel covers 64 -> 64
  if (d->Cin <= 4 && d->N == 64) return false;                     // conv_first_tc.cu
  if (d->Cin >= 8 && (d->ld_in % 8)) return false;
  if (d->ld_out % 8) return false;
  bool seen[9] = {false};
  for (int t = 0; t < 9; ++t) {
    const int dy = d->tap_dy[t], dx = d->tap_dx[t];
    if (dy < -1 || dy > 1 || dx < -1 || dx > 1 || seen[(dy + 1) * 3 + dx + 1]) return false;
    seen[(dy + 1) * 3 + dx + 1] = true;
  }
  return (long long)d->B * d->Hm * d->Wm < (1LL << 31) - 256;
}

int narrow_tc_supported(const unetb200_gconv_t* d, const void* x, const void* wp, const void* y) {
  if (!narrow_shape_ok(d)) return 0;
  if (!aligned16(y) || (reinterpret_cast<uintptr_t>(wp) & 1)) return 0;
  if (d->Cin >= 8 && !aligned16(x)) return 0;
  return 1;
}

static int narrow_grid(const unetb200_gconv_t* d, int* ntiles) {
  const long long npix = (long long)d->B * d->Hm * d->Wm;
  *ntiles = (int)((npix + 127) / 128);
  const int per_sm = (d->Cin <= 16 && d->N <= 32) ? 2 : 1;      // see __launch_bounds__ of narrow_conv_kernel
  const int slots = per_sm * sm_count();
  return *ntiles < slots ? *ntiles : slots;
}

long long narrow_tc_stats_rows(const unetb200_gconv_t* d) {
  if (!narrow_shape_ok(d)) return 0;
  int nt;
  return (long long)narrow_grid(d, &nt) * 4;
}

template <int CIN, int NT, int MODE>
static int narrow_launch(const NarrowParams& P, int grid, cudaStream_t s) {
  constexpr int smem = NarrowCfg<CIN, NT>::smem;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(&narrow_conv_kernel<CIN, NT, MODE>), smem, "narrow_conv smem attribute"))
    return rc;
  narrow_conv_kernel<CIN, NT, MODE><<<grid, kNwThreads, smem, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "narrow_conv launch");
  return 0;
}

template <int CIN>
static int narrow_dispatch_n(const NarrowParams& P, int grid, bool affine, cudaStream_t s) {
  const int nt = P.N <= 16 ? 16 : (P.N <= 32 ? 32 : 64);
  if (affine) {
    if (nt == 16) return narrow_launch<CIN, 16, 1>(P, grid, s);
    if (nt == 32) return narrow_launch<CIN, 32, 1>(P, grid, s);
    if constexpr (CIN < 64) return narrow_launch<CIN, 64, 1>(P, grid, s);      // 64 -> 64 is the pair kernel's
  } else {
    if (nt == 16) return narrow_launch<CIN, 16, 0>(P, grid, s);
    if (nt == 32) return narrow_launch<CIN, 32, 0>(P, grid, s);
    if constexpr (CIN < 64) return narrow_launch<CIN, 64, 0>(P, grid, s);
  }
  set_error("narrow_conv: unsupported channel counts");
  return UNETB200_E_INVALID;
}

int narrow_tc_fprop(const unetb200_gconv_t* d, const void* x, const void* wp, void* y, double* stats, float* stats_ws,
                    const float* affine, cudaStream_t s) {
  if (!narrow_tc_supported(d, x, wp, y)) { set_error("narrow_tc_fprop: unsupported shape"); return UNETB200_E_INVALID; }
  NarrowParams P;
  memset(&P, 0, sizeof(P));
  P.x = (const __nv_bfloat16*)x; P.wp = (const __nv_bfloat16*)wp; P.y = (__nv_bfloat16*)y;
  P.affine = affine;
  P.stats_ws = (stats && !affine) ? stats_ws : nullptr;
  P.ld_in = d->ld_in; P.ld_out = d->ld_out;
  P.npix = (long long)d->B * d->Hm * d->Wm;
  P.H = d->Hm; P.W = d->Wm; P.N = d->N;
  for (int t = 0; t < 9; ++t) {
    P.tap_dy[t] = d->tap_dy[t]; P.tap_dx[t] = d->tap_dx[t];
    P.tap_off[t] = ((long long)d->tap_dy[t] * d->Wm + d->tap_dx[t]) * d->ld_in;
  }
  const int grid = narrow_grid(d, &P.ntiles);
  int rc;
  switch (d->Cin) {
    case 1: rc = narrow_dispatch_n<1>(P, grid, affine != nullptr, s); break;
    case 2: rc = narrow_dispatch_n<2>(P, grid, affine != nullptr, s); break;
    case 3: rc = narrow_dispatch_n<3>(P, grid, affine != nullptr, s); break;
    case 4: rc = narrow_dispatch_n<4>(P, grid, affine != nullptr, s); break;
    case 8: rc = narrow_dispatch_n<8>(P, grid, affine != nullptr, s); break;
    case 16: rc = narrow_dispatch_n<16>(P, grid, affine != nullptr, s); break;
    case 32: rc = narrow_dispatch_n<32>(P, grid, affine != nullptr, s); break;
    default: rc = narrow_dispatch_n<64>(P, grid, affine != nullptr, s); break;
  }
  if (rc) return rc;
  if (P.stats_ws) return launch_stats_reduce(stats_ws, (long long)grid * 4, 2 * d->N, stats, s);
  return 0;
}

// ------------------------------------------------------------------------------------------ wgrad
// dW[(t, c)][n] = sum_p x[p + t][c] * g[p][n] for the same narrow layers.  The reduction runs over pixels, so both
// operands are MN-major: the producers build the SAME im2col tile as above (rows = pixels, 128 bytes = 64 values of
// k = t * C_in + c) -- read by the tensor core as A^T with M = 128 values of k (two chunk tiles LBO apart), K = 16
// pixel rows per MMA -- plus the pixel's dY row (N <= 64 channels) as the B tile.  A CTA walks its share of the pixel
// tiles accumulating ALL of dW in TMEM (ceil(K / 128) blocks x N columns <= 320) and writes one fp32 partial at the
// end; the partials (one per CTA) go through the ordinary split reduction.
struct NarrowWParams {
  const __nv_bfloat16* x;        // [npix][ld_in]
  const __nv_bfloat16* gy;       // [npix][ld_out]
  float* partials;               // [grid][9 * Cin][N]
  long long ld_in, ld_out;
  long long npix;
  int H, W, N;
  int tap_dy[9], tap_dx[9];
  long long tap_off[9];
  int ntiles;
};

template <int CIN, int NT>
struct NarrowWCfg {
  static constexpr int K = 9 * CIN, NKC = (K + 63) / 64, MB = (NKC + 1) / 2;
  static constexpr uint32_t kStage = (NKC + 1) * 16384u;            // im2col chunk tiles + the dY tile
  static constexpr int STAGES = kStage * 3 <= 200 * 1024 ? 3 : (kStage * 2 <= 200 * 1024 ? 2 : 1);
  static constexpr int kCols = MB * NT <= 32 ? 32 : (MB * NT <= 64 ? 64 : (MB * NT <= 128 ? 128 : (MB * NT <= 256 ? 256 : 512)));
  static constexpr int smem = STAGES * (int)kStage + 256 + 1024;
};

template <int CIN, int NT>
__global__ void __launch_bounds__(192, 1) narrow_wgrad_kernel(const __grid_constant__ NarrowWParams p) {
  using Cfg = NarrowWCfg<CIN, NT>;
  constexpr int K = Cfg::K, NKC = Cfg::NKC, MB = Cfg::MB, STAGES = Cfg::STAGES;
  constexpr uint32_t kStage = Cfg::kStage;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + STAGES * kStage);
  uint64_t* a_empty = a_full + STAGES;
  uint64_t* t_full = a_empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); }
    mbar_init(t_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, Cfg::kCols);
  for (uint32_t i = threadIdx.x; i < (STAGES * kStage) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);        // K padding / unused dY columns stay zero
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int my_tiles = 0;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) ++my_tiles;

  if (warp < 4) {
    const int r = threadIdx.x;
    uint32_t s = 0, ph = 1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      // (32-bit index arithmetic: npix < 2^31; a lone producer warp per scheduler is instruction-latency bound)
      const unsigned pix = (unsigned)tile * 128u + (unsigned)r;
      const bool live = pix < (unsigned)p.npix;
      const unsigned rest = live ? pix / (unsigned)p.W : 0u;
      const int j0 = live ? (int)(pix - rest * (unsigned)p.W) : 0;
      const int i0 = (int)(rest % (unsigned)p.H);
      const __nv_bfloat16* pc = p.x + (long long)(live ? pix : 0u) * p.ld_in;      // the pixel itself; taps are offsets from it
      mbar_wait(&a_empty[s], ph);
      const uint32_t row = smem_u32(smem) + s * kStage + r * 128;
      // the pixel's dY row -> B tile (chunk tile NKC of the stage)
      {
        const uint4* src = reinterpret_cast<const uint4*>(p.gy + (long long)(live ? pix : 0u) * p.ld_out);
        const int nv = p.N / 8;
        for (int q = 0; q < nv; ++q) {
          const uint4 w = live ? __ldg(src + q) : make_uint4(0u, 0u, 0u, 0u);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(row + NKC * 16384 + ((q ^ (r & 7)) << 4)), "r"(w.x),
                       "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
        }
      }
      if constexpr (CIN >= 8) {
        constexpr int VPT = CIN / 8;
        constexpr int TB = CIN <= 16 ? 9 : (CIN == 32 ? 5 : 3);
#pragma unroll
        for (int t0 = 0; t0 < 9; t0 += TB) {
          uint4 v[TB * VPT];
#pragma unroll
          for (int tt = 0; tt < TB; ++tt) {
            const int t = t0 + tt;
            if (t < 9) {
              const int yy = i0 + p.tap_dy[t], xx = j0 + p.tap_dx[t];
              const bool in = live && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
              const uint4* src = reinterpret_cast<const uint4*>(pc + p.tap_off[t]);
#pragma unroll
              for (int q = 0; q < VPT; ++q) v[tt * VPT + q] = in ? __ldg(src + q) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
#pragma unroll
          for (int tt = 0; tt < TB; ++tt)
#pragma unroll
            for (int q = 0; q < VPT; ++q) {
              if (t0 + tt < 9) {
                const int k = (t0 + tt) * CIN + 8 * q;           // compile-time
                const uint32_t dst = row + (k >> 6) * 16384 + ((((k & 63) >> 3) ^ (r & 7)) << 4);
                const uint4 w = v[tt * VPT + q];
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w)
                             : "memory");
              }
            }
        }
      } else {
        constexpr int KCH = ((K + 15) / 16) * 2;
        uint32_t v32[KCH * 4];
#pragma unroll
        for (int i = 0; i < KCH * 4; ++i) v32[i] = 0u;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int yy = i0 + p.tap_dy[t], xx = j0 + p.tap_dx[t];
          const bool in = live && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
          const __nv_bfloat16* src = pc + p.tap_off[t];
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            const int k = t * CIN + c;
            const uint32_t val = in ? (uint32_t)__bfloat16_as_ushort(__ldg(src + c)) : 0u;
            v32[k >> 1] |= val << ((k & 1) * 16);
          }
        }
#pragma unroll
        for (int c = 0; c < KCH; ++c)
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((c ^ (r & 7)) << 4)), "r"(v32[4 * c]),
                       "r"(v32[4 * c + 1]), "r"(v32[4 * c + 2]), "r"(v32[4 * c + 3]) : "memory");
      }
      fence_async_smem();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_full[s])) : "memory");
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    // ---- end of the walk: drain the accumulators, thread = row (value of k) of each 128-row block
    mbar_wait(t_full, 0);
    tc_fence_after();
    float* out = p.partials + (long long)blockIdx.x * K * p.N;
#pragma unroll 1
    for (int mb = 0; mb < MB; ++mb) {
      const int k = mb * 128 + r;
      const bool dup = (NKC & 1) && mb == MB - 1 && r >= 64;        // odd chunk count: rows 64.. of the last block repeat
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + mb * NT + c0, v);
        if (k < K && !dup) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c0 + e < p.N) out[(long long)k * p.N + c0 + e] = my_tiles > 0 ? __uint_as_float(v[e]) : 0.f;
        }
      }
    }
  } else if (warp == 4) {
    constexpr uint32_t idesc = make_idesc(false, true, true, 128, NT);
    uint32_t s = 0, ph = 0;
    int done = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      mbar_wait(&a_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t base = smem_u32(smem) + s * kStage;
        const uint64_t db = make_desc(base + NKC * 16384, 16384, 1024, kLayoutSW128);
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          const bool odd_last = (NKC & 1) && mb == MB - 1;
          const uint64_t da = make_desc(base + 2 * mb * 16384, odd_last ? 0u : 16384u, 1024, kLayoutSW128);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma<false>(tmem_base + mb * NT, da + ((ks * 2048) >> 4), db + ((ks * 2048) >> 4), idesc,
                        (done > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&a_empty[s]);
      }
      __syncwarp();
      ++done;
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(t_full);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tmem_dealloc(tmem_base, Cfg::kCols);
  }
}

static int narrow_wgrad_grid(const unetb200_gconv_t* d, int* ntiles) {
  const long long npix = (long long)d->B * d->Hm * d->Wm;
  *ntiles = (int)((npix + 127) / 128);
  const int slots = sm_count();                         // all of dW lives in a CTA's TMEM: one CTA per SM
  return *ntiles < slots ? *ntiles : slots;
}

int narrow_wgrad_supported(const unetb200_gconv_t* d, const void* x, const void* gy) {
  static const bool off = getenv("UNETB200_NO_NARROW_WGRAD") != nullptr;
  if (off || !narrow_shape_ok(d)) return 0;
  if (d->Cin >= 8 && x && !aligned16(x)) return 0;
  if (gy && !aligned16(gy)) return 0;
  return 1;
}

int narrow_wgrad_splits(const unetb200_gconv_t* d) {
  int nt;
  return narrow_wgrad_grid(d, &nt);
}

template <int CIN, int NT>
static int narrow_wgrad_launch(const NarrowWParams& P, int grid, cudaStream_t s) {
  constexpr int smem = NarrowWCfg<CIN, NT>::smem;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(&narrow_wgrad_kernel<CIN, NT>), smem, "narrow_wgrad smem attribute"))
    return rc;
  narrow_wgrad_kernel<CIN, NT><<<grid, 192, smem, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "narrow_wgrad launch");
  return 0;
}

template <int CIN>
static int narrow_wgrad_dispatch_n(const NarrowWParams& P, int grid, cudaStream_t s) {
  const int nt = P.N <= 16 ? 16 : (P.N <= 32 ? 32 : 64);
  if (nt == 16) return narrow_wgrad_launch<CIN, 16>(P, grid, s);
  if (nt == 32) return narrow_wgrad_launch<CIN, 32>(P, grid, s);
  if constexpr (CIN < 64) return narrow_wgrad_launch<CIN, 64>(P, grid, s);
  set_error("narrow_wgrad: unsupported channel counts");
  return UNETB200_E_INVALID;
}

int narrow_wgrad(const unetb200_gconv_t* d, const void* x, const void* gy, float* partials, int splits, cudaStream_t s) {
  if (!narrow_wgrad_supported(d, x, gy)) { set_error("narrow_wgrad: unsupported shape"); return UNETB200_E_INVALID; }
  NarrowWParams P;
  memset(&P, 0, sizeof(P));
  P.x = (const __nv_bfloat16*)x; P.gy = (const __nv_bfloat16*)gy; P.partials = partials;
  P.ld_in = d->ld_in; P.ld_out = d->ld_out;
  P.npix = (long long)d->B * d->Hm * d->Wm;
  P.H = d->Hm; P.W = d->Wm; P.N = d->N;
  for (int t = 0; t < 9; ++t) {
    P.tap_dy[t] = d->tap_dy[t]; P.tap_dx[t] = d->tap_dx[t];
    P.tap_off[t] = ((long long)d->tap_dy[t] * d->Wm + d->tap_dx[t]) * d->ld_in;
  }
  const int grid = narrow_wgrad_grid(d, &P.ntiles);
  if (grid != splits) { set_error("narrow_wgrad: the planned split count is %d, got %d", grid, splits); return UNETB200_E_INVALID; }
  switch (d->Cin) {
    case 1: return narrow_wgrad_dispatch_n<1>(P, grid, s);
    case 2: return narrow_wgrad_dispatch_n<2>(P, grid, s);
    case 3: return narrow_wgrad_dispatch_n<3>(P, grid, s);
    case 4: return narrow_wgrad_dispatch_n<4>(P, grid, s);
    case 8: return narrow_wgrad_dispatch_n<8>(P, grid, s);
    case 16: return narrow_wgrad_dispatch_n<16>(P, grid, s);
    case 32: return narrow_wgrad_dispatch_n<32>(P, grid, s);
    default: return narrow_wgrad_dispatch_n<64>(P, grid, s);
  }
}

}  // namespace ub
