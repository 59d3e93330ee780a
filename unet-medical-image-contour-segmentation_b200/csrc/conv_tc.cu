// tcgen05 / TMEM / TMA implicit-GEMM engine (sm_100a) for the generalised convolution of unetb200.h.
//
//   fprop-like  D[m][n] = sum_{t,c} A[m][(t,c)] Wp[n][(t,c)]      (conv3x3 fprop + dgrad, convT fprop + dgrad)
//     M tile = TH x TW = 128 pixels of one image, N tile = BLOCK_N channels, K chunk = one tap x 128 bytes
//     of channels.  A chunk = ONE 4-D TMA box {128 B of channels, TW, TH, 1} at the tap-shifted pixel
//     coordinates: out-of-bounds rows/cols are zero-filled by TMA, which *is* the conv padding.  The
//     box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle, i.e. exactly the
//     K-major SWIZZLE_128B operand layout of tcgen05.mma.  Weights are a 2-D TMA box of the packed
//     [N][K] matrix.  One elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=32 bytes) with the
//     fp32 accumulator in TMEM; 4 epilogue warps read it back with tcgen05.ld, add the bias, round to
//     the storage type, stage the tile in (swizzled) shared memory, TMA-store it to NHWC (clipped at
//     the tensor edge), and reduce the BatchNorm sum / sum-of-squares of the *rounded* values.
//
//   wgrad-like  dWp[(t,c)][n] = sum_m A[m][(t,c)] G[m][n]
//     both operands are MN-major (the reduction runs over pixels): the same TMA boxes, now 64 pixels
//     per stage, described to tcgen05.mma as MN-major SWIZZLE_128B.  M tile = two (tap, channel
//     chunk) sub-tiles, N tile = BLOCK_N output channels, split over pixel tiles across CTAs.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp % 4).  Every mbarrier wait is bounded: a stuck
// pipeline traps instead of hanging the GPU.
#include <mutex>

#include "tc_common.cuh"

namespace ub {

// ------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------
struct alignas(64) TcParams {
  CUtensorMap a_map[4];     // activation source views (index = tap when in_scale == 2, else 0)
  CUtensorMap b_map;        // fprop: packed weights [N][K];  wgrad: unused
  CUtensorMap o_map[4];     // destination views, one per quadrant
  int tap_dy[9], tap_dx[9];
  int ntaps, cchunks, Cin, in_scale;
  int tiles_w, tiles_h, TW, TH, tw_shift;
  int Hm, Wm, Cq, N, K;
  const float* bias;
  float* stats_ws;          // per-tile BN partials [tile][2][Cq] (nullptr: no statistics)
  float* partials;          // wgrad
  int nsub;                 // wgrad: number of (tap, chunk) sub-tiles
  int ptiles, ptiles_per_split;
};

constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;            // one A stage (fprop): 128 rows x 128 B

// ------------------------------------------------------------------------------------------
// fprop
// ------------------------------------------------------------------------------------------
template <typename T, int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(192) tc_fprop_kernel(const __grid_constant__ TcParams p) {
  constexpr bool TF32 = sizeof(T) == 4;
  constexpr int EPR = 128 / sizeof(T);                 // elements per 128-byte row
  constexpr int kBBytes = BLOCK_N * 128;
  constexpr int kStage = kABytes + kBBytes;
  constexpr int NSUB = BLOCK_N / EPR;                  // output staging sub-tiles (128 rows x 128 B each)
  static_assert(NSUB * kABytes <= STAGES * kStage, "staging must fit in the pipeline buffers");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStage);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bx = blockIdx.x;
  const int tj = bx % p.tiles_w;
  bx /= p.tiles_w;
  const int ti = bx % p.tiles_h;
  const int b = bx / p.tiles_h;
  const int i0 = ti * p.TH, j0 = tj * p.TW;
  const int n0 = blockIdx.y * BLOCK_N;
  const int q = n0 / p.Cq, co0 = n0 - q * p.Cq;
  const int num_k = p.ntaps * p.cchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]);
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&p.o_map[q]);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], kStage);
        const int t = kb / p.cchunks, cc = kb - t * p.cchunks;
        uint8_t* sa = smem + s * kStage;
        if (p.in_scale == 1)
          tma_load_4d(sa, &p.a_map[0], &full_bar[s], cc * EPR, j0 + p.tap_dx[t], i0 + p.tap_dy[t], b);
        else
          tma_load_4d(sa, &p.a_map[t], &full_bar[s], cc * EPR, j0, i0, b);
        tma_load_2d(sa + kABytes, &p.b_map, &full_bar[s], t * p.Cin + cc * EPR, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TF32, false, false, kTileM, BLOCK_N);
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * kStage);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint64_t da = make_desc(sa + k * 32, 16, 1024);
          uint64_t db = make_desc(sa + kABytes + k * 32, 16, 1024);
          umma<TF32>(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);       // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full);             // accumulator complete
    }
  } else {
    // ---------------- epilogue: warps 2..5, TMEM lane quadrant = warp % 4 ----------------
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                 // tile row = pixel (row / TW, row % TW)
    const int et = threadIdx.x - 64;                  // 0..127
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    uint8_t* stage_out = smem;                        // pipeline buffers are free now
    float* sstat = reinterpret_cast<float*>(smem + STAGES * kStage + 256);   // [2][BLOCK_N]
    if (p.stats_ws)
      for (int i = et; i < 2 * BLOCK_N; i += 128) sstat[i] = 0.f;          // ordered by the barrier below
#pragma unroll 1
    for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + ch * 32, v);
      float f[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        f[e] = __uint_as_float(v[e]);
        if (p.bias) f[e] += Elem<T>::round(__ldg(p.bias + co0 + ch * 32 + e));
      }
      if constexpr (TF32) {
        // 32 fp32 = one full 128-byte row of sub-tile `ch`
        uint8_t* dst = stage_out + ch * kABytes + row * 128;
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16)
          *reinterpret_cast<float4*>(dst + ((c16 ^ (row & 7)) << 4)) =
              make_float4(f[4 * c16], f[4 * c16 + 1], f[4 * c16 + 2], f[4 * c16 + 3]);
      } else {
        // 32 bf16 = half a row (4 x 16 B) of sub-tile ch/2
        uint8_t* dst = stage_out + (ch >> 1) * kABytes + row * 128;
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          uint4 r;
          r.x = pack_bf16x2(f[8 * c16 + 0], f[8 * c16 + 1]);
          r.y = pack_bf16x2(f[8 * c16 + 2], f[8 * c16 + 3]);
          r.z = pack_bf16x2(f[8 * c16 + 4], f[8 * c16 + 5]);
          r.w = pack_bf16x2(f[8 * c16 + 6], f[8 * c16 + 7]);
          const int chunk = (ch & 1) * 4 + c16;
          *reinterpret_cast<uint4*>(dst + ((chunk ^ (row & 7)) << 4)) = r;
        }
      }
    }
    tc_fence_before();
    fence_async_smem();
    epi_bar_sync();
    if (et == 0) {
#pragma unroll 1
      for (int sub = 0; sub < NSUB; ++sub)
        tma_store_4d(&p.o_map[q], stage_out + sub * kABytes, co0 + sub * EPR, j0, i0, b);
      tma_store_commit();
    }
    if (p.stats_ws) {
      // per-channel sum / sum of squares of the rounded tile, rows outside the M grid masked out.
      // words of a sub-tile row: bf16 -> 32 words of 2 channels; fp32 -> 32 words of 1 channel.
      // Row splits are combined in shared memory; the tile's partial goes to the workspace with plain
      // stores (no global atomics: 16M fp64 atomics on 128 addresses cost more than the MMAs).
      constexpr int WORDS = NSUB * 32;                 // 32-bit words per tile row (<= 128)
      constexpr int RSPLIT = 128 / WORDS;
      const int word = et % WORDS, rs = et / WORDS;
      const int sub = word >> 5, w32 = word & 31;
      const int r0 = rs * (128 / RSPLIT), r1 = r0 + 128 / RSPLIT;
      float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
      for (int r = r0; r < r1; ++r) {
        const int pi = i0 + (r >> p.tw_shift), pj = j0 + (r & (p.TW - 1));
        if (pi >= p.Hm || pj >= p.Wm) continue;
        const uint32_t u = *reinterpret_cast<const uint32_t*>(stage_out + sub * kABytes + r * 128 +
                                                              ((((w32 >> 2) ^ (r & 7)) << 4) | ((w32 & 3) << 2)));
        if constexpr (TF32) {
          float a = __uint_as_float(u);
          s0 += a; q0 += a * a;
        } else {
          float a = __uint_as_float(u << 16), c = __uint_as_float(u & 0xffff0000u);
          s0 += a; q0 += a * a; s1 += c; q1 += c * c;
        }
      }
      const int lc = TF32 ? sub * 32 + w32 : sub * 64 + w32 * 2;      // channel within the N block
      if constexpr (RSPLIT == 1) {
        sstat[lc] = s0; sstat[BLOCK_N + lc] = q0;
        if constexpr (!TF32) { sstat[lc + 1] = s1; sstat[BLOCK_N + lc + 1] = q1; }
      } else {
        atomicAdd(&sstat[lc], s0); atomicAdd(&sstat[BLOCK_N + lc], q0);
        if constexpr (!TF32) { atomicAdd(&sstat[lc + 1], s1); atomicAdd(&sstat[BLOCK_N + lc + 1], q1); }
      }
      epi_bar_sync();
      float* dst = p.stats_ws + (long long)blockIdx.x * 2 * p.Cq + co0;
      for (int i = et; i < 2 * BLOCK_N; i += 128) dst[(i / BLOCK_N) * p.Cq + (i % BLOCK_N)] = sstat[i];
    }
    if (et == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

// ------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------
constexpr int kWPix = 64;                        // pixels (reduction length) per stage
constexpr int kWSub = kWPix * 128;               // one 64-pixel x 128-byte sub-tile = 8 KB

template <typename T, int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(192) tc_wgrad_kernel(const __grid_constant__ TcParams p) {
  constexpr bool TF32 = sizeof(T) == 4;
  constexpr int EPR = 128 / sizeof(T);
  constexpr int MSUB = kTileM / EPR;                   // A sub-tiles per stage (2 bf16 / 4 fp32)
  constexpr int NSUBT = BLOCK_N / EPR;                 // G sub-tiles per stage
  constexpr int kStage = (MSUB + NSUBT) * kWSub;
  constexpr int UMMA_K = 32 / sizeof(T);               // pixels per MMA (16 / 8)
  constexpr int MMAS = kWPix / UMMA_K;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStage);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x;                           // M tile: sub-tiles mt*MSUB .. +MSUB
  const int n0 = blockIdx.y * BLOCK_N;
  const int q = n0 / p.Cq, co0 = n0 - q * p.Cq;
  const int split = blockIdx.z;
  const int pt_begin = split * p.ptiles_per_split;
  int pt_end = pt_begin + p.ptiles_per_split;
  if (pt_end > p.ptiles) pt_end = p.ptiles;
  const int num_k = pt_end - pt_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]);
    tma_prefetch_desc(&p.o_map[q]);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], kStage);
        int pt = pt_begin + kb;
        const int tj = pt % p.tiles_w;
        pt /= p.tiles_w;
        const int ti = pt % p.tiles_h;
        const int b = pt / p.tiles_h;
        const int i0 = ti * p.TH, j0 = tj * p.TW;
        uint8_t* sa = smem + s * kStage;
#pragma unroll
        for (int h = 0; h < MSUB; ++h) {
          int sidx = mt * MSUB + h;
          if (sidx >= p.nsub) sidx = p.nsub - 1;        // padding rows: valid data, results discarded
          const int t = sidx / p.cchunks, cc = sidx - t * p.cchunks;
          tma_load_4d(sa + h * kWSub, &p.a_map[0], &full_bar[s], cc * EPR, j0 + p.tap_dx[t], i0 + p.tap_dy[t], b);
        }
#pragma unroll
        for (int h = 0; h < NSUBT; ++h)
          tma_load_4d(sa + (MSUB + h) * kWSub, &p.o_map[q], &full_bar[s], co0 + h * EPR, j0, i0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TF32, true, true, kTileM, BLOCK_N);
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * kStage);
#pragma unroll
        for (int k = 0; k < MMAS; ++k) {
          // MN-major: LBO = distance between 128-byte-wide sub-tiles, SBO = one swizzle group of
          // pixel rows (8 rows for 16-bit operands; 4 rows with the 32-byte-atom swizzle for tf32)
          constexpr uint32_t kLay = TF32 ? kLayoutSW128_32B : kLayoutSW128;
          constexpr uint32_t kSbo = TF32 ? 512 : 1024;
          uint64_t da = make_desc(sa + k * UMMA_K * 128, kWSub, kSbo, kLay);
          uint64_t db = make_desc(sa + MSUB * kWSub + k * UMMA_K * 128, kWSub, kSbo, kLay);
          umma<TF32>(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                  // D row = (sub-tile row / EPR, channel row % EPR)
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int sidx = mt * MSUB + row / EPR;
    const bool valid = sidx < p.nsub && num_k > 0;
    const long long k = (long long)sidx * EPR + (row % EPR);
    float* out = p.partials + (long long)split * p.K * p.N + k * p.N + n0;
#pragma unroll 1
    for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + ch * 32, v);
      if (valid) {
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(out + ch * 32 + e) =
              make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                          __uint_as_float(v[e + 3]));
      }
    }
    if (sidx < p.nsub && num_k <= 0) {                 // empty split: its partial must still be defined
      for (int e = 0; e < BLOCK_N; ++e) out[e] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// 4-D NHWC view {C, W, H, B} with element strides (sw, sh, sb) and a {128 B, TW, TH, 1} box
int encode_act_box(CUtensorMap* m, int dtype, const void* base, int C, int W, int H, int B, long long sw, long long sh,
                   long long sb, int TW, int TH, bool mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not found"); return UNETB200_E_CUDA; }
  const size_t esz = dtype == UNETB200_BF16 ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)(sw * esz), (cuuint64_t)(sh * esz), (cuuint64_t)(sb * esz)};
  cuuint32_t box[4] = {(cuuint32_t)(128 / esz), (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, dtype == UNETB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   // MN-major fp32 (tf32 wgrad) operands need the 32-byte-atom flavour of the 128-byte swizzle
                   (mn_major && dtype == UNETB200_F32) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                       : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(activation C=%d W=%d H=%d B=%d sw=%lld) failed: %d", C, W, H, B, sw, (int)r);
    return UNETB200_E_CUDA;
  }
  return 0;
}
static int encode_act(CUtensorMap* m, int dtype, const void* base, int C, int W, int H, int B, long long sw,
                      long long sh, long long sb, int TW, int TH, bool mn_major = false) {
  return encode_act_box(m, dtype, base, C, W, H, B, sw, sh, sb, TW, TH, mn_major);
}
int encode_weights(CUtensorMap* m, int dtype, const void* base, int K, int N, int box_n) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not found"); return UNETB200_E_CUDA; }
  const size_t esz = dtype == UNETB200_BF16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)(K * esz)};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dtype == UNETB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weights K=%d N=%d) failed: %d", K, N, (int)r);
    return UNETB200_E_CUDA;
  }
  return 0;
}

static bool tc_common_ok(const unetb200_gconv_t* d) {
  const int esz = d->dtype == UNETB200_BF16 ? 2 : 4;
  const int epr = 128 / esz;
  const int Cq = d->N / d->nquad;
  if (d->Cin % epr || Cq % 64) return false;
  if ((d->ld_in * esz) % 16 || (d->ld_out * esz) % 16) return false;
  if (d->in_scale == 1 && (d->in_off_y || d->in_off_x)) return false;
  if (d->in_scale == 2) {
    if (d->ntaps != 4) return false;
    for (int t = 0; t < 4; ++t)
      if (d->tap_dy[t] != (t >> 1) || d->tap_dx[t] != (t & 1)) return false;
  }
  return true;
}

int tc_fprop_supported(const unetb200_gconv_t* d, const void* x, const void* wp, const void* y) {
  if (d->dtype != UNETB200_BF16 && d->dtype != UNETB200_F32) return 0;
  if (d->dtype == UNETB200_F32 && d->algo != UNETB200_ALGO_TC && d->algo != UNETB200_ALGO_PREFER_TC) return 0;   // fp32 defaults to exact FMA
  if (!tc_common_ok(d)) return 0;
  if (!aligned16(x) || !aligned16(wp) || !aligned16(y)) return 0;
  return 1;
}

int tc_wgrad_supported(const unetb200_gconv_t* d, const void* x, const void* gy) {
  if (d->dtype != UNETB200_BF16 && d->dtype != UNETB200_F32) return 0;
  if (d->dtype == UNETB200_F32 && d->algo != UNETB200_ALGO_TC && d->algo != UNETB200_ALGO_PREFER_TC) return 0;
  if (!tc_common_ok(d) || d->in_scale != 1) return 0;
  if ((x && !aligned16(x)) || (gy && !aligned16(gy))) return 0;
  return 1;
}

static int pick_block_n(int Cq, int dtype) {
  const int cap = dtype == UNETB200_BF16 ? 256 : 128;
  if (cap >= 256 && Cq % 256 == 0) return 256;
  if (Cq % 128 == 0) return 128;
  return 64;
}

static void tile_shape(int Hm, int Wm, int pixels, int* TW, int* TH) {
  // power-of-two tile width <= 16 (8 for the 64-pixel wgrad tile), as wide as the grid allows
  int tw = pixels == 128 ? 16 : 8;
  while (tw > 1 && tw / 2 >= Wm) tw >>= 1;
  *TW = tw;
  *TH = pixels / tw;
  (void)Hm;
}

static int fill_maps(const unetb200_gconv_t* d, const void* x, const void* y, TcParams* P, int TW, int TH,
                     bool mn_major = false) {
  const size_t esz = d->dtype == UNETB200_BF16 ? 2 : 4;
  const int Cq = d->N / d->nquad;
  int rc;
  if (d->in_scale == 1) {
    // coordinates carry in_off + tap; the view is the whole source grid so that padding == TMA zero fill
    const char* base = (const char*)x + ((long long)d->in_off_y * d->Win + d->in_off_x) * d->ld_in * (long long)esz;
    rc = encode_act(&P->a_map[0], d->dtype, base, d->Cin, d->Win - d->in_off_x, d->Hin - d->in_off_y, d->B, d->ld_in,
                    (long long)d->Win * d->ld_in, (long long)d->Hin * d->Win * d->ld_in, TW, TH, mn_major);
    if (rc) return rc;
  } else {
    for (int t = 0; t < 4; ++t) {
      const int a = t >> 1, c = t & 1;
      const int oy = d->in_off_y + a, ox = d->in_off_x + c;
      const int Hq = (d->Hin - oy + 1) / 2, Wq = (d->Win - ox + 1) / 2;
      if (Hq <= 0 || Wq <= 0) { set_error("tc: empty quadrant view"); return UNETB200_E_INVALID; }
      const char* base = (const char*)x + ((long long)oy * d->Win + ox) * d->ld_in * (long long)esz;
      rc = encode_act(&P->a_map[t], d->dtype, base, d->Cin, Wq, Hq, d->B, 2 * d->ld_in,
                      2LL * d->Win * d->ld_in, (long long)d->Hin * d->Win * d->ld_in, TW, TH, mn_major);
      if (rc) return rc;
    }
  }
  for (int qd = 0; qd < d->nquad; ++qd) {
    const int a = qd >> 1, c = qd & 1;
    const int oy = d->out_off_y + (d->out_scale == 2 ? a : 0), ox = d->out_off_x + (d->out_scale == 2 ? c : 0);
    const char* base = (const char*)y + ((long long)oy * d->Wout + ox) * d->ld_out * (long long)esz;
    rc = encode_act(&P->o_map[qd], d->dtype, base, Cq, d->Wm, d->Hm, d->B, (long long)d->out_scale * d->ld_out,
                    (long long)d->out_scale * d->Wout * d->ld_out, (long long)d->Hout * d->Wout * d->ld_out, TW, TH,
                    mn_major);
    if (rc) return rc;
  }
  return 0;
}

static void fill_common(const unetb200_gconv_t* d, const GconvDev& g, TcParams* P, int TW, int TH) {
  const int esz = d->dtype == UNETB200_BF16 ? 2 : 4;
  for (int t = 0; t < 9; ++t) { P->tap_dy[t] = d->tap_dy[t]; P->tap_dx[t] = d->tap_dx[t]; }
  P->ntaps = d->ntaps;
  P->Cin = d->Cin;
  P->cchunks = d->Cin / (128 / esz);
  P->in_scale = d->in_scale;
  P->TW = TW; P->TH = TH;
  int sh = 0;
  while ((1 << sh) < TW) ++sh;
  P->tw_shift = sh;
  P->tiles_w = (d->Wm + TW - 1) / TW;
  P->tiles_h = (d->Hm + TH - 1) / TH;
  P->Hm = d->Hm; P->Wm = d->Wm; P->Cq = g.Cq; P->N = d->N; P->K = g.K;
  P->bias = nullptr; P->stats_ws = nullptr; P->partials = nullptr;
  P->nsub = d->ntaps * P->cchunks;
  P->ptiles = d->B * P->tiles_w * P->tiles_h;
  P->ptiles_per_split = P->ptiles;
}

template <typename T, int BN, int ST>
static int launch_fprop(const TcParams& P, dim3 grid, cudaStream_t s) {
  constexpr int smem = ST * (kABytes + BN * 128) + 1024 + 256 + 2 * BN * 4;   // + barriers + BN-stat scratch
  static bool configured = false;   // benign race: attribute set is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_fprop_kernel<T, BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail(e, "tc_fprop smem attribute");
    configured = true;
  }
  tc_fprop_kernel<T, BN, ST><<<grid, 192, smem, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "tc_fprop launch");
  return 0;
}

long long tc_fprop_tiles(const unetb200_gconv_t* d) {
  int TW, TH;
  tile_shape(d->Hm, d->Wm, 128, &TW, &TH);
  return (long long)d->B * ((d->Wm + TW - 1) / TW) * ((d->Hm + TH - 1) / TH);
}

template <typename T, int BN, int ST>
static int launch_fprop(const TcParams& P, dim3 grid, cudaStream_t s);

int tc_fprop(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* wp, const float* bias, void* y,
             double* stats, float* stats_ws, cudaStream_t stream) {
  TcParams P;
  int TW, TH;
  tile_shape(d->Hm, d->Wm, 128, &TW, &TH);
  fill_common(d, g, &P, TW, TH);
  int rc = fill_maps(d, x, y, &P, TW, TH);
  if (rc) return rc;
  const int BN = pick_block_n(g.Cq, d->dtype);
  rc = encode_weights(&P.b_map, d->dtype, wp, g.K, d->N, BN);
  if (rc) return rc;
  P.bias = bias;
  P.stats_ws = stats ? stats_ws : nullptr;
  dim3 grid((unsigned)P.ptiles, (unsigned)(d->N / BN));
  if (d->dtype == UNETB200_BF16) {
    if (BN == 256) rc = launch_fprop<__nv_bfloat16, 256, 2>(P, grid, stream);
    else if (BN == 128) rc = launch_fprop<__nv_bfloat16, 128, 3>(P, grid, stream);
    else rc = launch_fprop<__nv_bfloat16, 64, 4>(P, grid, stream);
  } else {
    if (BN == 128) rc = launch_fprop<float, 128, 3>(P, grid, stream);
    else rc = launch_fprop<float, 64, 4>(P, grid, stream);
  }
  if (rc) return rc;
  if (stats) return launch_stats_reduce(stats_ws, P.ptiles, 2 * g.Cq, stats, stream);
  return 0;
}

static int wgrad_geometry(const unetb200_gconv_t* d, int* mtiles, int* ntiles, int* ptiles, int* BN) {
  const int esz = d->dtype == UNETB200_BF16 ? 2 : 4;
  const int epr = 128 / esz;
  const int nsub = d->ntaps * (d->Cin / epr);
  const int msub = kTileM / epr;
  *mtiles = (nsub + msub - 1) / msub;
  *BN = pick_block_n(d->N / d->nquad, d->dtype);
  *ntiles = d->N / *BN;
  int TW, TH;
  tile_shape(d->Hm, d->Wm, kWPix, &TW, &TH);
  *ptiles = d->B * ((d->Wm + TW - 1) / TW) * ((d->Hm + TH - 1) / TH);
  return 0;
}

int tc_wgrad_splits(const unetb200_gconv_t* d, const GconvDev& g) {
  (void)g;
  int mt, nt, pt, BN;
  wgrad_geometry(d, &mt, &nt, &pt, &BN);
  long long tiles = (long long)mt * nt;
  long long want = ((long long)sm_count() * 4 + tiles - 1) / tiles;
  long long max_by_k = (pt + 15) / 16;          // at least 16 pixel tiles (1024 pixels) per split
  if (want > max_by_k) want = max_by_k;
  if (want > 512) want = 512;
  if (want < 1) want = 1;
  return (int)want;
}

template <typename T, int BN, int ST>
static int launch_wgrad(const TcParams& P, dim3 grid, cudaStream_t s) {
  constexpr int EPR = 128 / sizeof(T);
  constexpr int smem = ST * ((kTileM / EPR + BN / EPR) * kWSub) + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel<T, BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail(e, "tc_wgrad smem attribute");
    configured = true;
  }
  tc_wgrad_kernel<T, BN, ST><<<grid, 192, smem, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "tc_wgrad launch");
  return 0;
}

int tc_wgrad(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* gy, float* partials, int splits,
             cudaStream_t stream) {
  TcParams P;
  int TW, TH;
  tile_shape(d->Hm, d->Wm, kWPix, &TW, &TH);
  fill_common(d, g, &P, TW, TH);
  int rc = fill_maps(d, x, gy, &P, TW, TH, /*mn_major=*/true);
  if (rc) return rc;
  int mt, nt, pt, BN;
  wgrad_geometry(d, &mt, &nt, &pt, &BN);
  P.partials = partials;
  P.ptiles_per_split = (P.ptiles + splits - 1) / splits;
  dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)splits);
  if (d->dtype == UNETB200_BF16) {
    if (BN == 256) return launch_wgrad<__nv_bfloat16, 256, 2>(P, grid, stream);
    if (BN == 128) return launch_wgrad<__nv_bfloat16, 128, 3>(P, grid, stream);
    return launch_wgrad<__nv_bfloat16, 64, 4>(P, grid, stream);
  }
  if (BN == 128) return launch_wgrad<float, 128, 3>(P, grid, stream);
  return launch_wgrad<float, 64, 4>(P, grid, stream);
}

}  // namespace ub
