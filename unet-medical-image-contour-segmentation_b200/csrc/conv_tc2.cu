// Persistent tcgen05 implicit-GEMM engine for the fprop-like generalised convolution (conv3x3 fprop and
// dgrad, ConvTranspose fprop and dgrad) -- the second-generation kernel, built around what bounded the
// first one (conv_tc.cu): shared-memory *fill* bandwidth from L2 and per-tile set-up cost.
//
//  * one CTA per SM, looping over tiles (n-block-major so the weights of an n-block stay hot);
//  * M tile = 256 pixels = two 128-row accumulators that share every weight tile (halves weight traffic);
//  * vertical tap reuse: for each column shift dx ONE TMA box {128 B of channels, box_w, tile_h + 2, 1} is
//    loaded and serves the three row shifts dy of a 3x3 filter -- the UMMA descriptor of tap dy simply
//    starts (dy - dy_min) * box_w * 128 B further into the box (always 1024-B aligned, so no swizzle
//    phase is involved).  Activation fill traffic drops 2.7x, padding is still TMA zero fill;
//  * separate activation and weight rings (mbarrier full/empty each);
//  * accumulators double-buffered in TMEM when 2 x 2 x BLOCK_N <= 512 columns, so the epilogue of tile
//    i overlaps the MMAs of tile i+1;
//  * epilogue per warp: tcgen05.ld -> +bias -> round -> private 4 KB swizzled staging -> TMA store of an
//    {128 B, 8, 4} box; BatchNorm partial sums are kept in registers across tiles and flushed once per
//    (CTA, n-block) with plain stores.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
#include <cstring>
#include <mutex>

#include "tc_common.cuh"

namespace ub {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct alignas(64) Tc2Params {
  CUtensorMap a_map[4];      // source view per tap group, box {128 B, box_w, box_h(g), 1}
  CUtensorMap b_map;         // packed weights [N][K], box {128 B, BLOCK_N}
  CUtensorMap o_map[4];      // destination view per quadrant, box {128 B, 8, 4, 1}
  int ngroups;
  int g_dx[4], g_dy0[4], g_ntaps[4], g_tap[4][3];
  uint32_t g_box_bytes[4];
  int cchunks, Cin;
  int tiles_w, tiles_h, tile_w, tile_h;
  int sub_di[2], sub_dj[2];
  uint32_t sub_off[2];       // byte offset of the sub-tile's first row inside the box
  uint32_t row_bytes;        // box_w * 128: bytes per image row of the box = SBO = bytes per dy step
  int m_tiles, total_tiles;
  int Hm, Wm, Cq;
  const float* bias;
  float* stats_ws;           // [n_block][cta][8 epilogue warps][2][BLOCK_N]
};

constexpr uint32_t kAStage = 36864;     // max box: 18 rows x 16 px (or 34 x 8) x 128 B
constexpr int kEpiStage = 4096;         // 32 rows x 128 B per epilogue warp
constexpr int kEpiWarps = 8;            // two per TMEM lane quadrant: one per 128-row sub-tile
constexpr int kTc2Threads = 64 + 32 * kEpiWarps;

template <typename T, int BLOCK_N, int SA, int SB, int ACC>
__global__ void __launch_bounds__(kTc2Threads, 1) tc2_fprop_kernel(const __grid_constant__ Tc2Params p) {
  constexpr bool TF32 = sizeof(T) == 4;
  constexpr int EPR = 128 / sizeof(T);
  constexpr uint32_t kBStage = BLOCK_N * 128;
  constexpr int NCB = BLOCK_N / EPR;                   // 128-byte channel blocks per accumulator row
  static_assert(2 * BLOCK_N * ACC <= 512, "TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  uint8_t* b_ring = a_ring + SA * kAStage;
  uint8_t* epi = b_ring + SB * kBStage;                // 8 warps x 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + kEpiWarps * kEpiStage);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + SA;
  uint64_t* b_full = a_empty + SA;
  uint64_t* b_empty = b_full + SB;
  uint64_t* t_full = b_empty + SB;
  uint64_t* t_empty = t_full + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.ngroups; ++g) tma_prefetch_desc(&p.a_map[g]);
    tma_prefetch_desc(&p.b_map);
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < ACC; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    // The whole warp runs the loop convergently and one elected lane issues: values stay in uniform
    // registers, so a TMA issue is a handful of instructions (a lane-0-only branch makes the compiler
    // wrap every UTMALDG / UTCHMMA in an elect + broadcast loop).
    uint32_t sa = 0, pa = 1, sb = 0, pb = 1;             // stage index, parity to wait for on *_empty
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int nb = t / p.m_tiles;
      int mt = t - nb * p.m_tiles;
      const int tj = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int ti = mt % p.tiles_h;
      const int b = mt / p.tiles_h;
      const int i0 = ti * p.tile_h, j0 = tj * p.tile_w, n0 = nb * BLOCK_N;
      for (int c = 0; c < p.cchunks; ++c) {
        for (int g = 0; g < p.ngroups; ++g) {
          mbar_wait(&a_empty[sa], pa);
          if (elect_one()) {
            mbar_expect_tx(&a_full[sa], p.g_box_bytes[g]);
            tma_load_4d(a_ring + sa * kAStage, &p.a_map[g], &a_full[sa], c * EPR, j0 + p.g_dx[g], i0 + p.g_dy0[g], b);
          }
          if (++sa == SA) { sa = 0; pa ^= 1; }
          const int nt = p.g_ntaps[g];
          for (int k = 0; k < nt; ++k) {
            mbar_wait(&b_empty[sb], pb);
            if (elect_one()) {
              mbar_expect_tx(&b_full[sb], kBStage);
              tma_load_2d(b_ring + sb * kBStage, &p.b_map, &b_full[sb], p.g_tap[g][k] * p.Cin + c * EPR, n0);
            }
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ---------------------------------------
    // Convergent warp, one elected lane issues.  Descriptors differ only in their 14-bit start-address
    // field: one template per operand, plus byte offsets >> 4.
    constexpr uint32_t idesc = make_idesc(TF32, false, false, 128, BLOCK_N);
    const uint64_t a_desc_t = make_desc(smem_u32(a_ring), 16, p.row_bytes);
    const uint64_t b_desc_t = make_desc(smem_u32(b_ring), 16, 1024);
    const uint32_t sub1 = p.sub_off[1] >> 4, rowq = p.row_bytes >> 4;
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, pacc = 1;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait(&t_empty[acc], pacc);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 2 * BLOCK_N;
      uint32_t accum = 0;
      for (int c = 0; c < p.cchunks; ++c) {
        for (int g = 0; g < p.ngroups; ++g) {
          mbar_wait(&a_full[sa], pa);
          const uint64_t a_desc0 = a_desc_t + sa * (kAStage >> 4);
          const int nt = p.g_ntaps[g];
          for (int k = 0; k < nt; ++k) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint64_t b_desc0 = b_desc_t + sb * (kBStage >> 4);
            const uint64_t a_desc1 = a_desc0 + k * rowq;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                umma<TF32>(d_tmem, a_desc1 + 2 * kk, b_desc0 + 2 * kk, idesc, accum | (uint32_t)kk);
                umma<TF32>(d_tmem + BLOCK_N, a_desc1 + sub1 + 2 * kk, b_desc0 + 2 * kk, idesc, accum | (uint32_t)kk);
              }
              umma_commit(&b_empty[sb]);
              if (k == nt - 1) umma_commit(&a_empty[sa]);
            }
            __syncwarp();
            accum = 1;
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
      }
      if (elect_one()) umma_commit(&t_full[acc]);
      __syncwarp();
      if (++acc == ACC) { acc = 0; pacc ^= 1; }
    }
  } else {
    // ------------------------------------ epilogue -----------------------------------------
    // 8 warps: warp handles TMEM lane quadrant (warp % 4) of sub-tile (warp - 2) / 4.
    const int quad = warp & 3;
    const int s = (warp - 2) >> 2;
    const int ew = s * 4 + quad;                                      // 0..7
    uint8_t* buf = epi + ew * kEpiStage;
    uint32_t it = 0;
    int cur_nb = -1;
    float st[NCB][TF32 ? 2 : 4];
#pragma unroll
    for (int i = 0; i < NCB; ++i)
#pragma unroll
      for (int j = 0; j < (TF32 ? 2 : 4); ++j) st[i][j] = 0.f;
    auto flush = [&](int nb) {
      if (!p.stats_ws || nb < 0) return;
      float* dst = p.stats_ws + (((long long)nb * gridDim.x + blockIdx.x) * kEpiWarps + ew) * 2 * BLOCK_N;
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb) {
        if constexpr (TF32) {
          dst[cb * 32 + lane] = st[cb][0];
          dst[BLOCK_N + cb * 32 + lane] = st[cb][1];
          st[cb][0] = st[cb][1] = 0.f;
        } else {
          dst[cb * 64 + 2 * lane] = st[cb][0];
          dst[BLOCK_N + cb * 64 + 2 * lane] = st[cb][1];
          dst[cb * 64 + 2 * lane + 1] = st[cb][2];
          dst[BLOCK_N + cb * 64 + 2 * lane + 1] = st[cb][3];
          st[cb][0] = st[cb][1] = st[cb][2] = st[cb][3] = 0.f;
        }
      }
    };
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int nb = t / p.m_tiles;
      int mt = t - nb * p.m_tiles;
      const int tj = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int ti = mt % p.tiles_h;
      const int b = mt / p.tiles_h;
      const int n0 = nb * BLOCK_N;
      if (nb != cur_nb) { flush(cur_nb); cur_nb = nb; }
      const uint32_t acc = it % ACC;
      const int pi0 = ti * p.tile_h + p.sub_di[s] + 4 * quad;        // first image row of this warp's 4 x 8 patch
      const int pj0 = tj * p.tile_w + p.sub_dj[s];
      // rows of this warp: r = lane -> pixel (pi0 + r / 8, pj0 + r % 8); bit r of valid_rows: inside the M grid
      const uint32_t valid_rows =
          __ballot_sync(0xffffffffu, (pi0 + (lane >> 3) < p.Hm) && (pj0 + (lane & 7) < p.Wm));
      mbar_wait(&t_full[acc], (it / ACC) & 1);
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb) {                              // unrolled: st[cb] must stay in registers
        // an n-block may span output quadrants (ConvTranspose: n = q * Cq + co): one quadrant per 128-byte block
        const int n_cb = n0 + cb * EPR;
        const int q = n_cb / p.Cq, co_cb = n_cb - q * p.Cq;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 2 * BLOCK_N + s * BLOCK_N + cb * EPR;
        uint32_t v[EPR];
#pragma unroll
        for (int h = 0; h < EPR / 32; ++h) tmem_ld32_nowait(taddr + h * 32, v + h * 32);
        if (lane == 0) tma_store_wait_read0();                        // the previous store has finished reading `buf`
        tmem_ld_wait();
        __syncwarp();
        uint8_t* dst = buf + lane * 128;
#pragma unroll
        for (int h = 0; h < EPR / 32; ++h) {
          float f[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[h * 32 + e]);
          if (p.bias) {                                                 // one warp-uniform branch, not 32 predicated loads
            const float4* bp = reinterpret_cast<const float4*>(p.bias + co_cb + h * 32);
#pragma unroll
            for (int e4 = 0; e4 < 8; ++e4) {
              const float4 b4 = __ldg(bp + e4);
              f[4 * e4 + 0] += Elem<T>::round(b4.x); f[4 * e4 + 1] += Elem<T>::round(b4.y);
              f[4 * e4 + 2] += Elem<T>::round(b4.z); f[4 * e4 + 3] += Elem<T>::round(b4.w);
            }
          }
          if constexpr (TF32) {
#pragma unroll
            for (int c16 = 0; c16 < 8; ++c16)
              *reinterpret_cast<float4*>(dst + ((c16 ^ (lane & 7)) << 4)) =
                  make_float4(f[4 * c16], f[4 * c16 + 1], f[4 * c16 + 2], f[4 * c16 + 3]);
          } else {
#pragma unroll
            for (int c16 = 0; c16 < 4; ++c16) {
              uint4 r;
              r.x = pack_bf16x2(f[8 * c16 + 0], f[8 * c16 + 1]);
              r.y = pack_bf16x2(f[8 * c16 + 2], f[8 * c16 + 3]);
              r.z = pack_bf16x2(f[8 * c16 + 4], f[8 * c16 + 5]);
              r.w = pack_bf16x2(f[8 * c16 + 6], f[8 * c16 + 7]);
              *reinterpret_cast<uint4*>(dst + (((h * 4 + c16) ^ (lane & 7)) << 4)) = r;
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&p.o_map[q], buf, co_cb, pj0, pi0, b);
          tma_store_commit();
        }
        if (p.stats_ws) {
          // lane = 32-bit word of the 128-byte row: sum the rounded values over the 32 rows (branch-free,
          // all loads issued up front; rows outside the M grid are multiplied by 0)
          float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
          uint32_t u[32];
#pragma unroll
          for (int r = 0; r < 32; ++r)
            u[r] = *reinterpret_cast<const uint32_t*>(buf + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
          if (valid_rows == 0xffffffffu) {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              if constexpr (TF32) {
                const float a = __uint_as_float(u[r]);
                s0 += a; q0 = fmaf(a, a, q0);
              } else {
                const float a = __uint_as_float(u[r] << 16), c2 = __uint_as_float(u[r] & 0xffff0000u);
                s0 += a; q0 = fmaf(a, a, q0); s1 += c2; q1 = fmaf(c2, c2, q1);
              }
            }
          } else {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const float m = ((valid_rows >> r) & 1u) ? 1.f : 0.f;
              if constexpr (TF32) {
                const float a = __uint_as_float(u[r]) * m;
                s0 += a; q0 = fmaf(a, a, q0);
              } else {
                const float a = __uint_as_float(u[r] << 16) * m, c2 = __uint_as_float(u[r] & 0xffff0000u) * m;
                s0 += a; q0 = fmaf(a, a, q0); s1 += c2; q1 = fmaf(c2, c2, q1);
              }
            }
          }
          st[cb][0] += s0; st[cb][1] += q0;
          if constexpr (!TF32) { st[cb][2] += s1; st[cb][3] += q1; }
        }
      }
      // this warp has drained its 32 lanes of its sub-tile of accumulator set `acc`
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
    }
    flush(cur_nb);
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// stats[which*Cq + nb*BN + c] += sum over (cta, warp) rows of ws[nb][row][which][c]
__global__ void __launch_bounds__(1024) tc2_stats_reduce_kernel(const float* __restrict__ ws, int rows, int BN, int Cq, double* __restrict__ stats) {
  // one block per (32 columns, n-block): 32 row lanes x 32 columns, fixed assignment of rows to lanes and a fixed
  // order of the final sum -> bit-reproducible statistics (no atomics between blocks; `stats` is += by this block only)
  __shared__ double red[32][33];
  const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int nb = blockIdx.y;
  const int col = blockIdx.x * 32 + lane;              // 0 .. 2*BN
  double acc = 0.0;
  if (col < 2 * BN) {
    const float* base = ws + (long long)nb * rows * 2 * BN + col;
    int r = ry;
    for (; r + 96 < rows; r += 128) {
      const float a = base[(long long)r * 2 * BN], b = base[(long long)(r + 32) * 2 * BN],
                  c = base[(long long)(r + 64) * 2 * BN], d = base[(long long)(r + 96) * 2 * BN];
      acc += (double)a; acc += (double)b; acc += (double)c; acc += (double)d;
    }
    for (; r < rows; r += 32) acc += (double)base[(long long)r * 2 * BN];
  }
  red[ry][lane] = acc;
  __syncthreads();
  if (ry == 0 && col < 2 * BN) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += red[i][lane];
    const int which = col / BN, c = col - which * BN;
    stats[which * Cq + nb * BN + c] += s;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Tc2Plan {
  int BN, tile_w, tile_h, box_w, ngroups;
  int g_dx[4], g_dy0[4], g_ntaps[4], g_tap[4][3], g_map_src[4];
  int m_tiles, n_blocks, grid;
};

static bool tc2_plan(const unetb200_gconv_t* d, Tc2Plan* pl) {
  const int esz = d->dtype == UNETB200_BF16 ? 2 : 4;
  const int epr = 128 / esz;
  const int Cq = d->N / d->nquad;
  if (d->dtype != UNETB200_BF16 && d->dtype != UNETB200_F32) return false;
  if (d->Cin % epr || Cq % 64) return false;
  if ((d->ld_in * esz) % 16 || (d->ld_out * esz) % 16) return false;
  if (d->in_scale == 1 && (d->in_off_y || d->in_off_x)) return false;
  // tap groups: same dx, consecutive dy (at most 3 per group, 4 groups)
  pl->ngroups = 0;
  if (d->in_scale == 2) {
    if (d->ntaps != 4) return false;
    for (int t = 0; t < 4; ++t) {
      if (d->tap_dy[t] != (t >> 1) || d->tap_dx[t] != (t & 1)) return false;
      pl->g_dx[t] = 0; pl->g_dy0[t] = 0; pl->g_ntaps[t] = 1; pl->g_tap[t][0] = t; pl->g_map_src[t] = t;
    }
    pl->ngroups = 4;
  } else {
    bool used[9] = {false};
    for (int t = 0; t < d->ntaps; ++t) {
      if (used[t]) continue;
      if (pl->ngroups == 4) return false;
      const int g = pl->ngroups++;
      pl->g_dx[g] = d->tap_dx[t];
      int lo = d->tap_dy[t];
      for (int u = 0; u < d->ntaps; ++u)
        if (d->tap_dx[u] == pl->g_dx[g] && d->tap_dy[u] < lo) lo = d->tap_dy[u];
      pl->g_dy0[g] = lo;
      pl->g_ntaps[g] = 0;
      pl->g_map_src[g] = 0;
      for (int k = 0; k < 3; ++k) {                    // taps at dy = lo, lo+1, lo+2 (must be gap-free)
        int found = -1;
        for (int u = 0; u < d->ntaps; ++u)
          if (!used[u] && d->tap_dx[u] == pl->g_dx[g] && d->tap_dy[u] == lo + k) { found = u; break; }
        if (found < 0) break;
        used[found] = true;
        pl->g_tap[g][pl->g_ntaps[g]++] = found;
      }
    }
    for (int t = 0; t < d->ntaps; ++t)
      if (!used[t]) return false;
  }
  // N block: 256 (one accumulator set: the epilogue does not overlap the next tile) or 128 (two sets).  Measured at
  // C2 (profiles/r2_bench_c2_c*.json): the ConvTranspose fprop (1 tap, 4 quadrants: epilogue heavy) runs 15 % faster
  // with 128, its dgrad (4 taps, K = 4 C_out) 14 % slower; UNETB200_TC2_MAXBN overrides both for A/B runs
  static const int env_bn = getenv("UNETB200_TC2_MAXBN") ? atoi(getenv("UNETB200_TC2_MAXBN")) : 0;
  const int max_bn = env_bn ? env_bn : (d->nquad == 4 ? 128 : 256);
  // the N block may span quadrants (the epilogue resolves the quadrant per 128-byte channel block), so the block
  // width follows N = nquad * Cq: a 64-channel ConvTranspose reads its input once (N = 256) instead of 4 times
  const int epr_n = 128 / esz;
  const int Nall = (Cq % epr_n == 0) ? d->N : Cq;
  const bool wide = d->dtype == UNETB200_BF16 && max_bn >= 256 && (Cq % 256 == 0 || (Cq == 64 && Nall == 256));
  pl->BN = wide ? 256 : (Cq % 128 == 0 ? 128 : (Nall % 128 == 0 && Cq == 64 && d->nquad == 4 ? 128 : 64));
  if (d->Wm > 8) { pl->tile_w = 16; pl->tile_h = 16; pl->box_w = 16; }
  else { pl->tile_w = 8; pl->tile_h = 32; pl->box_w = 8; }
  const int tiles_w = (d->Wm + pl->tile_w - 1) / pl->tile_w, tiles_h = (d->Hm + pl->tile_h - 1) / pl->tile_h;
  pl->m_tiles = d->B * tiles_w * tiles_h;
  pl->n_blocks = d->N / pl->BN;
  long long total = (long long)pl->m_tiles * pl->n_blocks;
  int sms = sm_count();
  pl->grid = (int)(total < sms ? total : sms);
  return true;
}

int tc2_fprop_supported(const unetb200_gconv_t* d, const void* x, const void* wp, const void* y) {
  if (d->dtype == UNETB200_F32 && d->algo != UNETB200_ALGO_TC && d->algo != UNETB200_ALGO_PREFER_TC) return 0;
  Tc2Plan pl;
  if (!tc2_plan(d, &pl)) return 0;
  if (!aligned16(x) || !aligned16(wp) || !aligned16(y)) return 0;
  return 1;
}

long long tc2_stats_workspace(const unetb200_gconv_t* d) {
  Tc2Plan pl;
  if (!tc2_plan(d, &pl)) return 0;
  return (long long)pl.n_blocks * pl.grid * kEpiWarps * 2 * pl.BN;
}

template <typename T, int BN, int SA, int SB, int ACC>
static int tc2_launch(const Tc2Params& P, int grid, cudaStream_t s) {
  constexpr int smem = SA * kAStage + SB * BN * 128 + kEpiWarps * kEpiStage + 1024 + 256;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(&tc2_fprop_kernel<T, BN, SA, SB, ACC>), smem,
                                    "tc2_fprop smem attribute"))
    return rc;
  tc2_fprop_kernel<T, BN, SA, SB, ACC><<<grid, kTc2Threads, smem, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "tc2_fprop launch");
  return 0;
}

int tc2_fprop(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* wp, const float* bias, void* y,
              double* stats, float* stats_ws, cudaStream_t stream) {
  Tc2Plan pl;
  if (!tc2_plan(d, &pl)) { set_error("tc2_fprop: unsupported shape"); return UNETB200_E_INVALID; }
  const size_t esz = d->dtype == UNETB200_BF16 ? 2 : 4;
  const int epr = 128 / (int)esz;
  Tc2Params P;
  memset(&P, 0, sizeof(P));
  P.ngroups = pl.ngroups;
  int rc;
  for (int gi = 0; gi < pl.ngroups; ++gi) {
    P.g_dx[gi] = pl.g_dx[gi]; P.g_dy0[gi] = pl.g_dy0[gi]; P.g_ntaps[gi] = pl.g_ntaps[gi];
    for (int k = 0; k < 3; ++k) P.g_tap[gi][k] = pl.g_tap[gi][k];
    const int box_h = pl.tile_h + pl.g_ntaps[gi] - 1;
    P.g_box_bytes[gi] = (uint32_t)(box_h * pl.box_w * 128);
    if (d->in_scale == 1) {
      rc = encode_act_box(&P.a_map[gi], d->dtype, x, d->Cin, d->Win, d->Hin, d->B, d->ld_in,
                          (long long)d->Win * d->ld_in, (long long)d->Hin * d->Win * d->ld_in, pl.box_w, box_h, false);
    } else {
      const int a = gi >> 1, c = gi & 1;
      const int oy = d->in_off_y + a, ox = d->in_off_x + c;
      const int Hq = (d->Hin - oy + 1) / 2, Wq = (d->Win - ox + 1) / 2;
      if (Hq <= 0 || Wq <= 0) { set_error("tc2: empty quadrant view"); return UNETB200_E_INVALID; }
      const char* base = (const char*)x + ((long long)oy * d->Win + ox) * d->ld_in * (long long)esz;
      rc = encode_act_box(&P.a_map[gi], d->dtype, base, d->Cin, Wq, Hq, d->B, 2 * d->ld_in, 2LL * d->Win * d->ld_in,
                          (long long)d->Hin * d->Win * d->ld_in, pl.box_w, box_h, false);
    }
    if (rc) return rc;
  }
  for (int qd = 0; qd < d->nquad; ++qd) {
    const int a = qd >> 1, c = qd & 1;
    const int oy = d->out_off_y + (d->out_scale == 2 ? a : 0), ox = d->out_off_x + (d->out_scale == 2 ? c : 0);
    const char* base = (const char*)y + ((long long)oy * d->Wout + ox) * d->ld_out * (long long)esz;
    rc = encode_act_box(&P.o_map[qd], d->dtype, base, g.Cq, d->Wm, d->Hm, d->B, (long long)d->out_scale * d->ld_out,
                        (long long)d->out_scale * d->Wout * d->ld_out, (long long)d->Hout * d->Wout * d->ld_out, 8, 4,
                        false);
    if (rc) return rc;
  }
  rc = encode_weights(&P.b_map, d->dtype, wp, g.K, d->N, pl.BN);
  if (rc) return rc;
  P.cchunks = d->Cin / epr;
  P.Cin = d->Cin;
  P.tile_w = pl.tile_w; P.tile_h = pl.tile_h;
  P.tiles_w = (d->Wm + pl.tile_w - 1) / pl.tile_w;
  P.tiles_h = (d->Hm + pl.tile_h - 1) / pl.tile_h;
  P.row_bytes = (uint32_t)pl.box_w * 128;
  if (pl.box_w == 16) {             // two 8-wide sub-tiles side by side
    P.sub_di[0] = 0; P.sub_dj[0] = 0; P.sub_off[0] = 0;
    P.sub_di[1] = 0; P.sub_dj[1] = 8; P.sub_off[1] = 8 * 128;
  } else {                          // two 16-row sub-tiles stacked
    P.sub_di[0] = 0; P.sub_dj[0] = 0; P.sub_off[0] = 0;
    P.sub_di[1] = 16; P.sub_dj[1] = 0; P.sub_off[1] = 16 * 8 * 128;
  }
  P.m_tiles = pl.m_tiles;
  P.total_tiles = pl.m_tiles * pl.n_blocks;
  P.Hm = d->Hm; P.Wm = d->Wm; P.Cq = g.Cq;
  P.bias = bias;
  P.stats_ws = stats ? stats_ws : nullptr;
  if (stats) {
    cudaError_t e = cudaMemsetAsync(stats_ws, 0, sizeof(float) * (size_t)tc2_stats_workspace(d), stream);
    if (e != cudaSuccess) return cuda_fail(e, "tc2 stats workspace memset");
  }
  if (d->dtype == UNETB200_BF16) {
    if (pl.BN == 256) rc = tc2_launch<__nv_bfloat16, 256, 2, 3, 1>(P, pl.grid, stream);
    else if (pl.BN == 128) rc = tc2_launch<__nv_bfloat16, 128, 3, 4, 2>(P, pl.grid, stream);
    else rc = tc2_launch<__nv_bfloat16, 64, 3, 6, 2>(P, pl.grid, stream);
  } else {
    if (pl.BN == 128) rc = tc2_launch<float, 128, 3, 4, 2>(P, pl.grid, stream);
    else rc = tc2_launch<float, 64, 3, 6, 2>(P, pl.grid, stream);
  }
  if (rc) return rc;
  if (stats) {
    const int rows = pl.grid * kEpiWarps;
    tc2_stats_reduce_kernel<<<dim3((2 * pl.BN + 31) / 32, pl.n_blocks), 1024, 0, stream>>>(stats_ws, rows, pl.BN, g.Cq,
                                                                                         stats);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "tc2_stats_reduce");
  }
  return 0;
}


// ------------------------------------------------------------------------------------------
// wgrad, second generation: vertical tap reuse on the activation operand
// ------------------------------------------------------------------------------------------
//   dWp[(t,c)][n] = sum_pixels A[pixel][(t,c)] * G[pixel][n]      (both operands MN-major, K = pixels)
// K step = an 8x8 pixel tile.  A "unit" is (tap group g, 128-byte channel chunk c): ONE TMA box
// {128 B, 8, 8 + ndy - 1, 1} of x at column shift dx(g) serves the ndy row shifts of the group -- tap dy
// starts (dy - dy_min) * 1024 B into the box.  An M tile = MSUB consecutive units (128 rows) and owns ndy
// accumulators (one per dy) in TMEM, all fed by the same boxes and the same dY sub-tiles:
// fill traffic per FLOP is about half of the first-generation kernel's.
struct alignas(64) Tc2WParams {
  CUtensorMap a_map[4];      // x view per tap group, box {128 B, 8, 8 + ndy - 1, 1}
  CUtensorMap o_map[4];      // dY view per quadrant, box {128 B, 8, 8, 1}
  int ngroups, ndy;
  int g_dx[4], g_dy0[4], g_tap[4][3];
  int cchunks, Cin, nunits;
  int tiles_w, tiles_h;
  int ptiles, ptiles_per_split;
  int Cq, N, K;
  uint32_t a_box_bytes;      // (8 + ndy - 1) * 8 * 128, a multiple of 1024
  float* partials;
};

template <typename T, int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(192, 1) tc2_wgrad_kernel(const __grid_constant__ Tc2WParams p) {
  constexpr bool TF32 = sizeof(T) == 4;
  constexpr int EPR = 128 / sizeof(T);
  constexpr int MSUB = 128 / EPR;                      // units per M tile (2 bf16 / 4 tf32)
  constexpr int NSUBT = BLOCK_N / EPR;                 // dY sub-tiles per stage
  constexpr uint32_t kGSub = 64 * 128;                 // one 8x8-pixel x 128-byte dY sub-tile
  constexpr uint32_t kABox = 10 * 8 * 128;             // largest x box (ndy = 3)
  constexpr uint32_t kStage = MSUB * kABox + NSUBT * kGSub;
  constexpr int UMMA_K = 32 / sizeof(T);               // pixels per MMA (16 / 8)
  constexpr int MMAS = 64 / UMMA_K;
  constexpr uint32_t kLay = TF32 ? kLayoutSW128_32B : kLayoutSW128;
  constexpr uint32_t kSbo = TF32 ? 512 : 1024;
  constexpr uint32_t kTmemCols = 3 * BLOCK_N <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStage);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x;
  const int n0 = blockIdx.y * BLOCK_N;
  const int q = n0 / p.Cq, co0 = n0 - q * p.Cq;
  const int split = blockIdx.z;
  const int pt_begin = split * p.ptiles_per_split;
  int pt_end = pt_begin + p.ptiles_per_split;
  if (pt_end > p.ptiles) pt_end = p.ptiles;
  const int num_k = pt_end - pt_begin;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.ngroups; ++g) tma_prefetch_desc(&p.a_map[g]);
    tma_prefetch_desc(&p.o_map[q]);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // convergent warp, one elected lane issues (see tc2_fprop_kernel)
    const uint32_t stage_bytes = MSUB * p.a_box_bytes + NSUBT * kGSub;
    int pt = pt_begin;
    int tj = pt % p.tiles_w;
    int rest = pt / p.tiles_w;
    int ti = rest % p.tiles_h;
    int b = rest / p.tiles_h;
    // the MSUB units of this M tile are fixed for the whole kernel
    int uc[MSUB], ug[MSUB];
#pragma unroll
    for (int h = 0; h < MSUB; ++h) {
      int u = mt * MSUB + h;
      if (u >= p.nunits) u = p.nunits - 1;              // padding rows: valid data, results discarded
      uc[h] = u / p.ngroups;
      ug[h] = u - uc[h] * p.ngroups;
    }
    uint32_t s = 0, ph = 1;
    for (int kb = 0; kb < num_k; ++kb) {
      mbar_wait(&empty_bar[s], ph);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[s], stage_bytes);
        const int i0 = ti * 8, j0 = tj * 8;
        uint8_t* sa = smem + s * kStage;
#pragma unroll
        for (int h = 0; h < MSUB; ++h)
          tma_load_4d(sa + h * kABox, &p.a_map[ug[h]], &full_bar[s], uc[h] * EPR, j0 + p.g_dx[ug[h]], i0 + p.g_dy0[ug[h]], b);
#pragma unroll
        for (int h = 0; h < NSUBT; ++h)
          tma_load_4d(sa + MSUB * kABox + h * kGSub, &p.o_map[q], &full_bar[s], co0 + h * EPR, j0, i0, b);
      }
      if (++s == STAGES) { s = 0; ph ^= 1; }
      if (++tj == p.tiles_w) { tj = 0; if (++ti == p.tiles_h) { ti = 0; ++b; } }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(TF32, true, true, 128, BLOCK_N);
    // MN-major: LBO = distance between 128-byte-wide sub-tiles, SBO = one swizzle group of pixel rows
    const uint64_t da_t = make_desc(smem_u32(smem), kABox, kSbo, kLay);
    const uint64_t db_t = make_desc(smem_u32(smem) + MSUB * kABox, kGSub, kSbo, kLay);
    uint32_t s = 0, ph = 0;
    for (int kb = 0; kb < num_k; ++kb) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint64_t da0 = da_t + s * (kStage >> 4), db0 = db_t + s * (kStage >> 4);
      if (elect_one()) {
        for (int a = 0; a < p.ndy; ++a) {
#pragma unroll
          for (int k = 0; k < MMAS; ++k)
            umma<TF32>(tmem_base + a * BLOCK_N, da0 + ((a * 1024 + k * UMMA_K * 128) >> 4), db0 + ((k * UMMA_K * 128) >> 4),
                       idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(tmem_full);
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                  // D row = (unit row / EPR, channel row % EPR)
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int u = mt * MSUB + row / EPR;
    const bool live = u < p.nunits;
    const int c = live ? u / p.ngroups : 0, g = live ? u - c * p.ngroups : 0;
    for (int a = 0; a < p.ndy; ++a) {
      const long long k = (long long)p.g_tap[g][a] * p.Cin + c * EPR + (row % EPR);
      float* out = p.partials + (long long)split * p.K * p.N + k * p.N + n0;
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + a * BLOCK_N + ch * 32, v);
        if (live) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            float4 o4 = num_k > 0 ? make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(out + ch * 32 + e) = o4;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

struct Tc2WPlan {
  int BN, ngroups, ndy, nunits, mtiles, ntiles, ptiles, tiles_w, tiles_h;
  int g_dx[4], g_dy0[4], g_tap[4][3];
};

static bool tc2_wgrad_plan(const unetb200_gconv_t* d, Tc2WPlan* w) {
  if (d->in_scale != 1) return false;
  Tc2Plan pl;
  if (!tc2_plan(d, &pl)) return false;
  const int esz = d->dtype == UNETB200_BF16 ? 2 : 4;
  const int epr = 128 / esz;
  w->ngroups = pl.ngroups;
  w->ndy = pl.g_ntaps[0];
  for (int g = 0; g < pl.ngroups; ++g) {
    if (pl.g_ntaps[g] != w->ndy) return false;         // every group needs the same number of row shifts
    w->g_dx[g] = pl.g_dx[g]; w->g_dy0[g] = pl.g_dy0[g];
    for (int k = 0; k < 3; ++k) w->g_tap[g][k] = pl.g_tap[g][k];
  }
  const int Cq = d->N / d->nquad;
  w->BN = Cq % 128 == 0 ? 128 : 64;
  w->nunits = pl.ngroups * (d->Cin / epr);
  const int msub = 128 / epr;
  w->mtiles = (w->nunits + msub - 1) / msub;
  w->ntiles = d->N / w->BN;
  w->tiles_w = (d->Wm + 7) / 8;
  w->tiles_h = (d->Hm + 7) / 8;
  w->ptiles = d->B * w->tiles_w * w->tiles_h;
  return true;
}

int tc2_wgrad_supported(const unetb200_gconv_t* d, const void* x, const void* gy) {
  if (d->dtype == UNETB200_F32 && d->algo != UNETB200_ALGO_TC && d->algo != UNETB200_ALGO_PREFER_TC) return 0;
  Tc2WPlan w;
  if (!tc2_wgrad_plan(d, &w)) return 0;
  if ((x && !aligned16(x)) || (gy && !aligned16(gy))) return 0;
  return 1;
}

int tc2_wgrad_splits(const unetb200_gconv_t* d) {
  Tc2WPlan w;
  if (!tc2_wgrad_plan(d, &w)) return 1;
  const long long tiles = (long long)w.mtiles * w.ntiles;
  const long long max_by_k = (w.ptiles + 15) / 16;      // at least 16 pixel tiles (1024 pixels) per split
  return pick_splits(tiles, sm_count(), max_by_k);      // one CTA per SM: whole waves (297 CTAs would cost 3 waves)
}

template <typename T, int BN, int ST>
static int tc2_launch_wgrad(const Tc2WParams& P, dim3 grid, cudaStream_t s) {
  constexpr int EPR = 128 / sizeof(T);
  constexpr int smem = ST * ((128 / EPR) * 10240 + (BN / EPR) * 8192) + 1024 + 256;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(&tc2_wgrad_kernel<T, BN, ST>), smem,
                                    "tc2_wgrad smem attribute"))
    return rc;
  tc2_wgrad_kernel<T, BN, ST><<<grid, 192, smem, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "tc2_wgrad launch");
  return 0;
}

int tc2_wgrad(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* gy, float* partials, int splits,
              cudaStream_t stream) {
  Tc2WPlan w;
  if (!tc2_wgrad_plan(d, &w)) { set_error("tc2_wgrad: unsupported shape"); return UNETB200_E_INVALID; }
  const size_t esz = d->dtype == UNETB200_BF16 ? 2 : 4;
  const int epr = 128 / (int)esz;
  Tc2WParams P;
  memset(&P, 0, sizeof(P));
  P.ngroups = w.ngroups; P.ndy = w.ndy;
  int rc;
  for (int gi = 0; gi < w.ngroups; ++gi) {
    P.g_dx[gi] = w.g_dx[gi]; P.g_dy0[gi] = w.g_dy0[gi];
    for (int k = 0; k < 3; ++k) P.g_tap[gi][k] = w.g_tap[gi][k];
    rc = encode_act_box(&P.a_map[gi], d->dtype, x, d->Cin, d->Win, d->Hin, d->B, d->ld_in, (long long)d->Win * d->ld_in,
                        (long long)d->Hin * d->Win * d->ld_in, 8, 8 + w.ndy - 1, true);
    if (rc) return rc;
  }
  for (int qd = 0; qd < d->nquad; ++qd) {
    const int a = qd >> 1, c = qd & 1;
    const int oy = d->out_off_y + (d->out_scale == 2 ? a : 0), ox = d->out_off_x + (d->out_scale == 2 ? c : 0);
    const char* base = (const char*)gy + ((long long)oy * d->Wout + ox) * d->ld_out * (long long)esz;
    rc = encode_act_box(&P.o_map[qd], d->dtype, base, g.Cq, d->Wm, d->Hm, d->B, (long long)d->out_scale * d->ld_out,
                        (long long)d->out_scale * d->Wout * d->ld_out, (long long)d->Hout * d->Wout * d->ld_out, 8, 8,
                        true);
    if (rc) return rc;
  }
  P.cchunks = d->Cin / epr; P.Cin = d->Cin; P.nunits = w.nunits;
  P.tiles_w = w.tiles_w; P.tiles_h = w.tiles_h;
  P.ptiles = w.ptiles;
  P.ptiles_per_split = (w.ptiles + splits - 1) / splits;
  P.Cq = g.Cq; P.N = d->N; P.K = g.K;
  P.a_box_bytes = (uint32_t)((8 + w.ndy - 1) * 8 * 128);
  P.partials = partials;
  dim3 grid((unsigned)w.mtiles, (unsigned)w.ntiles, (unsigned)splits);
  if (d->dtype == UNETB200_BF16) {
    if (w.BN == 128) return tc2_launch_wgrad<__nv_bfloat16, 128, 5>(P, grid, stream);
    return tc2_launch_wgrad<__nv_bfloat16, 64, 6>(P, grid, stream);
  }
  if (w.BN == 128) return tc2_launch_wgrad<float, 128, 3>(P, grid, stream);
  return tc2_launch_wgrad<float, 64, 3>(P, grid, stream);
}

}  // namespace ub
