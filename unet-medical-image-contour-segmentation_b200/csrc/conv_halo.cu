// 3x3 convolutions with 16 / 32 / 64 channels on either side (not 64 -> 64) on the tensor cores, TMA staged: the
// full-resolution layers of the reference's light variants UNet_S / UNet_T / UNet_SA (unet_model.py:52-189; UNet_S is
// what train.py:253 builds).  fprop / dgrad (halo_conv_kernel) and wgrad (halo_wgrad_kernel) of
// nn.Conv2d(k=3, padding=1, bias=False) (unet_parts.py:15,18).
//
// conv_narrow.cu builds the im2col rows of these layers with threads (nine gathers per pixel) and runs 5-10x above
// the HBM floor, instruction- and latency-bound (profiles/r2_narrow_full.md).  Here the tensor core reads the pixels
// where TMA put them.  A pixel with C channels is a 2C-byte row; 2C = 32 / 64 / 128 bytes is exactly a
// SWIZZLE_32B / 64B / 128B row, and both the TMA write and the UMMA read swizzle as a function of the absolute
// shared-memory address (tools/umma_probe.cu checks (d), (e); profiles/r2_umma_probe_narrow.txt), so
//   * fprop: ONE halo box {C, 18, 18} per 16 x 16 pixel tile serves all nine taps: the K-major A descriptor of tap
//     (dy, dx) starts ((dy + 1) * 18 + dx + 1) pixels into the box, 8-pixel row groups one box row apart (SBO); the
//     tile is two M = 128 accumulators (8 x 16 pixel strips); 9 * C / 16 MMAs per strip, weights resident in shared
//     memory for the whole kernel;
//   * wgrad: dW[dy][dx] = sum_{y', x} X[y'][x + dx] (x) G[y' - dy][x] -- per image row ONE MMA with all nine taps:
//     A = X row, MN-major, its M atoms one PIXEL apart (LBO = 2C: the dx shifts), B = G rows, MN-major, its N atoms one
//     ROW apart (LBO = row pitch: the dy shifts), K = 16 pixels; all of dW stays in TMEM, one fp32 partial per CTA.
// No thread touches an operand; what is left per tile is the epilogue (TMEM -> bf16 -> BatchNorm statistics / folded
// BatchNorm + ReLU -> coalesced stores).
#include <cstring>

#include "halo_common.cuh"

namespace ub {

// ------------------------------------------------------------------------------------------ fprop / dgrad
struct HaloParams {
  CUtensorMap x_map;             // {Cin, W, H, B}, box {Cin, 18, 18, 1}
  const void* wp;                // packed [N][9 * Cin], tap order of the descriptor (bf16, or fp32 for the TF32 form)
  void* y;                       // [B][H][W][ld_out]
  const float* affine;           // MODE 1: scale[N] then shift[N]
  float* stats_ws;               // MODE 0: [grid * 8][2][N] or null
  long long ld_out;
  int H, W;
  int tiles_w, tiles_h, ntiles;
  int tap_pix[9];                // (dy + 1) * 18 + dx + 1
  int cin_real, n_real;          // 8-channel tensors ride the 16-channel instantiation: TMA zero-fills channels 8..15 of
                                 // a pixel row (the box is wider than the tensor), weight rows / columns 8..15 are zero
  // MODE 2 (bf16 dgrad whose output g is the gradient of z = relu(bn(yprev))): the raw conv output of the layer below and
  // its BatchNorm coefficients [4][N] = mean, invstd, scale, shift; stats_ws then receives sum(g * mask) and
  // sum(g * mask * xhat) per channel -- the reduction pass of unetb200_bn_relu_bwd_reduce
  const __nv_bfloat16* yprev;
  long long ld_y;
  const float* bnc;
  // MODE 1, bf16, optional: MaxPool2d(2) of the activation as a second output, [B][H/2][W/2][ld_pool] (a warp's 4 x 8
  // pixel patch holds whole 2 x 2 windows; conv_tc3.cu does the same for the wide layers)
  __nv_bfloat16* pooled;
  long long ld_pool;
};

constexpr int kHaloThreads = 320;
constexpr int kHaloBW = 18;

// ES = 2: bf16 operands (kind::f16, K = 16 per MMA); ES = 4: fp32 operands read as TF32 (kind::tf32, K = 8 per MMA --
// the same 32 bytes; a 16 / 32-channel fp32 pixel row is a 64 / 128-byte swizzle row), fp32 output.  Used without
// autocast when TF32 is allowed (UNET_B200_PRECISION=tf32, the default when torch allows TF32 in cuDNN).
template <int CIN, int NT, int ES = 2>
struct HaloCfg {
  static constexpr int P = ES * CIN;
  static constexpr uint32_t kBox = kHaloBW * kHaloBW * P;
  static constexpr uint32_t kStage = (kBox + 1023) / 1024 * 1024;
  static constexpr uint32_t kBTile = NT * P < 1024 ? 1024u : (uint32_t)(NT * P);
  static constexpr int RS = NT * ES;
  static constexpr uint32_t kEpi = 8 * 32 * RS;
  static constexpr int kFixed = 9 * (int)kBTile + (int)kEpi + 128 * 4 + 256 + 1024;
  static constexpr int STAGES = P == 32 ? 4 : 3;
  static constexpr int smem = STAGES * (int)kStage + kFixed;
  static constexpr int CTAS = smem <= 110 * 1024 ? 2 : 1;
  static constexpr int kTmemCols = 4 * NT;
};

template <int CIN, int NT, int MODE, int ES = 2>
__global__ void __launch_bounds__(kHaloThreads, HaloCfg<CIN, NT, ES>::CTAS) halo_conv_kernel(const __grid_constant__ HaloParams p) {
  using Cfg = HaloCfg<CIN, NT, ES>;
  constexpr bool TF32 = ES == 4;
  constexpr int EPC = 16 / ES;                          // elements per 16-byte chunk
  constexpr int P = Cfg::P, STAGES = Cfg::STAGES, RS = Cfg::RS;
  constexpr uint32_t LAYOUT = halo_layout<P>();
  constexpr int CMASK = P / 16 - 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  uint8_t* b_tile = a_ring + STAGES * Cfg::kStage;      // 9 taps x [NT rows x P bytes]
  uint8_t* stage = b_tile + 9 * Cfg::kBTile;            // 8 epilogue warps x 32 rows x RS
  float* coef = reinterpret_cast<float*>(stage + Cfg::kEpi);           // scale[64] shift[64]
  uint64_t* full = reinterpret_cast<uint64_t*>(coef + 128);
  uint64_t* empty = full + STAGES;
  uint64_t* t_full = empty + STAGES;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 8); }
    fence_barrier_init();
    tma_prefetch_desc(&p.x_map);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  // weights: row n of tap tile t holds Wp[n][t * CIN .. + CIN), swizzled like a TMA write would
  for (int e = threadIdx.x; e < NT * 9 * (P / 16); e += blockDim.x) {
    const int n = e / (9 * (P / 16)), r = e - n * (9 * (P / 16));
    const int t = r / (P / 16), c = r - t * (P / 16);
    const uint32_t off = t * Cfg::kBTile + n * P;
    const uint32_t addr = smem_u32(b_tile) + off;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n < p.n_real && c * EPC < p.cin_real)
      v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.wp) +
                                          ((size_t)n * 9 * p.cin_real + t * p.cin_real + c * EPC) * ES);
    *reinterpret_cast<uint4*>(b_tile + off + ((c ^ ((addr >> 7) & CMASK)) << 4)) = v;
  }
  if (MODE == 1 && threadIdx.x < 128) {
    const int c = threadIdx.x & 63, which = threadIdx.x >> 6;
    coef[threadIdx.x] = c < p.n_real ? p.affine[which * p.n_real + c] : 0.f;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer: one halo box per tile
    if (lane == 0) {
      uint32_t s = 0, ph = 1;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
        const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
        mbar_wait(&empty[s], ph);
        mbar_expect_tx(&full[s], Cfg::kBox);
        tma_load_4d(a_ring + s * Cfg::kStage, &p.x_map, &full[s], 0, tx * 16 - 1, ty * 16 - 1, b);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc(TF32, false, false, 128, NT);
    const uint32_t a_base = smem_u32(a_ring), b_base = smem_u32(b_tile);
    uint32_t s = 0, ph = 0, acc = 0, pacc = 1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      mbar_wait(&full[s], ph);
      mbar_wait(&t_empty[acc], pacc);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int strip = 0; strip < 2; ++strip) {
          const uint32_t d = tmem_base + (acc * 2 + strip) * NT;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t a0 = a_base + s * Cfg::kStage + (p.tap_pix[t] + 8 * strip) * P;
#pragma unroll
            for (int kk = 0; kk < P / 32; ++kk)               // 32 bytes of K per MMA: 16 bf16 or 8 tf32 values
              umma<TF32>(d, make_desc(a0 + kk * 32, 16, kHaloBW * P, LAYOUT),
                          make_desc(b_base + t * Cfg::kBTile + kk * 32, 16, 8 * P, LAYOUT), idesc, (t | kk) ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);
        umma_commit(&t_full[acc]);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
  } else {
    // ------------------------------------------------ epilogue: warp drains TMEM lanes [32 (warp % 4), +32) of one strip
    const int e = warp - 2, strip = e >> 2, quad = warp & 3;
    const uint32_t stg = smem_u32(stage) + e * 32 * RS;
    auto swz = [](int row) { return RS == 128 ? (row & 7) : (RS == 64 ? ((row >> 1) & 3) : 0); };
    // statistics lanes: bf16 -> a lane owns a channel PAIR (one 32-bit word of the staged row), fp32 -> one channel
    constexpr int PP = TF32 ? NT : NT / 2, G = 32 / PP;     // lanes per row; row groups
    const int pair = lane % PP, grp = lane / PP;
    constexpr int cpr = RS / 16;                        // 16-byte chunks per stored row
    uint32_t acc = 0, pacc = 0;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
    float bmu[2] = {0.f, 0.f}, bis[2] = {0.f, 0.f}, bsc[2] = {0.f, 0.f}, bsh[2] = {0.f, 0.f};
    if constexpr (MODE == 2) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int ch = 2 * pair + k;
        if (ch < p.n_real) {
          bmu[k] = __ldg(p.bnc + ch); bis[k] = __ldg(p.bnc + p.n_real + ch);
          bsc[k] = __ldg(p.bnc + 2 * p.n_real + ch); bsh[k] = __ldg(p.bnc + 3 * p.n_real + ch);
        }
      }
    }
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
      const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
      const int y0 = ty * 16 + quad * 4, x0 = tx * 16 + strip * 8;       // row m = 8 * y + x of the strip
      const bool live = (y0 + (lane >> 3)) < p.H && (x0 + (lane & 7)) < p.W;
      // MODE 2: this lane's yprev words (rows grp, grp + G, ...; channel pair `pair`) are requested before the wait for
      // the accumulators, so their latency hides behind the MMAs of the tile
      uint32_t yv[MODE == 2 ? 32 / G : 1];
      if constexpr (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 32 / G; ++i) {
          const int rr = grp + i * G;
          const int yy = y0 + (rr >> 3), xx = x0 + (rr & 7);
          yv[i] = 0u;
          if (yy < p.H && xx < p.W && 2 * pair < p.n_real)
            yv[i] = __ldg(reinterpret_cast<const uint32_t*>(p.yprev + (((long long)b * p.H + yy) * p.W + xx) * p.ld_y) + pair);
        }
      }
      mbar_wait(&t_full[acc], pacc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (acc * 2 + strip) * NT;
      __syncwarp();
#pragma unroll
      for (int h = 0; h < (NT + 31) / 32; ++h) {
        constexpr int W = NT < 32 ? NT : 32;
        uint32_t v[W];
        if constexpr (NT < 32) halo_tmem_ld16(taddr, v);
        else tmem_ld32(taddr + h * 32, v);
        if (h == (NT + 31) / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&t_empty[acc])) : "memory");
        }
        if constexpr (TF32) {
#pragma unroll
          for (int c = 0; c < W; ++c) {
            float v0 = __uint_as_float(v[c]);
            if (MODE == 1) v0 = fmaxf(fmaf(v0, coef[h * 32 + c], coef[64 + h * 32 + c]), 0.f);
            v[c] = live ? __float_as_uint(v0) : 0u;
          }
#pragma unroll
          for (int c = 0; c < W / 4; ++c)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * RS + (((h * 8 + c) ^ swz(lane)) << 4)),
                         "r"(v[4 * c]), "r"(v[4 * c + 1]), "r"(v[4 * c + 2]), "r"(v[4 * c + 3]) : "memory");
        } else {
          uint32_t o[W / 2];
#pragma unroll
          for (int c = 0; c < W / 2; ++c) {
            float v0 = __uint_as_float(v[2 * c]), v1 = __uint_as_float(v[2 * c + 1]);
            if (MODE == 1) {
              const int ch = h * 32 + 2 * c;
              v0 = fmaxf(fmaf(v0, coef[ch], coef[64 + ch]), 0.f);
              v1 = fmaxf(fmaf(v1, coef[ch + 1], coef[64 + ch + 1]), 0.f);
            }
            o[c] = live ? halo_pack(v0, v1) : 0u;
          }
#pragma unroll
          for (int c = 0; c < W / 8; ++c)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * RS + (((h * 4 + c) ^ swz(lane)) << 4)),
                         "r"(o[4 * c]), "r"(o[4 * c + 1]), "r"(o[4 * c + 2]), "r"(o[4 * c + 3]) : "memory");
        }
      }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
      __syncwarp();
      // transposed store: one instruction = 32 / cpr rows x (16 cpr) bytes; 8 rows (one image row of the strip) are
      // contiguous in memory
      {
        constexpr int rpi = 32 / cpr;
        const int ch = lane % cpr;
#pragma unroll
        for (int i = 0; i < cpr; ++i) {
          const int rr = i * rpi + lane / cpr;
          uint4 q;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                       : "r"(stg + rr * RS + ((ch ^ swz(rr)) << 4)) : "memory");
          const int yy = y0 + (rr >> 3), xx = x0 + (rr & 7);
          if (yy < p.H && xx < p.W && ch * EPC < p.n_real)
            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.y) +
                                      ((((long long)b * p.H + yy) * p.W + xx) * p.ld_out + ch * EPC) * ES) = q;
        }
      }
      if constexpr (MODE == 1 && !TF32) {
        if (p.pooled) {
          // 8 pooled pixels x NT / 2 channel-pair words per patch; activations are >= 0, so the unsigned 16-bit SIMD
          // maximum of the bf16 bit patterns is the bf16 maximum
          const int Hp = p.H >> 1, Wp = p.W >> 1;
#pragma unroll
          for (int idx = lane; idx < 8 * (NT / 2); idx += 32) {
            const int pp = idx / (NT / 2), w = idx % (NT / 2);
            const int r0 = 16 * (pp >> 2) + 2 * (pp & 3);
            uint32_t m = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int rr = r0 + (k & 1) + 8 * (k >> 1);
              uint32_t u;
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u)
                           : "r"(stg + rr * RS + ((((w >> 2) ^ swz(rr)) << 4) | ((w & 3) << 2))) : "memory");
              m = __vmaxu2(m, u);
            }
            const int gy = (y0 >> 1) + (pp >> 2), gx = (x0 >> 1) + (pp & 3);
            if (gy < Hp && gx < Wp && 2 * w < p.n_real)
              reinterpret_cast<uint32_t*>(p.pooled + (((long long)b * Hp + gy) * Wp + gx) * p.ld_pool)[w] = m;
          }
        }
      }
      if constexpr (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 32 / G; ++i) {
          const int rr = grp + i * G;
          uint32_t u;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u)
                       : "r"(stg + rr * RS + ((((pair >> 2) ^ swz(rr)) << 4) | ((pair & 3) << 2))) : "memory");
          // rows outside the image were staged as zeros and fetched no yprev: they add nothing
          const float ya = __uint_as_float(yv[i] << 16), yb = __uint_as_float(yv[i] & 0xffff0000u);
          const float ga = (fmaf(ya, bsc[0], bsh[0]) > 0.f) ? __uint_as_float(u << 16) : 0.f;
          const float gb = (fmaf(yb, bsc[1], bsh[1]) > 0.f) ? __uint_as_float(u & 0xffff0000u) : 0.f;
          s0 += ga; q0 += ga * ((ya - bmu[0]) * bis[0]);
          s1 += gb; q1 += gb * ((yb - bmu[1]) * bis[1]);
        }
      }
      if (MODE == 0 && p.stats_ws) {
#pragma unroll
        for (int rr = grp; rr < 32; rr += G) {
          uint32_t u;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u)
                       : "r"(stg + rr * RS + ((((pair >> 2) ^ swz(rr)) << 4) | ((pair & 3) << 2))) : "memory");
          if constexpr (TF32) {
            const float a = __uint_as_float(u);
            s0 += a; q0 = fmaf(a, a, q0);
          } else {
            const float a = __uint_as_float(u << 16), bq = __uint_as_float(u & 0xffff0000u);
            s0 += a; q0 = fmaf(a, a, q0); s1 += bq; q1 = fmaf(bq, bq, q1);
          }
        }
      }
    }
    if ((MODE == 0 || MODE == 2) && p.stats_ws) {
#pragma unroll
      for (int o = PP; o < 32; o <<= 1) {               // combine the row groups (fixed order)
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
      }
      float* dst = p.stats_ws + ((long long)blockIdx.x * 8 + e) * 2 * p.n_real;
      if constexpr (TF32) {
        if (grp == 0 && pair < p.n_real) { dst[pair] = s0; dst[p.n_real + pair] = q0; }
      } else {
        if (grp == 0 && 2 * pair < p.n_real) {
          dst[2 * pair] = s0; dst[2 * pair + 1] = s1;
          dst[p.n_real + 2 * pair] = q0; dst[p.n_real + 2 * pair + 1] = q1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------ host side (fprop)
static bool halo_ch_ok(int c) { return c == 8 || c == 16 || c == 32 || c == 64; }
static int halo_ch_pad(int c) { return c < 16 ? 16 : c; }          // 8 channels ride the 16-channel instantiation

// fp32 (TF32 on the tensor cores, fprop / dgrad only): 8 / 16 / 32 channels on either side -- a 64-channel fp32 row is two
// swizzle rows, and tcgen05 has no MN-major 32-bit layout for 64-byte rows (the weight gradient stays exact fp32 on the
// CUDA cores, conv_simt_narrow.cu)
static bool halo_shape_ok(const unetb200_gconv_t* d, bool wgrad = false) {
  static const bool off = getenv("UNETB200_NO_HALO") != nullptr;
  static const bool off32 = getenv("UNETB200_NO_HALO_TF32") != nullptr;
  if (off || (d->dtype != UNETB200_BF16 && d->dtype != UNETB200_F32)) return false;
  if (d->dtype == UNETB200_F32) {
    if (wgrad || off32 || (d->algo != UNETB200_ALGO_TC && d->algo != UNETB200_ALGO_PREFER_TC)) return false;
    if (d->Cin > 32 || d->N > 32) return false;
  }
  if (d->ntaps != 9 || d->in_scale != 1 || d->out_scale != 1 || d->nquad != 1) return false;
  if (d->in_off_y || d->in_off_x || d->out_off_y || d->out_off_x) return false;
  if (d->Hm != d->Hout || d->Wm != d->Wout || d->Hm != d->Hin || d->Wm != d->Win) return false;
  if (!halo_ch_ok(d->Cin) || !halo_ch_ok(d->N)) return false;
  if (d->Cin == 64 && d->N == 64) return false;                    // the CTA-pair kernel (conv_tc3.cu) covers 64 -> 64
  const int vec = d->dtype == UNETB200_BF16 ? 8 : 4;               // 16-byte pixel strides
  if ((d->ld_in % vec) || (d->ld_out % vec)) return false;
  bool seen[9] = {false};
  for (int t = 0; t < 9; ++t) {
    const int dy = d->tap_dy[t], dx = d->tap_dx[t];
    if (dy < -1 || dy > 1 || dx < -1 || dx > 1 || seen[(dy + 1) * 3 + dx + 1]) return false;
    seen[(dy + 1) * 3 + dx + 1] = true;
  }
  return (long long)d->B * d->Hm * d->Wm < (1LL << 31) - 256;
}

int halo_fprop_supported(const unetb200_gconv_t* d, const void* x, const void* wp, const void* y) {
  if (!halo_shape_ok(d)) return 0;
  return aligned16(x) && aligned16(wp) && aligned16(y);
}

template <int CIN, int NT, int ES = 2>
static int halo_ctas() { return HaloCfg<CIN, NT, ES>::CTAS; }

static int halo_ctas_per_sm(int cin, int n, bool f32 = false) {
  cin = halo_ch_pad(cin); n = halo_ch_pad(n);
  if (f32) {
    if (cin == 16) return n == 16 ? halo_ctas<16, 16, 4>() : halo_ctas<16, 32, 4>();
    return n == 16 ? halo_ctas<32, 16, 4>() : halo_ctas<32, 32, 4>();
  }
  if (cin == 16) return n == 16 ? halo_ctas<16, 16>() : (n == 32 ? halo_ctas<16, 32>() : halo_ctas<16, 64>());
  if (cin == 32) return n == 16 ? halo_ctas<32, 16>() : (n == 32 ? halo_ctas<32, 32>() : halo_ctas<32, 64>());
  return n == 16 ? halo_ctas<64, 16>() : halo_ctas<64, 32>();
}

static int halo_grid(const unetb200_gconv_t* d, int* tiles_w, int* tiles_h, int* ntiles) {
  *tiles_w = (d->Wm + 15) / 16;
  *tiles_h = (d->Hm + 15) / 16;
  *ntiles = d->B * *tiles_w * *tiles_h;
  const int slots = halo_ctas_per_sm(d->Cin, d->N, d->dtype == UNETB200_F32) * sm_count();
  return *ntiles < slots ? *ntiles : slots;
}

long long halo_stats_rows(const unetb200_gconv_t* d) {
  if (!halo_shape_ok(d)) return 0;
  int tw, th, nt;
  return (long long)halo_grid(d, &tw, &th, &nt) * 8;
}

template <int CIN, int NT, int MODE, int ES = 2>
static int halo_launch(const HaloParams& P, int grid, cudaStream_t s) {
  constexpr int smem = HaloCfg<CIN, NT, ES>::smem;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(&halo_conv_kernel<CIN, NT, MODE, ES>), smem, "halo_conv smem attribute"))
    return rc;
  halo_conv_kernel<CIN, NT, MODE, ES><<<grid, kHaloThreads, smem, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "halo_conv launch");
  return 0;
}

template <int CIN>
static int halo_dispatch_n(const HaloParams& P, int n, int grid, bool affine, cudaStream_t s) {
  if (affine) {
    if (n == 16) return halo_launch<CIN, 16, 1>(P, grid, s);
    if (n == 32) return halo_launch<CIN, 32, 1>(P, grid, s);
    if constexpr (CIN < 64) return halo_launch<CIN, 64, 1>(P, grid, s);
  } else {
    if (n == 16) return halo_launch<CIN, 16, 0>(P, grid, s);
    if (n == 32) return halo_launch<CIN, 32, 0>(P, grid, s);
    if constexpr (CIN < 64) return halo_launch<CIN, 64, 0>(P, grid, s);
  }
  set_error("halo_conv: unsupported channel counts");
  return UNETB200_E_INVALID;
}

template <int CIN>
static int halo_dispatch_n32(const HaloParams& P, int n, int grid, bool affine, cudaStream_t s) {
  if (affine) return n == 16 ? halo_launch<CIN, 16, 1, 4>(P, grid, s) : halo_launch<CIN, 32, 1, 4>(P, grid, s);
  return n == 16 ? halo_launch<CIN, 16, 0, 4>(P, grid, s) : halo_launch<CIN, 32, 0, 4>(P, grid, s);
}

// dgrad + BatchNorm-backward reduction (MODE 2): bf16, at most 32 output channels (the prefetched yprev words live in
// registers: 32 / G per lane)
int halo_bnbwd_supported(const unetb200_gconv_t* d, const void* g, const void* wp, const void* gx) {
  static const bool off = getenv("UNETB200_NO_HALO_BNBWD") != nullptr;
  return !off && d->dtype == UNETB200_BF16 && d->N <= 32 && halo_fprop_supported(d, g, wp, gx);
}

int halo_fprop(const unetb200_gconv_t* d, const void* x, const void* wp, void* y, double* stats, float* stats_ws,
               const float* affine, cudaStream_t s, const void* yprev, long long ld_yprev, const float* bnc, void* pooled,
               long long ld_pool) {
  if (!halo_fprop_supported(d, x, wp, y)) { set_error("halo_fprop: unsupported shape"); return UNETB200_E_INVALID; }
  if (pooled && (!affine || d->dtype != UNETB200_BF16 || (ld_pool & 1) || (reinterpret_cast<uintptr_t>(pooled) & 3))) {
    set_error("halo_fprop: the pooled second output needs the bf16 affine epilogue and 4-byte aligned rows");
    return UNETB200_E_INVALID;
  }
  if (yprev && (!halo_bnbwd_supported(d, x, wp, y) || affine || !stats || !stats_ws || !bnc || (ld_yprev & 1) ||
                (reinterpret_cast<uintptr_t>(yprev) & 3))) {
    set_error("halo_fprop: the BatchNorm-backward epilogue needs bf16, N <= 32, sums, a workspace and coefficients");
    return UNETB200_E_INVALID;
  }
  HaloParams P;
  memset(&P, 0, sizeof(P));
  P.yprev = (const __nv_bfloat16*)yprev; P.ld_y = ld_yprev; P.bnc = bnc;
  P.pooled = (__nv_bfloat16*)pooled; P.ld_pool = ld_pool;
  const bool f32 = d->dtype == UNETB200_F32;
  if (int rc = encode_act_box_sw(&P.x_map, x, d->Cin, d->Wm, d->Hm, d->B, d->ld_in, (long long)d->Wm * d->ld_in,
                                 (long long)d->Hm * d->Wm * d->ld_in, kHaloBW, kHaloBW, f32 ? 4 : 2))
    return rc;
  P.cin_real = d->Cin; P.n_real = d->N;
  P.wp = wp; P.y = y;
  P.affine = affine;
  P.stats_ws = (stats && !affine) ? stats_ws : nullptr;
  P.ld_out = d->ld_out;
  P.H = d->Hm; P.W = d->Wm;
  for (int t = 0; t < 9; ++t) P.tap_pix[t] = (d->tap_dy[t] + 1) * kHaloBW + d->tap_dx[t] + 1;
  const int grid = halo_grid(d, &P.tiles_w, &P.tiles_h, &P.ntiles);
  int rc;
  if (yprev) {
    const int cp = halo_ch_pad(d->Cin), np = halo_ch_pad(d->N);
    if (cp == 16) rc = np == 16 ? halo_launch<16, 16, 2>(P, grid, s) : halo_launch<16, 32, 2>(P, grid, s);
    else if (cp == 32) rc = np == 16 ? halo_launch<32, 16, 2>(P, grid, s) : halo_launch<32, 32, 2>(P, grid, s);
    else rc = np == 16 ? halo_launch<64, 16, 2>(P, grid, s) : halo_launch<64, 32, 2>(P, grid, s);
  } else if (f32) {
    rc = halo_ch_pad(d->Cin) == 16 ? halo_dispatch_n32<16>(P, halo_ch_pad(d->N), grid, affine != nullptr, s)
                                   : halo_dispatch_n32<32>(P, halo_ch_pad(d->N), grid, affine != nullptr, s);
  } else
  switch (halo_ch_pad(d->Cin)) {
    case 16: rc = halo_dispatch_n<16>(P, halo_ch_pad(d->N), grid, affine != nullptr, s); break;
    case 32: rc = halo_dispatch_n<32>(P, halo_ch_pad(d->N), grid, affine != nullptr, s); break;
    default: rc = halo_dispatch_n<64>(P, halo_ch_pad(d->N), grid, affine != nullptr, s); break;
  }
  if (rc) return rc;
  if (P.stats_ws) return launch_stats_reduce(stats_ws, (long long)grid * 8, 2 * d->N, stats, s);
  return 0;
}

// ------------------------------------------------------------------------------------------ wgrad
// dW[(t, c)][n] = sum_p x[p + t][c] * g[p][n].  A tile is 16 x 16 pixels: X box {Cin, 18, 16} (columns x0 - 1 ..),
// G box {N, 16, 18} (rows y0 - 1 ..), zero filled outside the image.  Row r of the tile is ONE tcgen05.mma
// (K = 16 pixels): A = X row r, MN-major, M = 128 = atoms of Cin channels one pixel apart (atom h = column shift
// dx = h - 1; atoms h >= 3 are junk rows of D nobody reads), B = G rows r, r + 1, r + 2, MN-major, N = 3 atoms of N
// channels one box row apart (atom s = row shift dy = 1 - s).  Cin = 64 has two atoms per MMA and issues a second one
// (h = 2, 3) into a second accumulator.
struct HaloWParams {
  CUtensorMap x_map;             // {Cin, W, H, B}, box {Cin, 18, 16, 1}
  CUtensorMap g_map;             // {N, W, H, B}, box {N, 16, 18, 1}
  float* partials;               // [grid][9 * Cin][N]
  int tiles_w, tiles_h, ntiles;
  int tap_of[9];                 // [(dy + 1) * 3 + dx + 1] -> tap index of the descriptor
  int cin_real, n_real;          // 8-channel operands: atoms of 16 channels whose upper half TMA zero-fills
};

template <int CIN, int NT>
struct HaloWCfg {
  static constexpr int PX = 2 * CIN, PG = 2 * NT;
  static constexpr uint32_t kXBox = 16 * kHaloBW * PX, kGBox = kHaloBW * 16 * PG;
  static constexpr uint32_t kXStage = (kXBox + 8 * PX + 1023) / 1024 * 1024;      // + the junk atoms' overrun
  static constexpr uint32_t kGStage = (kGBox + 1023) / 1024 * 1024;
  static constexpr uint32_t kStage = kXStage + kGStage;
  static constexpr int NMMA = CIN == 64 ? 2 : 1;
  static constexpr int kUsed = NMMA * 3 * NT;
  static constexpr int kCols = kUsed <= 64 ? 64 : (kUsed <= 128 ? 128 : 256);
  static constexpr int kFixed = 256 + 1024;
  static constexpr int S2 = (110 * 1024 - kFixed) / (int)kStage;                  // stages if two CTAs share an SM
  static constexpr int CTAS = S2 >= 2 ? 2 : 1;
  static constexpr int S1 = (220 * 1024 - kFixed) / (int)kStage;
  static constexpr int STAGES = CTAS == 2 ? (S2 > 4 ? 4 : S2) : (S1 > 4 ? 4 : S1);
  static constexpr int smem = STAGES * (int)kStage + kFixed;
};

template <int CIN, int NT>
__global__ void __launch_bounds__(192, HaloWCfg<CIN, NT>::CTAS) halo_wgrad_kernel(const __grid_constant__ HaloWParams p) {
  using Cfg = HaloWCfg<CIN, NT>;
  constexpr int PX = Cfg::PX, PG = Cfg::PG, STAGES = Cfg::STAGES, NMMA = Cfg::NMMA;
  constexpr int K = 9 * CIN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStage);
  uint64_t* empty = full + STAGES;
  uint64_t* t_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(t_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.x_map);
    tma_prefetch_desc(&p.g_map);
  }
  if (warp == 5) tmem_alloc(tmem_slot, Cfg::kCols);
  // the junk atoms of the last box row read up to 8 pixels past the X box: keep those bytes finite
  for (int s = 0; s < STAGES; ++s)
    for (uint32_t i = Cfg::kXBox + threadIdx.x * 16; i < Cfg::kXStage; i += blockDim.x * 16)
      *reinterpret_cast<uint4*>(smem + s * Cfg::kStage + i) = make_uint4(0u, 0u, 0u, 0u);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 4) {
    if (lane == 0) {
      uint32_t s = 0, ph = 1;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
        const int ty = r / p.tiles_w, tx = r - ty * p.tiles_w;
        mbar_wait(&empty[s], ph);
        mbar_expect_tx(&full[s], Cfg::kXBox + Cfg::kGBox);
        uint8_t* st = smem + s * Cfg::kStage;
        tma_load_4d(st, &p.x_map, &full[s], 0, tx * 16 - 1, ty * 16, b);
        tma_load_4d(st + Cfg::kXStage, &p.g_map, &full[s], 0, tx * 16, ty * 16 - 1, b);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    constexpr uint32_t idesc = make_idesc(false, true, true, 128, 3 * NT);
    constexpr uint32_t LX = halo_layout<PX>(), LG = halo_layout<PG>();
    uint32_t s = 0, ph = 0;
    int done = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t xs = smem_u32(smem) + s * Cfg::kStage, gs = xs + Cfg::kXStage;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const uint64_t db = make_desc(gs + r * 16 * PG, 16 * PG, 8 * PG, LG);
#pragma unroll
          for (int mb = 0; mb < NMMA; ++mb)
            umma<false>(tmem_base + mb * 3 * NT, make_desc(xs + (r * kHaloBW + 2 * mb) * PX, PX, 8 * PX, LX), db, idesc,
                        (done > 0 || r > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
      ++done;
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(t_full);
    __syncwarp();
  } else {
    // ---- end of the walk: thread = row m = h * CIN + c of each accumulator
    const int m = threadIdx.x;
    mbar_wait(t_full, 0);
    tc_fence_after();
    float* out = p.partials + (long long)blockIdx.x * 9 * p.cin_real * p.n_real;
#pragma unroll 1
    for (int mb = 0; mb < NMMA; ++mb) {
      const int h = (CIN == 64 ? 2 * mb : 0) + m / CIN, c = m % CIN;
#pragma unroll 1
      for (int sft = 0; sft < 3; ++sft) {
        const int t = h < 3 ? p.tap_of[(2 - sft) * 3 + h] : 0;
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t v[16];
          halo_tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + mb * 3 * NT + sft * NT + c0, v);
          if (h < 3 && c < p.cin_real) {
            float4* dst = reinterpret_cast<float4*>(out + ((long long)t * p.cin_real + c) * p.n_real + c0);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (c0 + 4 * e < p.n_real)
                dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                                     __uint_as_float(v[4 * e + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tmem_dealloc(tmem_base, Cfg::kCols);
  }
}

template <int CIN, int NT>
static int halo_w_ctas() { return HaloWCfg<CIN, NT>::CTAS; }

static int halo_w_ctas_per_sm(int cin, int n) {
  cin = halo_ch_pad(cin); n = halo_ch_pad(n);
  if (cin == 16) return n == 16 ? halo_w_ctas<16, 16>() : (n == 32 ? halo_w_ctas<16, 32>() : halo_w_ctas<16, 64>());
  if (cin == 32) return n == 16 ? halo_w_ctas<32, 16>() : (n == 32 ? halo_w_ctas<32, 32>() : halo_w_ctas<32, 64>());
  return n == 16 ? halo_w_ctas<64, 16>() : halo_w_ctas<64, 32>();
}

static int halo_wgrad_grid(const unetb200_gconv_t* d, int* tiles_w, int* tiles_h, int* ntiles) {
  *tiles_w = (d->Wm + 15) / 16;
  *tiles_h = (d->Hm + 15) / 16;
  *ntiles = d->B * *tiles_w * *tiles_h;
  const int slots = halo_w_ctas_per_sm(d->Cin, d->N) * sm_count();
  return *ntiles < slots ? *ntiles : slots;
}

int halo_wgrad_supported(const unetb200_gconv_t* d, const void* x, const void* gy) {
  static const bool off = getenv("UNETB200_NO_HALO_WGRAD") != nullptr;
  if (off || !halo_shape_ok(d, true)) return 0;
  if ((x && !aligned16(x)) || (gy && !aligned16(gy))) return 0;
  return 1;
}

int halo_wgrad_splits(const unetb200_gconv_t* d) {
  int tw, th, nt;
  return halo_wgrad_grid(d, &tw, &th, &nt);
}

template <int CIN, int NT>
static int halo_wgrad_launch(const HaloWParams& P, int grid, cudaStream_t s) {
  constexpr int smem = HaloWCfg<CIN, NT>::smem;
  static_assert(smem <= 227 * 1024 && HaloWCfg<CIN, NT>::STAGES >= 2, "shared memory budget");
  if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(&halo_wgrad_kernel<CIN, NT>), smem, "halo_wgrad smem attribute"))
    return rc;
  halo_wgrad_kernel<CIN, NT><<<grid, 192, smem, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "halo_wgrad launch");
  return 0;
}

template <int CIN>
static int halo_wgrad_dispatch_n(const HaloWParams& P, int n, int grid, cudaStream_t s) {
  if (n == 16) return halo_wgrad_launch<CIN, 16>(P, grid, s);
  if (n == 32) return halo_wgrad_launch<CIN, 32>(P, grid, s);
  if constexpr (CIN < 64) return halo_wgrad_launch<CIN, 64>(P, grid, s);
  set_error("halo_wgrad: unsupported channel counts");
  return UNETB200_E_INVALID;
}

int halo_wgrad(const unetb200_gconv_t* d, const void* x, const void* gy, float* partials, int splits, cudaStream_t s) {
  if (!halo_wgrad_supported(d, x, gy)) { set_error("halo_wgrad: unsupported shape"); return UNETB200_E_INVALID; }
  HaloWParams P;
  memset(&P, 0, sizeof(P));
  if (int rc = encode_act_box_sw(&P.x_map, x, d->Cin, d->Wm, d->Hm, d->B, d->ld_in, (long long)d->Wm * d->ld_in,
                                 (long long)d->Hm * d->Wm * d->ld_in, kHaloBW, 16))
    return rc;
  if (int rc = encode_act_box_sw(&P.g_map, gy, d->N, d->Wm, d->Hm, d->B, d->ld_out, (long long)d->Wm * d->ld_out,
                                 (long long)d->Hm * d->Wm * d->ld_out, 16, kHaloBW))
    return rc;
  P.partials = partials;
  P.cin_real = d->Cin; P.n_real = d->N;
  for (int t = 0; t < 9; ++t) P.tap_of[(d->tap_dy[t] + 1) * 3 + d->tap_dx[t] + 1] = t;
  const int grid = halo_wgrad_grid(d, &P.tiles_w, &P.tiles_h, &P.ntiles);
  if (grid != splits) { set_error("halo_wgrad: the planned split count is %d, got %d", grid, splits); return UNETB200_E_INVALID; }
  switch (halo_ch_pad(d->Cin)) {
    case 16: return halo_wgrad_dispatch_n<16>(P, halo_ch_pad(d->N), grid, s);
    case 32: return halo_wgrad_dispatch_n<32>(P, halo_ch_pad(d->N), grid, s);
    default: return halo_wgrad_dispatch_n<64>(P, halo_ch_pad(d->N), grid, s);
  }
}

}  // namespace ub
