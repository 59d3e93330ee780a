// Exact-fp32 forward / data gradient and weight gradient of the narrow 3x3 layers on the CUDA cores (UNet_S / UNet_T / UNet_SA without autocast:
// the tcgen05 narrow kernels are bf16, and tcgen05 has no MN-major 32-bit layout for 64-byte rows -- DESIGN.md section 7).
//
// The generic split-K engine (gconv_wgrad_simt_kernel) re-decodes a pixel index per loaded vector and synchronises every
// 16 pixels: 2.6 ms per layer where the FMA floor is 0.26 ms.  Here dW[t][c][n] is register resident: a thread owns
// one input channel c and 8 output channels (72 accumulators = 9 taps x 8), a block = C_in x N/8 threads walks 8-row
// pixel tiles whose x halo and dY tile are staged in shared memory (zeros outside the image), sliding its 3x3 window
// of x down each pixel column: 72 FMAs per pixel against 3 + 2 shared-memory loads.  Persistent blocks, one fp32
// partial per block, reduced by the ordinary split reduction in a fixed order.
#include <cstring>

#include "gconv.cuh"

namespace ub {

constexpr int kWnH = 8;

template <int CIN, int N>
struct WnCfg {
  // TPB threads = one (input channel, 8-output-channel group) each; below a warp (8-channel tensors of UNet_T) LANES
  // copies of them share the block, walking interleaved pixel columns, and are combined by shuffles at the end
  static constexpr int NG = N / 8, TPB = CIN * NG, LANES = TPB >= 32 ? 1 : 32 / TPB, THREADS = TPB * LANES;
  // small blocks (<= 2 warps) take 8-column tiles: 10 KB of shared memory each, so 16 blocks (the register limit at 128
  // registers) are resident -- at 9 blocks of one warp the FMA pipe was 41 % active (ncu), two warps per scheduler
  static constexpr int TW = (CIN >= 64 || (CIN == 32 && N == 64) || TPB <= 64) ? 8 : 16;
  static constexpr int XS = (kWnH + 2) * (TW + 2) * CIN, GS = kWnH * TW * N;     // floats
  static constexpr int smem = (XS + GS) * 4;
};

struct WnParams {
  const float* x;                // [B][H][W][ld_in]
  const float* gy;               // [B][H][W][ld_out]
  float* partials;               // [grid][9 * Cin][N]
  long long ld_in, ld_out;
  int H, W, tiles_w, tiles_h, ntiles;
  int tap_of[9];
};

template <int CIN, int N>
__global__ void __launch_bounds__(WnCfg<CIN, N>::THREADS) wgrad_narrow_f32_kernel(const WnParams p) {
  using Cfg = WnCfg<CIN, N>;
  constexpr int TW = Cfg::TW, NT = Cfg::THREADS;
  extern __shared__ __align__(16) float wn_smem[];
  float* xs = wn_smem;                       // [kWnH + 2][TW + 2][CIN]
  float* gs = wn_smem + Cfg::XS;             // [kWnH][TW][N]
  constexpr int TPB = Cfg::TPB, LANES = Cfg::LANES;
  const int cl = threadIdx.x / TPB, tq = threadIdx.x % TPB;        // column lane; (cin, ng) index
  const int cin = tq % CIN, ng = tq / CIN;
  float acc[3][3][8];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[a][c][k] = 0.f;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int tj = tile % p.tiles_w, rest = tile / p.tiles_w;
    const int b = rest / p.tiles_h, i0 = (rest % p.tiles_h) * kWnH, j0 = tj * TW;
    __syncthreads();
    for (int e = threadIdx.x; e < (kWnH + 2) * (TW + 2) * (CIN / 4); e += NT) {
      const int c4 = e % (CIN / 4), px = e / (CIN / 4);
      const int r = px / (TW + 2), c = px - r * (TW + 2);
      const int gi = i0 - 1 + r, gj = j0 - 1 + c;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((unsigned)gi < (unsigned)p.H && (unsigned)gj < (unsigned)p.W)
        v = *reinterpret_cast<const float4*>(p.x + ((long long)(b * p.H + gi) * p.W + gj) * p.ld_in + c4 * 4);
      *reinterpret_cast<float4*>(xs + px * CIN + c4 * 4) = v;
    }
    for (int e = threadIdx.x; e < kWnH * TW * (N / 4); e += NT) {
      const int n4 = e % (N / 4), px = e / (N / 4);
      const int r = px / TW, c = px - r * TW;
      const int gi = i0 + r, gj = j0 + c;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gi < p.H && gj < p.W)
        v = *reinterpret_cast<const float4*>(p.gy + ((long long)(b * p.H + gi) * p.W + gj) * p.ld_out + n4 * 4);
      *reinterpret_cast<float4*>(gs + px * N + n4 * 4) = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int col = cl; col < TW; col += LANES) {
      float win[3][3];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) win[a][c] = xs[(a * (TW + 2) + col + c) * CIN + cin];
#pragma unroll
      for (int r = 0; r < kWnH; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) win[2][c] = xs[((r + 2) * (TW + 2) + col + c) * CIN + cin];
        const float4 g0 = *reinterpret_cast<const float4*>(gs + (r * TW + col) * N + ng * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(gs + (r * TW + col) * N + ng * 8 + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[a][c][k] = fmaf(win[a][c], g[k], acc[a][c][k]);
#pragma unroll
        for (int c = 0; c < 3; ++c) { win[0][c] = win[1][c]; win[1][c] = win[2][c]; }
      }
    }
  }
  if constexpr (LANES > 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int o = TPB; o < 32; o <<= 1) acc[a][c][k] += __shfl_xor_sync(0xffffffffu, acc[a][c][k], o);
    if (cl != 0) return;
  }
  float* out = p.partials + (long long)blockIdx.x * 9 * CIN * N;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int t = p.tap_of[a * 3 + c];
      float4* dst = reinterpret_cast<float4*>(out + ((long long)t * CIN + cin) * N + ng * 8);
      dst[0] = make_float4(acc[a][c][0], acc[a][c][1], acc[a][c][2], acc[a][c][3]);
      dst[1] = make_float4(acc[a][c][4], acc[a][c][5], acc[a][c][6], acc[a][c][7]);
    }
}

// ------------------------------------------------------------------------------------------ fprop / dgrad
// y[p][n] = sum_{t,c} x[p + t][c] * Wp[n][t * Cin + c].  A thread owns 4 horizontally adjacent pixels x 8 output channels
// (32 accumulators); the x halo of an 8-row tile sits in shared memory channel-planar ([c][row][col]: lanes = adjacent
// pixel groups read adjacent 16-byte chunks), the weights as [(t, c)][N] (all lanes of a warp share the channel group:
// broadcast reads).  Per (c, tap row): 2 x-loads + 6 weight loads feed 96 FMAs.  BatchNorm statistics of the stored
// values per thread, reduced in a fixed order at the end (one row of partials per block).
template <int CIN, int N>
struct FnfCfg {
  static constexpr int NG = N / 8;
  static constexpr int TW = (N == 64 || CIN == 64) ? 16 : 32;
  static constexpr int PG = kWnH * TW / 4;              // pixel groups (4 pixels) per tile = threads per channel group
  static constexpr int THREADS = NG * PG;
  static constexpr int LD = TW + 4;                     // row pitch of a channel plane (floats, multiple of 4)
  static constexpr int XS = CIN * (kWnH + 2) * LD, WS = 9 * CIN * N;
  static constexpr int smem = (XS + WS + (THREADS / 32) * 16) * 4;
};

struct FnfParams {
  const float* x;
  const float* wp;               // [N][9 * Cin]
  float* y;
  float* stats_ws;               // [grid][2][N] or null
  long long ld_in, ld_out;
  int H, W, tiles_w, tiles_h, ntiles;
  int tap_of[9];
};

template <int CIN, int N>
__global__ void __launch_bounds__(FnfCfg<CIN, N>::THREADS) fprop_narrow_f32_kernel(const FnfParams p) {
  using Cfg = FnfCfg<CIN, N>;
  constexpr int TW = Cfg::TW, LD = Cfg::LD, NT = Cfg::THREADS, PG = Cfg::PG;
  extern __shared__ __align__(16) float wn_smem[];
  float* xs = wn_smem;                       // [CIN][kWnH + 2][LD], column 0 = image column j0 - 1
  float* ws = wn_smem + Cfg::XS;             // [(a * 3 + c) * CIN + cin][N]
  float* red = ws + Cfg::WS;                 // [warps][16]
  const int ng = threadIdx.x / PG, pg = threadIdx.x % PG;
  const int row = pg / (TW / 4), cg = pg % (TW / 4);
  for (int e = threadIdx.x; e < 9 * CIN * N; e += NT) {
    const int n = e % N, r = e / N;
    const int cin = r % CIN, ac = r / CIN;
    ws[e] = p.wp[(long long)n * 9 * CIN + p.tap_of[ac] * CIN + cin];
  }
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int tj = tile % p.tiles_w, rest = tile / p.tiles_w;
    const int b = rest / p.tiles_h, i0 = (rest % p.tiles_h) * kWnH, j0 = tj * TW;
    __syncthreads();
    for (int e = threadIdx.x; e < (kWnH + 2) * (TW + 2) * (CIN / 4); e += NT) {
      const int c4 = e % (CIN / 4), px = e / (CIN / 4);
      const int r = px / (TW + 2), c = px - r * (TW + 2);
      const int gi = i0 - 1 + r, gj = j0 - 1 + c;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((unsigned)gi < (unsigned)p.H && (unsigned)gj < (unsigned)p.W)
        v = *reinterpret_cast<const float4*>(p.x + ((long long)(b * p.H + gi) * p.W + gj) * p.ld_in + c4 * 4);
      float* dst = xs + ((c4 * 4) * (kWnH + 2) + r) * LD + c;
      dst[0] = v.x; dst[(kWnH + 2) * LD] = v.y; dst[2 * (kWnH + 2) * LD] = v.z; dst[3 * (kWnH + 2) * LD] = v.w;
    }
    __syncthreads();
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
#pragma unroll 1
    for (int cin = 0; cin < CIN; ++cin) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float* xr = xs + (cin * (kWnH + 2) + row + a) * LD + cg * 4;
        const float4 x0 = *reinterpret_cast<const float4*>(xr);
        const float2 x1 = *reinterpret_cast<const float2*>(xr + 4);
        const float xv[6] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* wr = ws + ((a * 3 + c) * CIN + cin) * N + ng * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[i][k] = fmaf(xv[i + c], w[k], acc[i][k]);
        }
      }
    }
    const int gi = i0 + row;
    if (gi < p.H) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gj = j0 + cg * 4 + i;
        if (gj < p.W) {
          float4* dst = reinterpret_cast<float4*>(p.y + ((long long)(b * p.H + gi) * p.W + gj) * p.ld_out + ng * 8);
          dst[0] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          dst[1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
#pragma unroll
          for (int k = 0; k < 8; ++k) { s[k] += acc[i][k]; q[k] = fmaf(acc[i][k], acc[i][k], q[k]); }
        }
      }
    }
  }
  if (p.stats_ws) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;      // a warp lies inside one channel group (PG % 32 == 0)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
      }
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { red[warp * 16 + k] = s[k]; red[warp * 16 + 8 + k] = q[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * N) {
      const int which = threadIdx.x / N, n = threadIdx.x % N;
      constexpr int WPG = PG / 32;                      // warps per channel group
      float v = 0.f;
      for (int w8 = 0; w8 < WPG; ++w8) v += red[((n / 8) * WPG + w8) * 16 + which * 8 + (n % 8)];
      p.stats_ws[((long long)blockIdx.x * 2 + which) * N + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
static bool wn_ch(int c) { return c == 8 || c == 16 || c == 32 || c == 64; }
static bool fnf_ch(int c) { return c == 16 || c == 32 || c == 64; }

static bool wn_shape_ok(const unetb200_gconv_t* d) {
  static const bool off = getenv("UNETB200_NO_SIMT_NARROW") != nullptr;
  if (off || d->dtype != UNETB200_F32) return false;
  if (d->ntaps != 9 || d->in_scale != 1 || d->out_scale != 1 || d->nquad != 1) return false;
  if (d->in_off_y || d->in_off_x || d->out_off_y || d->out_off_x) return false;
  if (d->Hm != d->Hout || d->Wm != d->Wout || d->Hm != d->Hin || d->Wm != d->Win) return false;
  if (!wn_ch(d->Cin) || !wn_ch(d->N) || d->Cin * (d->N / 8) > 256) return false;
  if ((d->ld_in % 4) || (d->ld_out % 4)) return false;
  bool seen[9] = {false};
  for (int t = 0; t < 9; ++t) {
    const int dy = d->tap_dy[t], dx = d->tap_dx[t];
    if (dy < -1 || dy > 1 || dx < -1 || dx > 1 || seen[(dy + 1) * 3 + dx + 1]) return false;
    seen[(dy + 1) * 3 + dx + 1] = true;
  }
  return (long long)d->B * d->Hm * d->Wm < (1LL << 31) - 256;
}

int wgrad_narrow_f32_supported(const unetb200_gconv_t* d, const void* x, const void* gy) {
  if (!wn_shape_ok(d)) return 0;
  if ((x && !aligned16(x)) || (gy && !aligned16(gy))) return 0;
  return 1;
}

template <int CIN, int N>
static int wn_tw() { return WnCfg<CIN, N>::TW; }
template <int CIN, int N>
static int wn_bps() {                       // resident blocks per SM: shared memory and threads
  int by_smem = (200 * 1024) / (WnCfg<CIN, N>::smem + 1024), by_thr = 512 / WnCfg<CIN, N>::THREADS;   // 128 registers per thread
  int b = by_smem < by_thr ? by_smem : by_thr;
  return b < 1 ? 1 : (b > 16 ? 16 : b);
}

#define UB_WN_CASES(X)                                                                                            \
  X(16, 16) X(16, 32) X(16, 64) X(32, 16) X(32, 32) X(32, 64) X(64, 16) X(64, 32)
#define UB_WN8_CASES(X) X(8, 8) X(8, 16) X(8, 32) X(8, 64) X(16, 8) X(32, 8) X(64, 8)

static int wn_grid(const unetb200_gconv_t* d, WnParams* P) {
  int tw = 16, bps = 1;
#define UB_WN_Q(C, NN) if (d->Cin == C && d->N == NN) { tw = wn_tw<C, NN>(); bps = wn_bps<C, NN>(); }
  UB_WN_CASES(UB_WN_Q)
  UB_WN8_CASES(UB_WN_Q)
#undef UB_WN_Q
  P->tiles_w = (d->Wm + tw - 1) / tw;
  P->tiles_h = (d->Hm + kWnH - 1) / kWnH;
  P->ntiles = d->B * P->tiles_w * P->tiles_h;
  const int slots = bps * sm_count();
  return P->ntiles < slots ? P->ntiles : slots;
}

int wgrad_narrow_f32_splits(const unetb200_gconv_t* d) {
  WnParams P;
  return wn_grid(d, &P);
}

template <int CIN, int N>
static int wn_launch(const WnParams& P, int grid, cudaStream_t s) {
  using Cfg = WnCfg<CIN, N>;
  if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(&wgrad_narrow_f32_kernel<CIN, N>), Cfg::smem, "wgrad_narrow_f32 smem attribute"))
    return rc;
  wgrad_narrow_f32_kernel<CIN, N><<<grid, Cfg::THREADS, Cfg::smem, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "wgrad_narrow_f32 launch");
  return 0;
}

int wgrad_narrow_f32(const unetb200_gconv_t* d, const void* x, const void* gy, float* partials, int splits, cudaStream_t s) {
  if (!wgrad_narrow_f32_supported(d, x, gy)) { set_error("wgrad_narrow_f32: unsupported shape"); return UNETB200_E_INVALID; }
  WnParams P;
  memset(&P, 0, sizeof(P));
  P.x = (const float*)x; P.gy = (const float*)gy; P.partials = partials;
  P.ld_in = d->ld_in; P.ld_out = d->ld_out; P.H = d->Hm; P.W = d->Wm;
  for (int t = 0; t < 9; ++t) P.tap_of[(d->tap_dy[t] + 1) * 3 + d->tap_dx[t] + 1] = t;
  const int grid = wn_grid(d, &P);
  if (grid != splits) { set_error("wgrad_narrow_f32: the planned split count is %d, got %d", grid, splits); return UNETB200_E_INVALID; }
#define UB_WN_L(C, NN) if (d->Cin == C && d->N == NN) return wn_launch<C, NN>(P, grid, s);
  UB_WN_CASES(UB_WN_L)
  UB_WN8_CASES(UB_WN_L)
#undef UB_WN_L
  set_error("wgrad_narrow_f32: unsupported channel counts");
  return UNETB200_E_INVALID;
}

// ---- fprop / dgrad host side
int fprop_narrow_f32_supported(const unetb200_gconv_t* d, const void* x, const void* wp, const void* y) {
  if (!wn_shape_ok(d) || !fnf_ch(d->Cin) || !fnf_ch(d->N)) return 0;
  if ((x && !aligned16(x)) || (y && !aligned16(y)) || (wp && (reinterpret_cast<uintptr_t>(wp) & 3))) return 0;
  return 1;
}

template <int CIN, int N>
static int fnf_bps() {
  int by_smem = (200 * 1024) / (FnfCfg<CIN, N>::smem + 1024), by_thr = 1024 / FnfCfg<CIN, N>::THREADS;
  int b = by_smem < by_thr ? by_smem : by_thr;
  return b < 1 ? 1 : (b > 8 ? 8 : b);
}

static int fnf_grid(const unetb200_gconv_t* d, FnfParams* P) {
  int tw = 16, bps = 1;
#define UB_FNF_Q(C, NN) if (d->Cin == C && d->N == NN) { tw = FnfCfg<C, NN>::TW; bps = fnf_bps<C, NN>(); }
  UB_WN_CASES(UB_FNF_Q)
#undef UB_FNF_Q
  P->tiles_w = (d->Wm + tw - 1) / tw;
  P->tiles_h = (d->Hm + kWnH - 1) / kWnH;
  P->ntiles = d->B * P->tiles_w * P->tiles_h;
  const int slots = bps * sm_count();
  return P->ntiles < slots ? P->ntiles : slots;
}

long long fprop_narrow_f32_rows(const unetb200_gconv_t* d) {
  if (!wn_shape_ok(d) || !fnf_ch(d->Cin) || !fnf_ch(d->N)) return 0;
  FnfParams P;
  return fnf_grid(d, &P);
}

template <int CIN, int N>
static int fnf_launch(const FnfParams& P, int grid, cudaStream_t s) {
  using Cfg = FnfCfg<CIN, N>;
  static_assert(Cfg::smem <= 200 * 1024 && Cfg::PG % 32 == 0, "shared memory budget / warp alignment");
  if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(&fprop_narrow_f32_kernel<CIN, N>), Cfg::smem, "fprop_narrow_f32 smem attribute"))
    return rc;
  fprop_narrow_f32_kernel<CIN, N><<<grid, Cfg::THREADS, Cfg::smem, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "fprop_narrow_f32 launch");
  return 0;
}

int fprop_narrow_f32(const unetb200_gconv_t* d, const void* x, const void* wp, void* y, double* stats, float* stats_ws,
                     cudaStream_t s) {
  if (!fprop_narrow_f32_supported(d, x, wp, y)) { set_error("fprop_narrow_f32: unsupported shape"); return UNETB200_E_INVALID; }
  FnfParams P;
  memset(&P, 0, sizeof(P));
  P.x = (const float*)x; P.wp = (const float*)wp; P.y = (float*)y;
  P.stats_ws = stats ? stats_ws : nullptr;
  P.ld_in = d->ld_in; P.ld_out = d->ld_out; P.H = d->Hm; P.W = d->Wm;
  for (int t = 0; t < 9; ++t) P.tap_of[(d->tap_dy[t] + 1) * 3 + d->tap_dx[t] + 1] = t;
  const int grid = fnf_grid(d, &P);
  int rc = UNETB200_E_INVALID;
#define UB_FNF_L(C, NN) if (d->Cin == C && d->N == NN) rc = fnf_launch<C, NN>(P, grid, s);
  UB_WN_CASES(UB_FNF_L)
#undef UB_FNF_L
  if (rc) return rc;
  if (P.stats_ws) return launch_stats_reduce(stats_ws, grid, 2 * d->N, stats, s);
  return 0;
}

}  // namespace ub
