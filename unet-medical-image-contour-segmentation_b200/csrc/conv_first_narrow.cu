// First layer of the reference's light variants (UNet_S / UNet_T / UNet_SA, unet_model.py:52-189): Conv2d(1, N, 3,
// padding=1, bias=False) with N = 8 / 16 / 32 output channels in bf16 -- fprop (+ BatchNorm batch statistics, or the
// folded eval-mode BatchNorm + ReLU) and wgrad.
//
// With one input channel there is nothing for a tensor core to contract over (K = 9): the work is 9 N FMAs per pixel
// against 2 N bytes of HBM traffic, i.e. HBM-bound on the CUDA cores if nothing else gets in the way.  The im2col
// kernels (conv_narrow.cu, K padded to 16) spent 0.15 / 0.37 ms on it where the HBM floor is 0.02 ms.  Here a block
// owns a (256 / G) x 8 pixel tile (G = N / 8 channel groups): the input halo sits in shared memory as fp32, a thread
// owns one pixel column and one group of 8 channels, keeps its 3 x 3 window in registers while walking down the 8
// rows, and moves dY / Y as one 16-byte vector per pixel.  Statistics and weight-gradient partials are reduced in a
// fixed order (shuffles inside a warp, then warp 0..7 in order): bit-reproducible, no atomics.
#include <cstring>

#include "gconv.cuh"

namespace ub {

constexpr int kFnH = 8;

template <int G>
struct FnCfg {
  static constexpr int TW = 256 / G, LD = TW + 4;
};

__device__ __forceinline__ void fn_unpack8(const uint4& r, float v[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint32_t fn_pack(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

struct FnParams {
  const __nv_bfloat16* x;        // [B][H][W][ld_in], channel 0
  const __nv_bfloat16* w;        // packed [N][9] (fprop)
  __nv_bfloat16* y;              // fprop output / wgrad dY  [B][H][W][ld_out]
  const float* affine;           // MODE 1: scale[N] then shift[N]
  float* ws;                     // fprop: stats rows [grid][2][N]; wgrad: partials [grid][9][N]
  long long ld_in, ld_out;
  int H, W, N;
  int tiles_w, tiles_h, ntiles;
  int tap_of[9];                 // [(dy + 1) * 3 + dx + 1] -> tap index of the descriptor
};

template <int G>
__device__ __forceinline__ void fn_load_halo(const FnParams& p, float (*halo)[FnCfg<G>::LD], int b, int i0, int j0) {
  constexpr int TW = FnCfg<G>::TW;
  for (int e = threadIdx.x; e < (kFnH + 2) * (TW + 2); e += 256) {
    const int r = e / (TW + 2), c = e - r * (TW + 2);
    const int gi = i0 - 1 + r, gj = j0 - 1 + c;
    float v = 0.f;
    if ((unsigned)gi < (unsigned)p.H && (unsigned)gj < (unsigned)p.W)
      v = __bfloat162float(p.x[((long long)(b * p.H + gi) * p.W + gj) * p.ld_in]);
    halo[r][c] = v;
  }
}

// MODE 0: y = conv(x), statistics of the rounded y;  MODE 1: y = relu(conv(x) * scale + shift)
template <int G, int MODE>
__global__ void __launch_bounds__(256, 2) first_narrow_fprop_kernel(const FnParams p) {
  constexpr int TW = FnCfg<G>::TW;
  __shared__ float halo[kFnH + 2][FnCfg<G>::LD];
  __shared__ float red[8][2 * 8 * G];
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  float wt[3][3][8];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int t = p.tap_of[a * 3 + c];
#pragma unroll
      for (int k = 0; k < 8; ++k) wt[a][c][k] = __bfloat162float(p.w[(g * 8 + k) * 9 + t]);
    }
  float sc[8], sh[8], s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s[k] = q[k] = 0.f;
    sc[k] = MODE == 1 ? p.affine[g * 8 + k] : 1.f;
    sh[k] = MODE == 1 ? p.affine[p.N + g * 8 + k] : 0.f;
  }
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int tj = tile % p.tiles_w, rest = tile / p.tiles_w;
    const int b = rest / p.tiles_h, i0 = (rest % p.tiles_h) * kFnH, j0 = tj * TW;
    __syncthreads();
    fn_load_halo<G>(p, halo, b, i0, j0);
    __syncthreads();
    const int j = j0 + pl;
    float win[3][3];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) win[a][c] = halo[a][pl + c];
#pragma unroll
    for (int r = 0; r < kFnH; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) win[2][c] = halo[r + 2][pl + c];
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = fmaf(win[a][c], wt[a][c][k], o[k]);
      if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaxf(fmaf(o[k], sc[k], sh[k]), 0.f);
      }
      const int i = i0 + r;
      if (i < p.H && j < p.W) {
        uint4 pk;
        pk.x = fn_pack(o[0], o[1]); pk.y = fn_pack(o[2], o[3]); pk.z = fn_pack(o[4], o[5]); pk.w = fn_pack(o[6], o[7]);
        *reinterpret_cast<uint4*>(p.y + ((long long)(b * p.H + i) * p.W + j) * p.ld_out + g * 8) = pk;
        if (MODE == 0) {
          float rv[8];
          fn_unpack8(pk, rv);
#pragma unroll
          for (int k = 0; k < 8; ++k) { s[k] += rv[k]; q[k] = fmaf(rv[k], rv[k], q[k]); }
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) { win[0][c] = win[1][c]; win[1][c] = win[2][c]; }
    }
  }
  if (MODE == 0 && p.ws) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int o = G; o < 32; o <<= 1) {               // lanes with the same channel group
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
      }
      if (lane < G) { red[warp][g * 8 + k] = s[k]; red[warp][8 * G + g * 8 + k] = q[k]; }
    }
    __syncthreads();
    float* out = p.ws + (long long)blockIdx.x * 2 * p.N;
    for (int e = threadIdx.x; e < 2 * 8 * G; e += 256) {
      float v = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) v += red[w8][e];
      out[e] = v;
    }
  }
}

template <int G>
__global__ void __launch_bounds__(256, 2) first_narrow_wgrad_kernel(const FnParams p) {
  constexpr int TW = FnCfg<G>::TW;
  __shared__ float halo[kFnH + 2][FnCfg<G>::LD];
  __shared__ float wred[8][9 * 8 * G];
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  float acc[3][3][8];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[a][c][k] = 0.f;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int tj = tile % p.tiles_w, rest = tile / p.tiles_w;
    const int b = rest / p.tiles_h, i0 = (rest % p.tiles_h) * kFnH, j0 = tj * TW;
    const int j = j0 + pl;
    // the eight dY vectors of this thread's pixel column are in flight while the halo is staged
    uint4 gv[kFnH];
#pragma unroll
    for (int r = 0; r < kFnH; ++r) {
      const int i = i0 + r;
      gv[r] = make_uint4(0u, 0u, 0u, 0u);
      if (i < p.H && j < p.W) gv[r] = *reinterpret_cast<const uint4*>(p.y + ((long long)(b * p.H + i) * p.W + j) * p.ld_out + g * 8);
    }
    __syncthreads();
    fn_load_halo<G>(p, halo, b, i0, j0);
    __syncthreads();
    float win[3][3];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) win[a][c] = halo[a][pl + c];
#pragma unroll
    for (int r = 0; r < kFnH; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) win[2][c] = halo[r + 2][pl + c];
      float v[8];
      fn_unpack8(gv[r], v);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[a][c][k] = fmaf(win[a][c], v[k], acc[a][c][k]);
#pragma unroll
      for (int c = 0; c < 3; ++c) { win[0][c] = win[1][c]; win[1][c] = win[2][c]; }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int t = p.tap_of[a * 3 + c];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = acc[a][c][k];
#pragma unroll
        for (int o = G; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < G) wred[warp][t * 8 * G + g * 8 + k] = v;
      }
    }
  __syncthreads();
  float* out = p.ws + (long long)blockIdx.x * 9 * p.N;
  for (int e = threadIdx.x; e < 9 * 8 * G; e += 256) {
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) v += wred[w8][e];
    out[e] = v;
  }
}

// ------------------------------------------------------------------------------------------ host side
static bool fn_shape_ok(const unetb200_gconv_t* d) {
  static const bool off = getenv("UNETB200_NO_FIRST_NARROW") != nullptr;
  if (off || d->dtype != UNETB200_BF16) return false;
  if (d->Cin != 1 || (d->N != 8 && d->N != 16 && d->N != 32)) return false;
  if (d->ntaps != 9 || d->in_scale != 1 || d->out_scale != 1 || d->nquad != 1) return false;
  if (d->in_off_y || d->in_off_x || d->out_off_y || d->out_off_x) return false;
  if (d->Hm != d->Hout || d->Wm != d->Wout || d->Hm != d->Hin || d->Wm != d->Win) return false;
  if (d->ld_out % 8) return false;
  bool seen[9] = {false};
  for (int t = 0; t < 9; ++t) {
    const int dy = d->tap_dy[t], dx = d->tap_dx[t];
    if (dy < -1 || dy > 1 || dx < -1 || dx > 1 || seen[(dy + 1) * 3 + dx + 1]) return false;
    seen[(dy + 1) * 3 + dx + 1] = true;
  }
  return (long long)d->B * d->Hm * d->Wm < (1LL << 31) - 256;
}

int first_narrow_supported(const unetb200_gconv_t* d, const void* y) {
  return fn_shape_ok(d) && (!y || aligned16(y)) ? 1 : 0;
}

static int fn_grid(const unetb200_gconv_t* d, FnParams* P) {
  const int tw = 256 / (d->N / 8);
  P->tiles_w = (d->Wm + tw - 1) / tw;
  P->tiles_h = (d->Hm + kFnH - 1) / kFnH;
  P->ntiles = d->B * P->tiles_w * P->tiles_h;
  const int slots = 2 * sm_count();
  return P->ntiles < slots ? P->ntiles : slots;
}

long long first_narrow_rows(const unetb200_gconv_t* d) {
  if (!fn_shape_ok(d)) return 0;
  FnParams P;
  return fn_grid(d, &P);
}

static void fn_fill(const unetb200_gconv_t* d, FnParams* P) {
  memset(P, 0, sizeof(*P));
  P->ld_in = d->ld_in; P->ld_out = d->ld_out;
  P->H = d->Hm; P->W = d->Wm; P->N = d->N;
  for (int t = 0; t < 9; ++t) P->tap_of[(d->tap_dy[t] + 1) * 3 + d->tap_dx[t] + 1] = t;
}

int first_narrow_fprop(const unetb200_gconv_t* d, const void* x, const void* wp, void* y, double* stats, float* stats_ws,
                       const float* affine, cudaStream_t s) {
  if (!first_narrow_supported(d, y)) { set_error("first_narrow_fprop: unsupported shape"); return UNETB200_E_INVALID; }
  FnParams P;
  fn_fill(d, &P);
  P.x = (const __nv_bfloat16*)x; P.w = (const __nv_bfloat16*)wp; P.y = (__nv_bfloat16*)y;
  P.affine = affine;
  P.ws = (stats && !affine) ? stats_ws : nullptr;
  const int grid = fn_grid(d, &P);
  const int G = d->N / 8;
  if (affine) {
    if (G == 1) first_narrow_fprop_kernel<1, 1><<<grid, 256, 0, s>>>(P);
    else if (G == 2) first_narrow_fprop_kernel<2, 1><<<grid, 256, 0, s>>>(P);
    else first_narrow_fprop_kernel<4, 1><<<grid, 256, 0, s>>>(P);
  } else {
    if (G == 1) first_narrow_fprop_kernel<1, 0><<<grid, 256, 0, s>>>(P);
    else if (G == 2) first_narrow_fprop_kernel<2, 0><<<grid, 256, 0, s>>>(P);
    else first_narrow_fprop_kernel<4, 0><<<grid, 256, 0, s>>>(P);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "first_narrow_fprop launch");
  if (P.ws) return launch_stats_reduce(stats_ws, grid, 2 * d->N, stats, s);
  return 0;
}

int first_narrow_wgrad_splits(const unetb200_gconv_t* d) {
  FnParams P;
  return fn_grid(d, &P);
}

int first_narrow_wgrad(const unetb200_gconv_t* d, const void* x, const void* gy, float* partials, int splits, cudaStream_t s) {
  if (!first_narrow_supported(d, gy)) { set_error("first_narrow_wgrad: unsupported shape"); return UNETB200_E_INVALID; }
  FnParams P;
  fn_fill(d, &P);
  P.x = (const __nv_bfloat16*)x; P.y = (__nv_bfloat16*)const_cast<void*>(gy); P.ws = partials;
  const int grid = fn_grid(d, &P);
  if (grid != splits) { set_error("first_narrow_wgrad: the planned split count is %d, got %d", grid, splits); return UNETB200_E_INVALID; }
  const int G = d->N / 8;
  if (G == 1) first_narrow_wgrad_kernel<1><<<grid, 256, 0, s>>>(P);
  else if (G == 2) first_narrow_wgrad_kernel<2><<<grid, 256, 0, s>>>(P);
  else first_narrow_wgrad_kernel<4><<<grid, 256, 0, s>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "first_narrow_wgrad launch");
  return 0;
}

}  // namespace ub
