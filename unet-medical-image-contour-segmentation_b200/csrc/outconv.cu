// OutConv: nn.Conv2d(C, n_classes, kernel_size=1) with bias (unet_parts.py:100-106), n_classes <= 8.
// N = 2..4 makes this HBM-bound (one read of the last activation), so it stays on the CUDA cores:
// LPP = C/8 lanes cooperate on one pixel (each owns 8 channels = one 16-byte load), partial dot
// products are combined with warp shuffles.  Backward produces gx, and dW/dbias through per-block
// partials + a deterministic second-stage reduce.
#include "common.cuh"

namespace ub {

constexpr int kMaxK = 8;

// K (= n_classes) is a template parameter so that the class loops carry no predicated-off iterations, and every
// thread keeps UNR pixels' 16-byte loads in flight (the kernel is a pure stream over the last activation).
constexpr int kOutUnr = 4;
template <typename T, int K>
__global__ void __launch_bounds__(256) outconv_fwd_vec_kernel(const T* __restrict__ x, int64_t ld_x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, T* __restrict__ out,
                                                              int64_t npix, int C, int LPP) {
  extern __shared__ float sw[];   // [K][C] rounded to T, then bias[K]
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) sw[i] = Elem<T>::round(w[i]);
  for (int i = threadIdx.x; i < K; i += blockDim.x) sw[K * C + i] = bias ? Elem<T>::round(bias[i]) : 0.f;
  __syncthreads();
  const int sub = threadIdx.x % LPP;
  const int ppb = blockDim.x / LPP;
  const int c0 = sub * 8;
  float wr[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) wr[k][i] = sw[k * C + c0 + i];
  // the trip count is block-uniform so that the shuffles below are executed by full warps
  for (int64_t base = (int64_t)blockIdx.x * ppb * kOutUnr; base < npix; base += (int64_t)gridDim.x * ppb * kOutUnr) {
    float v[kOutUnr][8];
#pragma unroll
    for (int u = 0; u < kOutUnr; ++u) {
      const int64_t p = base + u * ppb + threadIdx.x / LPP;
      if (p < npix) {
        load8(x + p * ld_x + c0, v[u]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[u][i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kOutUnr; ++u) {
      const int64_t p = base + u * ppb + threadIdx.x / LPP;
      float acc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        acc[k] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k] = fmaf(v[u][i], wr[k][i], acc[k]);
      }
      for (int o = LPP >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      }
      if (sub == 0 && p < npix) {
#pragma unroll
        for (int k = 0; k < K; ++k) Elem<T>::st(out + p * K + k, acc[k] + sw[K * C + k]);
      }
    }
  }
}

template <typename T>
__global__ void outconv_fwd_scalar_kernel(const T* __restrict__ x, int64_t ld_x, const float* __restrict__ w,
                                          const float* __restrict__ bias, T* __restrict__ out, int64_t npix, int C,
                                          int K) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    for (int k = 0; k < K; ++k) {
      float a = 0.f;
      for (int c = 0; c < C; ++c) a = fmaf(Elem<T>::ld(x + p * ld_x + c), Elem<T>::round(w[k * C + c]), a);
      Elem<T>::st(out + p * K + k, a + (bias ? Elem<T>::round(bias[k]) : 0.f));
    }
  }
}

// backward, vector path.  partial layout per block: float[K*C] dW then float[K] dbias.
template <typename T, int K>
__global__ void __launch_bounds__(256) outconv_bwd_vec_kernel(const T* __restrict__ x, int64_t ld_x, const float* __restrict__ w,
                                                              const T* __restrict__ g, T* __restrict__ gx, int64_t ld_gx,
                                                              float* __restrict__ partial, int64_t npix, int C, int LPP) {
  extern __shared__ float sm[];   // [K][C] weights; then reduction scratch [K][C] + [K]
  float* sw = sm;
  float* red = sm + K * C;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) sw[i] = Elem<T>::round(w[i]);
  for (int i = threadIdx.x; i < K * C + K; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int sub = threadIdx.x % LPP;
  const int ppb = blockDim.x / LPP;
  const int c0 = sub * 8;
  float wr[K][8], dw[K][8], db[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    db[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { dw[k][i] = 0.f; wr[k][i] = sw[k * C + c0 + i]; }
  }
  for (int64_t base = (int64_t)blockIdx.x * ppb * kOutUnr; base < npix; base += (int64_t)gridDim.x * ppb * kOutUnr) {
    float v[kOutUnr][8], gk[kOutUnr][K];
#pragma unroll
    for (int u = 0; u < kOutUnr; ++u) {
      const int64_t p = base + u * ppb + threadIdx.x / LPP;
      if (p < npix) {
        load8(x + p * ld_x + c0, v[u]);
#pragma unroll
        for (int k = 0; k < K; ++k) gk[u][k] = Elem<T>::ld(g + p * K + k);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[u][i] = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) gk[u][k] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kOutUnr; ++u) {
      const int64_t p = base + u * ppb + threadIdx.x / LPP;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        db[k] += gk[u][k];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o[i] = fmaf(gk[u][k], wr[k][i], o[i]);
          dw[k][i] = fmaf(gk[u][k], v[u][i], dw[k][i]);
        }
      }
      if (gx && p < npix) store8(gx + p * ld_gx + c0, o);
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&red[k * C + c0 + i], dw[k][i]);
    if (sub == 0) atomicAdd(&red[K * C + k], db[k]);
  }
  __syncthreads();
  float* dst = partial + (int64_t)blockIdx.x * (K * C + K);
  for (int i = threadIdx.x; i < K * C + K; i += blockDim.x) dst[i] = red[i];
}

template <typename T>
__global__ void outconv_bwd_scalar_kernel(const T* __restrict__ x, int64_t ld_x, const float* __restrict__ w,
                                          const T* __restrict__ g, T* __restrict__ gx, int64_t ld_gx,
                                          float* __restrict__ partial, int64_t npix, int C, int K) {
  // one block per (k, c) pair group is overkill for the generic path: blocks stride over pixels and
  // accumulate dW through shared-memory atomics.
  extern __shared__ float red[];   // [K*C + K]
  for (int i = threadIdx.x; i < K * C + K; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    for (int c = 0; c < C; ++c) {
      float xv = Elem<T>::ld(x + p * ld_x + c);
      float o = 0.f;
      for (int k = 0; k < K; ++k) {
        float gk = Elem<T>::ld(g + p * K + k);
        o = fmaf(gk, Elem<T>::round(w[k * C + c]), o);
        atomicAdd(&red[k * C + c], gk * xv);
      }
      if (gx) Elem<T>::st(gx + p * ld_gx + c, o);
    }
    for (int k = 0; k < K; ++k) atomicAdd(&red[K * C + k], Elem<T>::ld(g + p * K + k));
  }
  __syncthreads();
  float* dst = partial + (int64_t)blockIdx.x * (K * C + K);
  for (int i = threadIdx.x; i < K * C + K; i += blockDim.x) dst[i] = red[i];
}

__global__ void outconv_reduce_kernel(const float* __restrict__ partial, int nblocks, int KC, int K, float* dw,
                                      float* dbias) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= KC + K) return;
  double s = 0;
  for (int b = 0; b < nblocks; ++b) s += partial[(int64_t)b * (KC + K) + i];
  if (i < KC) dw[i] = (float)s;
  else if (dbias) dbias[i - KC] = (float)s;
}

static int outconv_blocks(int64_t npix) {
  int64_t b = (npix + 2047) / 2048;
  int64_t cap = (int64_t)sm_count() * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
static bool outconv_vec(int C, int64_t ld, const void* p, size_t esz) {
  int l = C / 8;
  return C % 8 == 0 && l >= 1 && l <= 32 && (l & (l - 1)) == 0 && ld % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(p) % (8 * esz)) == 0;
}

}  // namespace ub

using namespace ub;
typedef __nv_bfloat16 bf16;

extern "C" {

int unetb200_outconv_fwd(const void* x, int64_t ld_x, const float* w, const float* bias, void* logits, int dtype,
                         int64_t npix, int C, int ncls, void* stream) {
  UB_CHECK_ARG(dtype == UNETB200_F32 || dtype == UNETB200_BF16, "outconv_fwd: dtype");
  UB_CHECK_ARG(npix > 0 && C > 0 && ncls >= 1 && ncls <= kMaxK && ld_x >= C,
               "outconv_fwd: npix=%lld C=%d n_classes=%d (n_classes <= %d)", (long long)npix, C, ncls, kMaxK);
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = outconv_blocks(npix);
  size_t esz = dtype == UNETB200_BF16 ? 2 : 4;
  size_t smem = sizeof(float) * ((size_t)ncls * C + ncls);
  if (outconv_vec(C, ld_x, x, esz) && smem <= 48 * 1024) {
#define UB_OUTCONV_FWD(KK)                                                                                           \
  case KK:                                                                                                           \
    if (dtype == UNETB200_BF16)                                                                                      \
      outconv_fwd_vec_kernel<bf16, KK><<<blocks, 256, smem, s>>>((const bf16*)x, ld_x, w, bias, (bf16*)logits, npix, C, C / 8); \
    else                                                                                                             \
      outconv_fwd_vec_kernel<float, KK><<<blocks, 256, smem, s>>>((const float*)x, ld_x, w, bias, (float*)logits, npix, C, C / 8); \
    break;
    switch (ncls) {
      UB_OUTCONV_FWD(1) UB_OUTCONV_FWD(2) UB_OUTCONV_FWD(3) UB_OUTCONV_FWD(4)
      UB_OUTCONV_FWD(5) UB_OUTCONV_FWD(6) UB_OUTCONV_FWD(7) UB_OUTCONV_FWD(8)
    }
#undef UB_OUTCONV_FWD
  } else {
    if (dtype == UNETB200_BF16)
      outconv_fwd_scalar_kernel<bf16><<<blocks, 256, 0, s>>>((const bf16*)x, ld_x, w, bias, (bf16*)logits, npix, C,
                                                             ncls);
    else
      outconv_fwd_scalar_kernel<float><<<blocks, 256, 0, s>>>((const float*)x, ld_x, w, bias, (float*)logits, npix,
                                                              C, ncls);
  }
  UB_LAUNCH_CHECK("outconv_fwd");
  return 0;
}

int64_t unetb200_outconv_bwd_workspace(int64_t npix, int C, int ncls) {
  return (int64_t)148 * 4 * 2 * ((int64_t)ncls * C + ncls) + 64 + 0 * npix;   // floats; >= blocks * (K*C+K)
}

int unetb200_outconv_bwd(const void* x, int64_t ld_x, const float* w, const void* glogits, void* gx, int64_t ld_gx,
                         float* dw, float* dbias, float* workspace, int dtype, int64_t npix, int C, int ncls,
                         void* stream) {
  UB_CHECK_ARG(dtype == UNETB200_F32 || dtype == UNETB200_BF16, "outconv_bwd: dtype");
  UB_CHECK_ARG(npix > 0 && C > 0 && ncls >= 1 && ncls <= kMaxK && ld_x >= C, "outconv_bwd: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = outconv_blocks(npix);
  if (blocks > 148 * 4 * 2) blocks = 148 * 4 * 2;
  size_t esz = dtype == UNETB200_BF16 ? 2 : 4;
  const int KC = ncls * C;
  size_t smem_v = sizeof(float) * (2 * (size_t)KC + ncls);
  size_t smem_s = sizeof(float) * ((size_t)KC + ncls);
  UB_CHECK_ARG(smem_s <= 48 * 1024, "outconv_bwd: n_classes*C too large (%d)", KC);
  bool vec = outconv_vec(C, ld_x, x, esz) && (!gx || outconv_vec(C, ld_gx, gx, esz)) && smem_v <= 48 * 1024;
  if (vec) {
#define UB_OUTCONV_BWD(KK)                                                                                           \
  case KK:                                                                                                           \
    if (dtype == UNETB200_BF16)                                                                                      \
      outconv_bwd_vec_kernel<bf16, KK><<<blocks, 256, smem_v, s>>>((const bf16*)x, ld_x, w, (const bf16*)glogits, (bf16*)gx, \
                                                                   ld_gx, workspace, npix, C, C / 8);                 \
    else                                                                                                             \
      outconv_bwd_vec_kernel<float, KK><<<blocks, 256, smem_v, s>>>((const float*)x, ld_x, w, (const float*)glogits,  \
                                                                    (float*)gx, ld_gx, workspace, npix, C, C / 8);    \
    break;
    switch (ncls) {
      UB_OUTCONV_BWD(1) UB_OUTCONV_BWD(2) UB_OUTCONV_BWD(3) UB_OUTCONV_BWD(4)
      UB_OUTCONV_BWD(5) UB_OUTCONV_BWD(6) UB_OUTCONV_BWD(7) UB_OUTCONV_BWD(8)
    }
#undef UB_OUTCONV_BWD
  } else {
    if (dtype == UNETB200_BF16)
      outconv_bwd_scalar_kernel<bf16><<<blocks, 256, smem_s, s>>>((const bf16*)x, ld_x, w, (const bf16*)glogits,
                                                                  (bf16*)gx, ld_gx, workspace, npix, C, ncls);
    else
      outconv_bwd_scalar_kernel<float><<<blocks, 256, smem_s, s>>>((const float*)x, ld_x, w, (const float*)glogits,
                                                                   (float*)gx, ld_gx, workspace, npix, C, ncls);
  }
  outconv_reduce_kernel<<<(KC + ncls + 127) / 128, 128, 0, s>>>(workspace, blocks, KC, ncls, dw, dbias);
  UB_LAUNCH_CHECK("outconv_bwd");
  return 0;
}
}
