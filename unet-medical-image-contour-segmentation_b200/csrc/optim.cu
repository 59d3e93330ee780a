// Optimizer side of the training step (train.py:80-84,153-159): clip_grad_norm_ + RMSprop(momentum) as two
// multi-tensor kernels instead of PyTorch's ~25 foreach launches over the 64 parameter tensors.
//   1. grad_sqnorm : sum of squares of all gradients -> one double (the squared total norm of clip_grad_norm_)
//   2. rmsprop_step: per element, exactly torch.optim.RMSprop (centered=False):
//        g  = grad * min(1, max_norm / (total_norm + 1e-6))          (clip_grad_norm_, optional)
//        g += weight_decay * w
//        sq = alpha * sq + (1 - alpha) * g * g ;  avg = sqrt(sq) + eps
//        buf = momentum * buf + g / avg ;  w -= lr * buf             (momentum > 0)
//        w -= lr * g / avg                                            (momentum == 0)
// The tensor table travels by value in the kernel parameters (no device-side table to maintain): up to
// kMaxTensors tensors per launch, the host loops over longer lists.  Memory-bound: 4 reads + 3 writes per element.
#include "common.cuh"

namespace ub {

constexpr int kMaxTensors = 48;
constexpr int kChunk = 8192;          // elements per block

struct alignas(16) OptTable {
  float* w[kMaxTensors];
  float* g[kMaxTensors];
  float* sq[kMaxTensors];
  float* mom[kMaxTensors];
  int chunk_start[kMaxTensors + 1];   // prefix sum of chunks per tensor
  long long numel[kMaxTensors];
  int ntensors;
};

__device__ __forceinline__ int find_tensor(const OptTable& t, int chunk) {
  int lo = 0, hi = t.ntensors - 1;
  while (lo < hi) {                   // last tensor with chunk_start <= chunk
    const int mid = (lo + hi + 1) >> 1;
    if (t.chunk_start[mid] <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const __grid_constant__ OptTable t, double* __restrict__ out) {
  const int ti = find_tensor(t, blockIdx.x);
  const long long n = t.numel[ti];
  const long long begin = (long long)(blockIdx.x - t.chunk_start[ti]) * kChunk;
  long long end = begin + kChunk;
  if (end > n) end = n;
  const float* g = t.g[ti];
  float s = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const long long nvec = (end - begin) >> 2;
    for (long long v = threadIdx.x; v < nvec; v += 256) {
      const float4 x = *reinterpret_cast<const float4*>(g + begin + 4 * v);
      s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    for (long long i = begin + 4 * nvec + threadIdx.x; i < end; i += 256) s += g[i] * g[i];
  } else {
    for (long long i = begin + threadIdx.x; i < end; i += 256) s += g[i] * g[i];
  }
  double d = warp_sum((double)s);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a += red[i];
    atomicAdd(out, a);
  }
}

__global__ void __launch_bounds__(256) rmsprop_step_kernel(const __grid_constant__ OptTable t, const double* __restrict__ sumsq,
                                                           float max_norm, float lr, float alpha, float eps,
                                                           float weight_decay, float momentum, int write_clipped_grad) {
  const int ti = find_tensor(t, blockIdx.x);
  const long long n = t.numel[ti];
  const long long begin = (long long)(blockIdx.x - t.chunk_start[ti]) * kChunk;
  long long end = begin + kChunk;
  if (end > n) end = n;
  float coef = 1.f;
  if (sumsq) {
    const float total = (float)sqrt(*sumsq);
    // torch.nn.utils.clip_grad_norm_: clamp(max_norm / (total + 1e-6), max=1); a non-finite norm propagates to
    // every gradient as it does in torch (fminf alone would drop a NaN)
    const float c = max_norm / (total + 1e-6f);
    coef = (c == c) ? fminf(c, 1.f) : c;
  }
  float* w = t.w[ti];
  float* g = t.g[ti];
  float* sq = t.sq[ti];
  float* mom = t.mom[ti];
  const float one_m_alpha = 1.f - alpha;
  auto update = [&](float gi, float wi, float si, float mi, float& w_out, float& s_out, float& m_out, float& g_out) {
    gi *= coef;
    g_out = gi;
    if (weight_decay != 0.f) gi = fmaf(weight_decay, wi, gi);
    const float s = fmaf(one_m_alpha * gi, gi, si * alpha);       // square_avg.mul_(alpha).addcmul_(g, g, value=1-alpha)
    s_out = s;
    const float avg = sqrtf(s) + eps;
    if (mom) {
      const float b = fmaf(momentum, mi, gi / avg);               // buf.mul_(momentum).addcdiv_(g, avg)
      m_out = b;
      w_out = fmaf(-lr, b, wi);
    } else {
      m_out = 0.f;
      w_out = fmaf(-lr, gi / avg, wi);
    }
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(sq) |
                     reinterpret_cast<uintptr_t>(mom)) & 15) == 0;
  long long i0 = begin;
  if (vec) {
    const long long nvec = (end - begin) >> 2;
    for (long long v = threadIdx.x; v < nvec; v += 256) {
      const long long i = begin + 4 * v;
      const float4 g4 = *reinterpret_cast<const float4*>(g + i), w4 = *reinterpret_cast<const float4*>(w + i),
                   s4 = *reinterpret_cast<const float4*>(sq + i);
      const float4 m4 = mom ? *reinterpret_cast<const float4*>(mom + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 wo, so, mo, go;
      update(g4.x, w4.x, s4.x, m4.x, wo.x, so.x, mo.x, go.x);
      update(g4.y, w4.y, s4.y, m4.y, wo.y, so.y, mo.y, go.y);
      update(g4.z, w4.z, s4.z, m4.z, wo.z, so.z, mo.z, go.z);
      update(g4.w, w4.w, s4.w, m4.w, wo.w, so.w, mo.w, go.w);
      *reinterpret_cast<float4*>(w + i) = wo;
      *reinterpret_cast<float4*>(sq + i) = so;
      if (mom) *reinterpret_cast<float4*>(mom + i) = mo;
      if (write_clipped_grad) *reinterpret_cast<float4*>(g + i) = go;
    }
    i0 = begin + 4 * nvec;
  }
  for (long long i = i0 + threadIdx.x; i < end; i += 256) {
    float wo, so, mo, go;
    update(g[i], w[i], sq[i], mom ? mom[i] : 0.f, wo, so, mo, go);
    w[i] = wo;
    sq[i] = so;
    if (mom) mom[i] = mo;
    if (write_clipped_grad) g[i] = go;
  }
}

static int build_table(OptTable* T, float* const* w, float* const* g, float* const* sq, float* const* mom,
                       const int64_t* numel, int first, int count) {
  int chunks = 0;
  for (int i = 0; i < count; ++i) {
    T->w[i] = w ? w[first + i] : nullptr;
    T->g[i] = g[first + i];
    T->sq[i] = sq ? sq[first + i] : nullptr;
    T->mom[i] = mom ? mom[first + i] : nullptr;
    T->numel[i] = numel[first + i];
    T->chunk_start[i] = chunks;
    chunks += (int)((numel[first + i] + kChunk - 1) / kChunk);
  }
  T->chunk_start[count] = chunks;
  T->ntensors = count;
  return chunks;
}

}  // namespace ub

using namespace ub;

extern "C" {

int unetb200_grad_sqnorm(float* const* grads, const int64_t* numel, int ntensors, double* out, void* stream) {
  UB_CHECK_ARG(grads && numel && out && ntensors >= 1, "grad_sqnorm: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), s);
  if (e != cudaSuccess) return cuda_fail(e, "grad_sqnorm memset");
  for (int first = 0; first < ntensors; first += kMaxTensors) {
    const int count = ntensors - first < kMaxTensors ? ntensors - first : kMaxTensors;
    OptTable T;
    for (int i = 0; i < count; ++i) UB_CHECK_ARG(grads[first + i] && numel[first + i] > 0, "grad_sqnorm: tensor %d", first + i);
    const int chunks = build_table(&T, nullptr, grads, nullptr, nullptr, numel, first, count);
    grad_sqnorm_kernel<<<chunks, 256, 0, s>>>(T, out);
  }
  UB_LAUNCH_CHECK("grad_sqnorm");
  return 0;
}

int unetb200_rmsprop_step(float* const* w, float* const* g, float* const* sq, float* const* mom, const int64_t* numel,
                          int ntensors, const double* sumsq, float max_norm, float lr, float alpha, float eps,
                          float weight_decay, float momentum, int write_clipped_grad, void* stream) {
  UB_CHECK_ARG(w && g && sq && numel && ntensors >= 1, "rmsprop_step: bad args");
  UB_CHECK_ARG((momentum != 0.f) == (mom != nullptr), "rmsprop_step: momentum buffers must be given iff momentum != 0");
  cudaStream_t s = (cudaStream_t)stream;
  for (int first = 0; first < ntensors; first += kMaxTensors) {
    const int count = ntensors - first < kMaxTensors ? ntensors - first : kMaxTensors;
    OptTable T;
    for (int i = 0; i < count; ++i)
      UB_CHECK_ARG(w[first + i] && g[first + i] && sq[first + i] && numel[first + i] > 0, "rmsprop_step: tensor %d", first + i);
    const int chunks = build_table(&T, w, g, sq, mom, numel, first, count);
    rmsprop_step_kernel<<<chunks, 256, 0, s>>>(T, sumsq, max_norm, lr, alpha, eps, weight_decay, momentum,
                                               write_clipped_grad);
  }
  UB_LAUNCH_CHECK("rmsprop_step");
  return 0;
}
}
