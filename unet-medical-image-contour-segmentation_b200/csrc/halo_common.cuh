// Helpers shared by the TMA-staged narrow-channel kernels (conv_halo.cu, conv_halo_t.cu).
#pragma once
#include "tc_common.cuh"

namespace ub {

constexpr uint32_t kLayoutSW64 = 4, kLayoutSW32 = 6;
template <int P>
__host__ __device__ constexpr uint32_t halo_layout() { return P == 128 ? kLayoutSW128 : (P == 64 ? kLayoutSW64 : kLayoutSW32); }

int encode_act_box_sw(CUtensorMap* m, const void* base, int C, int W, int H, int B, long long sw, long long sh, long long sb,
                      int box_w, int box_h, int esz = 2);

__device__ __forceinline__ void halo_tmem_ld16(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t halo_pack(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void tma_load_5d(void* smem, const void* desc, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// generic bf16 tiled map with the 32 / 64 / 128-byte swizzle chosen by the innermost box (tc_host.cu)
int encode_bf16_box(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                    const unsigned* box);

}  // namespace ub
