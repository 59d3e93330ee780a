// Third-generation tcgen05 engine for the 3x3 convolutions (fprop and dgrad of nn.Conv2d(k=3, padding=1),
// unet_parts.py:15,18 and their autograd dgrad): CTA pairs (cta_group::2) + full 3x3 tap reuse.
//
// What bounded the second generation (conv_tc2.cu), measured with tools/umma_probe.cu and ncu:
//   * a 1-CTA tcgen05.mma of M=128 runs at 77 / 87 / 128 cycles for N = 64 / 128 / 256 (floor 32 / 64 / 128):
//     the A operand is re-read from shared memory for every MMA.  As a CTA pair (M = 256, each SM reads its
//     own A and HALF of B) the same instructions take 49 / 64 / 128 cycles;
//   * the L2 -> shared-memory fill stream ran at 7-8 TB/s chip-wide, its ceiling.  Per 64-channel chunk and
//     256-pixel tile the old kernel fetched 3 activation boxes (one per column shift, 108 KB) + 9 weight
//     tiles; here ONE halo box {128 B, tile_w + 2, tile_h + 2} (41.5 KB) serves all nine taps -- the UMMA
//     descriptor of tap (dy, dx) starts ((dy+1) * box_w + (dx+1)) * 128 B into the box, its 8-row groups are
//     box_w * 128 B apart (SBO), and the 128-byte swizzle is a function of the absolute shared-memory
//     address, so neither needs 1024-byte alignment (tools/umma_probe.cu check (b)) -- and each CTA of the
//     pair fetches half of every weight tile.
//
// Structure per CTA (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only) + TMEM owner,
// warps 2..9 = epilogue (as in conv_tc2.cu: tcgen05.ld -> +bias -> round -> swizzled staging -> TMA store,
// BatchNorm partial sums in registers).  Each CTA of a pair owns one 256-pixel tile (two 128-row
// accumulators); the pair shares the n-block.  Barriers: the `full` barriers live in the leader and count
// the bytes of both CTAs' TMA loads (the peer's loads signal the leader's barrier); MMA completion is
// multicast to the `empty` barriers of both CTAs; the epilogue warps of both CTAs arrive on the leader's
// `t_empty`.
#include <cstring>

#include "tc_common.cuh"

namespace ub {

// ---------------------------------------------------------------------------- cluster / cta_group::2 PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int CG>
__device__ __forceinline__ void tma_load_4d_cg(void* smem, const void* desc, uint32_t bar_addr, int c0, int c1, int c2,
                                               int c3) {
  if constexpr (CG == 2) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tma_load_2d_cg(void* smem, const void* desc, uint32_t bar_addr, int c0, int c1) {
  if constexpr (CG == 2) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* dst_smem, uint32_t ncols) {
  if constexpr (CG == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    tmem_alloc(dst_smem, ncols);
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  } else {
    tmem_dealloc(taddr, ncols);
  }
}
template <int CG, bool KIND_TF32>
__device__ __forceinline__ void umma_cg(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 2) {
    if constexpr (KIND_TF32) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
          : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
          : "memory");
    }
  } else {
    umma<KIND_TF32>(d_tmem, a_desc, b_desc, idesc, acc);
  }
}
// MMA completion -> mbarrier at the same shared-memory offset in every CTA of the pair
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint64_t* bar) {
  if constexpr (CG == 2) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
  } else {
    umma_commit(bar);
  }
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t v[32]) {
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
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld_fence() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

struct alignas(64) Tc3Params {
  CUtensorMap a_map;         // source view, box {128 B, box_w, box_h, 1}
  CUtensorMap b_map;         // packed weights [N][K], box {128 B, BLOCK_N / CG}
  CUtensorMap o_map;         // destination view, box {128 B, 8, 4, 1}
  uint32_t tap_aoff[9];      // (((dy + 1) * box_w + (dx + 1)) * 128) >> 4
  uint32_t a_box_bytes;
  uint32_t sub1_off;         // byte offset >> 4 of sub-tile 1's first pixel inside the box
  uint32_t row_bytes;        // box_w * 128 = SBO
  int cchunks, Cin;
  int tiles_w, tiles_h, tile_w, tile_h;
  int sub1_di, sub1_dj;
  int m_tiles, m_groups, total_groups;   // m_groups = ceil(m_tiles / CG); one "group" = CG tiles of one n-block
  int Hm, Wm, B;
  float* stats_ws;           // [n_block][cta][8 epilogue warps][2][BLOCK_N]
  const float* affine;       // AFFINE kernels only: scale[N] then shift[N] (eval-mode BatchNorm folded into the epilogue)
  int N;
  // BNBWD kernels only (dgrad whose output is the gradient of z = relu(bn(yprev))): the raw conv output of the
  // previous layer and its BatchNorm coefficients [4][N] = mean, invstd, scale, shift
  CUtensorMap y_map;         // yprev view, box {128 B, 8, 4, 1} (the geometry of o_map)
  const float* bnc;
  // AFFINE_OUT kernels only: OutConv weights [ncls][64] / bias [ncls] (fp32, rounded to bf16 like the stand-alone
  // kernel), destination logits [B][Hm][Wm][ncls] (bf16)
  const float* oc_w;
  const float* oc_b;
  void* logits;
  int ncls;
  // AFFINE bf16 kernels, optional: MaxPool2d(2) of the activation as a second output (Down's pool reads what this
  // epilogue just staged: a warp's 4 x 8 pixel patch holds whole 2 x 2 windows), [B][Hm/2][Wm/2][ld_pool] bf16
  void* pooled;
  long long ld_pool;
};

constexpr uint32_t kA3Stage = 44032;    // 43 KB >= the largest halo box: 34 rows x 10 px x 128 B
constexpr int kEpi3Stage = 4096;        // 32 rows x 128 B per epilogue warp
constexpr int kEpi3Warps = 8;           // two per TMEM lane quadrant: one per 128-row sub-tile
constexpr int kTc3Threads = 64 + 32 * kEpi3Warps;

// AFFINE: inference form of conv -> BatchNorm(running statistics) -> ReLU (unet_parts.py:15-20 under .eval()): the
// epilogue stores relu(acc * scale[n] + shift[n]) instead of the raw convolution, so the activation is written once
// and the separate 4 B/element BatchNorm pass disappears.  A separate instantiation: the training kernels are unchanged.
//
// EPI_BNBWD: the convolution is a dgrad whose output g is the gradient of z = relu(bn(yprev)) of the previous layer
// (DoubleConv: conv2's dgrad feeds conv1's BatchNorm + ReLU backward).  The epilogue, which holds the rounded g tile
// anyway, also reads the matching yprev tile and accumulates sum(g * mask) and sum(g * mask * xhat) per channel --
// the reduction pass of the BatchNorm backward (unetb200_bn_relu_bwd_reduce: one read of g and of yprev, 4 B per
// element) disappears.  Partial sums take the same route as the forward statistics (per-warp registers -> workspace
// -> fp64 reduce, fixed order).
//
// EPI_AFFINE_OUT: the last DoubleConv of the network in inference (N = 64 = all channels of the layer in one lane's
// row) followed by OutConv (nn.Conv2d(64, n_classes, 1) + bias, unet_parts.py:100-106): each lane holds a pixel's 64
// activated channels, so it also takes the n_classes dot products (weights as [channel][8 classes] in shared memory:
// one warp-uniform 16/32-byte load per channel) and stores the logits; the last activation is never written (one
// 2 B x 64 write and one read of it per pixel, and the OutConv pass, disappear).
constexpr int EPI_PLAIN = 0, EPI_AFFINE = 1, EPI_BNBWD = 2, EPI_AFFINE_OUT = 3;
template <typename T, int BLOCK_N, int SA, int SB, int ACC, int CG, int TPS, int EPI = EPI_PLAIN>
__global__ void __launch_bounds__(kTc3Threads, 1) tc3_conv_kernel(const __grid_constant__ Tc3Params p) {
  constexpr bool OUTC = EPI == EPI_AFFINE_OUT, AFFINE = EPI == EPI_AFFINE || OUTC, BNBWD = EPI == EPI_BNBWD;
  static_assert(!OUTC || (BLOCK_N == 64 && sizeof(T) == 2), "the OutConv epilogue is the bf16 N = 64 kernel");
  constexpr bool TF32 = sizeof(T) == 4;
  constexpr int EPR = 128 / sizeof(T);
  constexpr uint32_t kBTap = (BLOCK_N / CG) * 128;     // one tap's weight tile (this CTA's half of it)
  constexpr uint32_t kBStage = TPS * kBTap;            // a B stage holds TPS taps: one barrier round trip per TPS * 8 MMAs
  static_assert(9 % TPS == 0, "taps per stage");
  constexpr int NCB = BLOCK_N / EPR;                   // 128-byte channel blocks per accumulator row
  static_assert(2 * BLOCK_N * ACC <= 512, "TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  uint8_t* b_ring = a_ring + SA * kA3Stage;
  uint8_t* epi = b_ring + SB * kBStage;                // 8 warps x 4 KB
  uint8_t* yst = epi + kEpi3Warps * kEpi3Stage;        // BNBWD: 8 warps x 4 KB, the yprev patch of the warp's current block
  uint64_t* bars = reinterpret_cast<uint64_t*>(yst + (BNBWD ? kEpi3Warps * kEpi3Stage : 0));
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + SA;
  uint64_t* b_full = a_empty + SA;
  uint64_t* b_empty = b_full + SB;
  uint64_t* t_full = b_empty + SB;
  uint64_t* t_empty = t_full + ACC;
  uint64_t* y_full = t_empty + ACC;                    // BNBWD: one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_full + kEpi3Warps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int group0 = blockIdx.x / CG, ngroups = gridDim.x / CG;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map);
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&p.o_map);
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < ACC; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], kEpi3Warps * CG); }
    if constexpr (BNBWD) {
      tma_prefetch_desc(&p.y_map);
      for (int s = 0; s < kEpi3Warps; ++s) mbar_init(&y_full[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_cg<CG>(tmem_slot, 512);
  if constexpr (OUTC) {
    // OutConv weights as [channel][8 classes] + bias[8] at the start of the (otherwise unused) staging area
    float* ocs = reinterpret_cast<float*>(epi);
    for (int i = threadIdx.x; i < 64 * 8 + 8; i += blockDim.x) {
      const int c = i >> 3, k = i & 7;
      float v = 0.f;
      if (k < p.ncls) v = c < 64 ? __ldg(p.oc_w + k * 64 + c) : (p.oc_b ? __ldg(p.oc_b + k) : 0.f);
      ocs[i] = __bfloat162float(__float2bfloat16_rn(v));
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();           // barrier inits + TMEM allocation of both CTAs are in place
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------ TMA producer (both CTAs) ------------------------
    uint32_t sa = 0, pa = 1, sb = 0, pb = 1;             // stage index, parity to wait for on *_empty
    for (int gt = group0; gt < p.total_groups; gt += ngroups) {
      const int nb = gt / p.m_groups;
      int mt = (gt - nb * p.m_groups) * CG + (int)rank;
      const bool valid = mt < p.m_tiles;
      const int tj = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int ti = mt % p.tiles_h;
      const int b = valid ? mt / p.tiles_h : p.B;       // a tile past the end reads zeros (TMA out-of-bounds fill)
      const int i0 = ti * p.tile_h - 1, j0 = tj * p.tile_w - 1;
      const int n0 = nb * BLOCK_N + (int)rank * (BLOCK_N / CG);
      for (int c = 0; c < p.cchunks; ++c) {
        mbar_wait(&a_empty[sa], pa);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&a_full[sa], CG * p.a_box_bytes);
          tma_load_4d_cg<CG>(a_ring + sa * kA3Stage, &p.a_map, mapa_u32(smem_u32(&a_full[sa]), 0), c * EPR, j0, i0, b);
        }
        if (++sa == SA) { sa = 0; pa ^= 1; }
        for (int t = 0; t < 9; t += TPS) {
          mbar_wait(&b_empty[sb], pb);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&b_full[sb], CG * kBStage);
            const uint32_t bar = mapa_u32(smem_u32(&b_full[sb]), 0);
#pragma unroll
            for (int j = 0; j < TPS; ++j)
              tma_load_2d_cg<CG>(b_ring + sb * kBStage + j * kBTap, &p.b_map, bar, (t + j) * p.Cin + c * EPR, n0);
          }
          if (++sb == SB) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer (leader CTA) --------------------------
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(TF32, false, false, 128 * CG, BLOCK_N);
      const uint64_t a_desc_t = make_desc(smem_u32(a_ring), 16, p.row_bytes);
      const uint64_t b_desc_t = make_desc(smem_u32(b_ring), 16, 1024);
      const uint32_t sub1 = p.sub1_off;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, pacc = 1;
      for (int gt = group0; gt < p.total_groups; gt += ngroups) {
        mbar_wait(&t_empty[acc], pacc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 2 * BLOCK_N;
        uint32_t accum = 0;
        for (int c = 0; c < p.cchunks; ++c) {
          mbar_wait(&a_full[sa], pa);
          const uint64_t a_desc0 = a_desc_t + sa * (kA3Stage >> 4);
          for (int t = 0; t < 9; t += TPS) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint64_t b_desc0 = b_desc_t + sb * (kBStage >> 4);
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < TPS; ++j) {
                const uint64_t a_desc1 = a_desc0 + p.tap_aoff[t + j];
                const uint64_t b_desc1 = b_desc0 + j * (kBTap >> 4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const uint32_t on = accum | (uint32_t)(j | kk);
                  umma_cg<CG, TF32>(d_tmem, a_desc1 + 2 * kk, b_desc1 + 2 * kk, idesc, on);
                  umma_cg<CG, TF32>(d_tmem + BLOCK_N, a_desc1 + sub1 + 2 * kk, b_desc1 + 2 * kk, idesc, on);
                }
              }
              umma_commit_cg<CG>(&b_empty[sb]);
              if (t + TPS == 9) umma_commit_cg<CG>(&a_empty[sa]);
            }
            __syncwarp();
            accum = 1;
            if (++sb == SB) { sb = 0; pb ^= 1; }
          }
          if (++sa == SA) { sa = 0; pa ^= 1; }
        }
        if (elect_one()) umma_commit_cg<CG>(&t_full[acc]);
        __syncwarp();
        if (++acc == ACC) { acc = 0; pacc ^= 1; }
      }
    }
  } else {
    // ------------------------------------ epilogue (both CTAs) -----------------------------
    // 8 warps: warp handles TMEM lane quadrant (warp % 4) of sub-tile (warp - 2) / 4.
    const int quad = warp & 3;
    const int s = (warp - 2) >> 2;
    const int ew = s * 4 + quad;                                      // 0..7
    uint8_t* buf = epi + ew * kEpi3Stage;
    const uint32_t buf_s = smem_u32(buf);
    uint32_t acc = 0, pacc = 0;
    int cur_nb = -1;
    float st[NCB][TF32 ? 2 : 4];
#pragma unroll
    for (int i = 0; i < NCB; ++i)
#pragma unroll
      for (int j = 0; j < (TF32 ? 2 : 4); ++j) st[i][j] = 0.f;
    auto flush = [&](int nb) {
      if (!p.stats_ws || nb < 0) return;
      float* dst = p.stats_ws + (((long long)nb * gridDim.x + blockIdx.x) * kEpi3Warps + ew) * 2 * BLOCK_N;
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
    const uint32_t t_empty_leader0 = mapa_u32(smem_u32(&t_empty[0]), 0);
    constexpr int CPL = TF32 ? 1 : 2;                                   // channels per lane and 128-byte block
    float bmu[BNBWD ? NCB : 1][CPL], bis[BNBWD ? NCB : 1][CPL], bsc[BNBWD ? NCB : 1][CPL], bsh[BNBWD ? NCB : 1][CPL];
    // BNBWD: the warp's 32 x 128 B patch of yprev for (group gt2, channel block cb2) is fetched by TMA into the warp's
    // own buffer as soon as that buffer is free -- i.e. one block ahead, while the warp still waits for the MMAs of
    // the block -- so the loads cost no registers and no exposed latency; tiles / rows outside the tensor are
    // zero-filled by TMA.  Every (group, block) of this CTA is fetched and awaited, valid or not.
    uint8_t* ybuf = yst + ew * kEpi3Stage;
    const uint32_t ybuf_s = smem_u32(ybuf);
    uint32_t yph = 0;
    auto y_fetch = [&](int gt2, int cb2) {                               // lane 0 only
      const int nb2 = gt2 / p.m_groups;
      int mt2 = (gt2 - nb2 * p.m_groups) * CG + (int)rank;
      const bool valid2 = mt2 < p.m_tiles;
      const int tj2 = mt2 % p.tiles_w;
      mt2 /= p.tiles_w;
      const int ti2 = mt2 % p.tiles_h;
      const int b2 = valid2 ? mt2 / p.tiles_h : p.B;
      mbar_expect_tx(&y_full[ew], kEpi3Stage);
      tma_load_4d(ybuf, &p.y_map, &y_full[ew], nb2 * BLOCK_N + cb2 * EPR, tj2 * p.tile_w + s * p.sub1_dj,
                  ti2 * p.tile_h + s * p.sub1_di + 4 * quad, b2);
    };
    if constexpr (BNBWD) {
      if (lane == 0 && group0 < p.total_groups) y_fetch(group0, 0);
    }
    for (int gt = group0; gt < p.total_groups; gt += ngroups) {
      const int nb = gt / p.m_groups;
      int mt = (gt - nb * p.m_groups) * CG + (int)rank;
      const bool valid = mt < p.m_tiles;
      const int tj = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int ti = mt % p.tiles_h;
      const int b = mt / p.tiles_h;
      const int n0 = nb * BLOCK_N;
      if (nb != cur_nb) {
        flush(cur_nb);
        cur_nb = nb;
        if constexpr (BNBWD) {
#pragma unroll
          for (int cb = 0; cb < NCB; ++cb)
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
              const int ch = n0 + cb * EPR + CPL * lane + k;
              bmu[cb][k] = __ldg(p.bnc + ch);
              bis[cb][k] = __ldg(p.bnc + p.N + ch);
              bsc[cb][k] = __ldg(p.bnc + 2 * p.N + ch);
              bsh[cb][k] = __ldg(p.bnc + 3 * p.N + ch);
            }
        }
      }
      const int pi0 = ti * p.tile_h + s * p.sub1_di + 4 * quad;       // first image row of this warp's 4 x 8 patch
      const int pj0 = tj * p.tile_w + s * p.sub1_dj;
      // rows of this warp: r = lane -> pixel (pi0 + r / 8, pj0 + r % 8); bit r of valid_rows: inside the M grid
      const uint32_t valid_rows =
          __ballot_sync(0xffffffffu, valid && (pi0 + (lane >> 3) < p.Hm) && (pj0 + (lane & 7) < p.Wm));
      mbar_wait(&t_full[acc], pacc);
      tc_fence_after();
      if (valid_rows != 0u) {
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {                            // unrolled: st[cb] must stay in registers
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 2 * BLOCK_N + s * BLOCK_N + cb * EPR;
          uint32_t v[EPR];
#pragma unroll
          for (int h = 0; h < EPR / 32; ++h) tmem_ld32_issue(taddr + h * 32, v + h * 32);
          if (lane == 0) tma_store_wait_read_all();                   // the previous store has finished reading `buf`
          tmem_ld_fence();
          __syncwarp();
          if constexpr (AFFINE) {
            // channel of v[k] is n0 + cb * EPR + k for every lane: warp-uniform 16-byte coefficient loads
            const float4* sc4 = reinterpret_cast<const float4*>(p.affine + n0 + cb * EPR);
            const float4* sh4 = reinterpret_cast<const float4*>(p.affine + p.N + n0 + cb * EPR);
#pragma unroll
            for (int k4 = 0; k4 < EPR / 4; ++k4) {
              const float4 a = __ldg(sc4 + k4), c = __ldg(sh4 + k4);
              v[4 * k4 + 0] = __float_as_uint(fmaxf(fmaf(__uint_as_float(v[4 * k4 + 0]), a.x, c.x), 0.f));
              v[4 * k4 + 1] = __float_as_uint(fmaxf(fmaf(__uint_as_float(v[4 * k4 + 1]), a.y, c.y), 0.f));
              v[4 * k4 + 2] = __float_as_uint(fmaxf(fmaf(__uint_as_float(v[4 * k4 + 2]), a.z, c.z), 0.f));
              v[4 * k4 + 3] = __float_as_uint(fmaxf(fmaf(__uint_as_float(v[4 * k4 + 3]), a.w, c.w), 0.f));
            }
          }
          if constexpr (OUTC) {
            const float* ocs = reinterpret_cast<const float*>(epi);
            float lg[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) lg[k] = 0.f;
#pragma unroll
            for (int c = 0; c < 64; ++c) {
              const float z = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[c])));     // the activation as stored
              const float4 w0 = *reinterpret_cast<const float4*>(ocs + c * 8);
              lg[0] = fmaf(z, w0.x, lg[0]); lg[1] = fmaf(z, w0.y, lg[1]);
              lg[2] = fmaf(z, w0.z, lg[2]); lg[3] = fmaf(z, w0.w, lg[3]);
              if (p.ncls > 4) {
                const float4 w1 = *reinterpret_cast<const float4*>(ocs + c * 8 + 4);
                lg[4] = fmaf(z, w1.x, lg[4]); lg[5] = fmaf(z, w1.y, lg[5]);
                lg[6] = fmaf(z, w1.z, lg[6]); lg[7] = fmaf(z, w1.w, lg[7]);
              }
            }
            if ((valid_rows >> lane) & 1u) {
              __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.logits) +
                                   (((long long)b * p.Hm + pi0 + (lane >> 3)) * p.Wm + pj0 + (lane & 7)) * p.ncls;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (k < p.ncls) out[k] = __float2bfloat16_rn(lg[k] + ocs[64 * 8 + k]);
            }
            continue;
          }
          const uint32_t dst = buf_s + lane * 128;
#pragma unroll
          for (int h = 0; h < EPR / 32; ++h) {
            if constexpr (TF32) {
#pragma unroll
              for (int c16 = 0; c16 < 8; ++c16)
                sts_v4(dst + ((c16 ^ (lane & 7)) << 4), v[4 * c16], v[4 * c16 + 1], v[4 * c16 + 2], v[4 * c16 + 3]);
            } else {
#pragma unroll
              for (int c16 = 0; c16 < 4; ++c16) {
                const uint32_t* w = v + h * 32 + 8 * c16;
                sts_v4(dst + (((h * 4 + c16) ^ (lane & 7)) << 4),
                       pack_bf16x2(__uint_as_float(w[0]), __uint_as_float(w[1])),
                       pack_bf16x2(__uint_as_float(w[2]), __uint_as_float(w[3])),
                       pack_bf16x2(__uint_as_float(w[4]), __uint_as_float(w[5])),
                       pack_bf16x2(__uint_as_float(w[6]), __uint_as_float(w[7])));
              }
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&p.o_map, buf, n0 + cb * EPR, pj0, pi0, b);
            tma_store_commit();
          }
          if constexpr (AFFINE && !OUTC && !TF32) {
            if (p.pooled) {
              // lane = 32-bit word (two channels) of the 128-byte row; pooled pixel (py, px) of the patch = max over rows
              // 16 py + 2 px + {0, 1, 8, 9}.  The activations are >= 0 (ReLU), so the unsigned 16-bit SIMD maximum of the
              // bf16 bit patterns is the bf16 maximum.
              const int Hp = p.Hm >> 1, Wp = p.Wm >> 1;
              const int gy0 = pi0 >> 1, gx0 = pj0 >> 1;
#pragma unroll
              for (int py = 0; py < 2; ++py)
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                  const int r0 = 16 * py + 2 * px;
                  uint32_t m = 0u;
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const int r = r0 + (k & 1) + 8 * (k >> 1);
                    m = __vmaxu2(m, lds_u32(buf_s + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2))));
                  }
                  if (valid && gy0 + py < Hp && gx0 + px < Wp)
                    reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.pooled) +
                                                (((long long)b * Hp + gy0 + py) * Wp + gx0 + px) * p.ld_pool + n0 + cb * EPR)[lane] = m;
                }
            }
          }
          if constexpr (BNBWD) {
            // lane = 32-bit word of the 128-byte row (one fp32 / two bf16 channels), over the 32 rows (pixels) of the
            // patch: g from the staged (rounded) tile, yprev from the TMA-fetched patch (same swizzle); rows outside
            // the M grid contribute nothing
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
            uint32_t u[32], yv[32];
            mbar_wait(&y_full[ew], yph);
            yph ^= 1;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const uint32_t off = r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2));
              u[r] = lds_u32(buf_s + off);
              yv[r] = lds_u32(ybuf_s + off);
            }
            __syncwarp();                                                 // every lane has read the yprev patch
            if (lane == 0) {
              if (cb + 1 < NCB) y_fetch(gt, cb + 1);
              else if (gt + ngroups < p.total_groups) y_fetch(gt + ngroups, 0);
            }
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const bool live = (valid_rows >> r) & 1u;
              if constexpr (TF32) {
                const float yy = __uint_as_float(yv[r]);
                const float gg = (live && fmaf(yy, bsc[cb][0], bsh[cb][0]) > 0.f) ? __uint_as_float(u[r]) : 0.f;
                s0 += gg; q0 += gg * ((yy - bmu[cb][0]) * bis[cb][0]);
              } else {
                const float ya = __uint_as_float(yv[r] << 16), yb = __uint_as_float(yv[r] & 0xffff0000u);
                const float ga = (live && fmaf(ya, bsc[cb][0], bsh[cb][0]) > 0.f) ? __uint_as_float(u[r] << 16) : 0.f;
                const float gb = (live && fmaf(yb, bsc[cb][1], bsh[cb][1]) > 0.f) ? __uint_as_float(u[r] & 0xffff0000u) : 0.f;
                s0 += ga; q0 += ga * ((ya - bmu[cb][0]) * bis[cb][0]);
                s1 += gb; q1 += gb * ((yb - bmu[cb][1]) * bis[cb][1]);
              }
            }
            st[cb][0] += s0; st[cb][1] += q0;
            if constexpr (!TF32) { st[cb][2] += s1; st[cb][3] += q1; }
          } else if (p.stats_ws) {
            // lane = 32-bit word of the 128-byte row: sum the rounded values over the 32 rows (branch-free,
            // all loads issued up front; rows outside the M grid are multiplied by 0)
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
            uint32_t u[32];
#pragma unroll
            for (int r = 0; r < 32; ++r)
              u[r] = lds_u32(buf_s + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
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
      }
      if constexpr (BNBWD) {
        if (valid_rows == 0u) {                                         // keep the yprev pipeline in step
#pragma unroll 1
          for (int cb = 0; cb < NCB; ++cb) {
            mbar_wait(&y_full[ew], yph);
            yph ^= 1;
            __syncwarp();
            if (lane == 0) {
              if (cb + 1 < NCB) y_fetch(gt, cb + 1);
              else if (gt + ngroups < p.total_groups) y_fetch(gt + ngroups, 0);
            }
          }
        }
      }
      // this warp has drained its 32 lanes of its sub-tile of accumulator set `acc`
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(t_empty_leader0 + acc * 8);
        else mbar_arrive_local(&t_empty[acc]);
      }
      if (++acc == ACC) { acc = 0; pacc ^= 1; }
    }
    flush(cur_nb);
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();           // no CTA exits (or frees TMEM) while its peer still signals it
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_cg<CG>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Tc3Plan {
  int BN, CG, tile_w, tile_h, box_w, box_h;
  int m_tiles, m_groups, n_blocks, grid;
};

static int tc3_cluster_size() {
  static const int cg = getenv("UNETB200_TC3_CG1") ? 1 : 2;        // 1-CTA variant of the same kernel, for A/B runs
  return cg;
}

static bool tc3_plan(const unetb200_gconv_t* d, Tc3Plan* pl) {
  const int esz = d->dtype == UNETB200_BF16 ? 2 : 4;
  const int epr = 128 / esz;
  if (d->dtype != UNETB200_BF16 && d->dtype != UNETB200_F32) return false;
  if (d->nquad != 1 || d->in_scale != 1 || d->out_scale != 1 || d->ntaps != 9) return false;
  if (d->in_off_y || d->in_off_x) return false;
  if (d->Cin % epr || d->N % 64) return false;
  if ((d->ld_in * esz) % 16 || (d->ld_out * esz) % 16) return false;
  bool seen[9] = {false};
  for (int t = 0; t < 9; ++t) {
    const int dy = d->tap_dy[t], dx = d->tap_dx[t];
    if (dy < -1 || dy > 1 || dx < -1 || dx > 1) return false;
    if (seen[(dy + 1) * 3 + dx + 1]) return false;
    seen[(dy + 1) * 3 + dx + 1] = true;
  }
  pl->CG = tc3_cluster_size();
  // N block: 128 by default -- as a CTA pair the N = 128 MMA already runs at the tensor rate (64 cycles) and two
  // accumulator sets fit in TMEM, so the epilogue overlaps the next tile (N = 256 fills TMEM with one set and
  // measured 8-15% slower on the C >= 256 layers); UNETB200_TC3_MAXBN=256 restores the wide block for A/B runs
  static const int max_bn = getenv("UNETB200_TC3_MAXBN") ? atoi(getenv("UNETB200_TC3_MAXBN")) : 128;
  pl->BN = (d->dtype == UNETB200_BF16 && d->N % 256 == 0 && max_bn >= 256) ? 256 : (d->N % 128 == 0 ? 128 : 64);
  if (d->Wm > 8) { pl->tile_w = 16; pl->tile_h = 16; }
  else { pl->tile_w = 8; pl->tile_h = 32; }
  pl->box_w = pl->tile_w + 2;
  pl->box_h = pl->tile_h + 2;
  const int tiles_w = (d->Wm + pl->tile_w - 1) / pl->tile_w, tiles_h = (d->Hm + pl->tile_h - 1) / pl->tile_h;
  pl->m_tiles = d->B * tiles_w * tiles_h;
  pl->m_groups = (pl->m_tiles + pl->CG - 1) / pl->CG;
  pl->n_blocks = d->N / pl->BN;
  const long long total = (long long)pl->m_groups * pl->n_blocks;
  const int slots = sm_count() / pl->CG;
  pl->grid = (int)(total < slots ? total : slots) * pl->CG;
  return true;
}

int tc3_fprop_supported(const unetb200_gconv_t* d, const void* x, const void* wp, const float* bias, const void* y) {
  static const bool off = getenv("UNETB200_NO_TC3") != nullptr;
  if (off || bias) return 0;      // the 3x3 convolutions of the path are bias-free (unet_parts.py:15,18)
  if (d->dtype == UNETB200_F32 && d->algo != UNETB200_ALGO_TC && d->algo != UNETB200_ALGO_PREFER_TC) return 0;
  Tc3Plan pl;
  if (!tc3_plan(d, &pl)) return 0;
  if (!aligned16(x) || !aligned16(wp) || !aligned16(y)) return 0;
  return 1;
}

long long tc3_stats_workspace(const unetb200_gconv_t* d) {
  Tc3Plan pl;
  if (!tc3_plan(d, &pl)) return 0;
  return (long long)pl.n_blocks * pl.grid * kEpi3Warps * 2 * pl.BN;
}

template <typename T, int BN, int SA, int SB, int ACC, int CG, int TPS, int EPI = EPI_PLAIN>
static int tc3_launch(const Tc3Params& P, int grid, cudaStream_t s) {
  constexpr int smem = SA * kA3Stage + SB * TPS * (BN / CG) * 128 + (EPI == EPI_BNBWD ? 2 : 1) * kEpi3Warps * kEpi3Stage + 1024 + 256;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  if (int rc = set_max_dynamic_smem(reinterpret_cast<const void*>(&tc3_conv_kernel<T, BN, SA, SB, ACC, CG, TPS, EPI>), smem,
                                    "tc3_conv smem attribute"))
    return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kTc3Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc3_conv_kernel<T, BN, SA, SB, ACC, CG, TPS, EPI>, P);
  if (e != cudaSuccess) return cuda_fail(e, "tc3_conv launch");
  return 0;
}

__global__ void __launch_bounds__(1024) tc3_stats_reduce_kernel(const float* __restrict__ ws, int rows, int BN, int N, double* __restrict__ stats) {
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
    stats[which * N + nb * BN + c] += s;
  }
}

int tc3_bnbwd_supported(const unetb200_gconv_t* d) {
  Tc3Plan pl;
  if (!tc3_plan(d, &pl)) return 0;
  return pl.CG == 2 && pl.BN != 256;
}

int tc3_affine_outconv_supported(const unetb200_gconv_t* d, int ncls) {
  Tc3Plan pl;
  if (!tc3_plan(d, &pl)) return 0;
  return d->dtype == UNETB200_BF16 && pl.CG == 2 && pl.BN == 64 && d->N == 64 && ncls >= 1 && ncls <= 8;
}

int tc3_fprop(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* wp, void* y, double* stats,
              float* stats_ws, cudaStream_t stream, const float* affine, const void* yprev, long long ld_yprev,
              const float* bnc, const Tc3OutConv* oc, void* pooled, long long ld_pool) {
  Tc3Plan pl;
  if (pooled && (!affine || oc || d->dtype != UNETB200_BF16 || (ld_pool & 1) || (reinterpret_cast<uintptr_t>(pooled) & 3))) {
    set_error("tc3_fprop: the pooled second output needs the bf16 affine epilogue and 4-byte aligned rows");
    return UNETB200_E_INVALID;
  }
  if (yprev && (affine || !stats || !stats_ws || !bnc || !aligned16(yprev) ||
                (ld_yprev * (d->dtype == UNETB200_BF16 ? 2 : 4)) % 16)) {
    set_error("tc3_fprop: the BatchNorm-backward epilogue needs sums, a workspace, coefficients and 16-byte aligned yprev rows");
    return UNETB200_E_INVALID;
  }
  if (affine && (stats || (reinterpret_cast<uintptr_t>(affine) & 15))) {
    set_error("tc3_fprop: the affine epilogue takes no statistics and needs 16-byte aligned coefficients");
    return UNETB200_E_INVALID;
  }
  if (!tc3_plan(d, &pl)) { set_error("tc3_fprop: unsupported shape"); return UNETB200_E_INVALID; }
  const int esz = d->dtype == UNETB200_BF16 ? 2 : 4;
  Tc3Params P;
  memset(&P, 0, sizeof(P));
  int rc = encode_act_box(&P.a_map, d->dtype, x, d->Cin, d->Win, d->Hin, d->B, d->ld_in, (long long)d->Win * d->ld_in,
                          (long long)d->Hin * d->Win * d->ld_in, pl.box_w, pl.box_h, false);
  if (rc) return rc;
  const char* obase = (const char*)y + ((long long)d->out_off_y * d->Wout + d->out_off_x) * d->ld_out * (long long)esz;
  rc = encode_act_box(&P.o_map, d->dtype, obase, d->N, d->Wm, d->Hm, d->B, d->ld_out, (long long)d->Wout * d->ld_out,
                      (long long)d->Hout * d->Wout * d->ld_out, 8, 4, false);
  if (rc) return rc;
  rc = encode_weights(&P.b_map, d->dtype, wp, g.K, d->N, pl.BN / pl.CG);
  if (rc) return rc;
  for (int t = 0; t < 9; ++t)
    P.tap_aoff[t] = (uint32_t)(((d->tap_dy[t] + 1) * pl.box_w + (d->tap_dx[t] + 1)) * 128) >> 4;
  P.a_box_bytes = (uint32_t)(pl.box_w * pl.box_h * 128);
  P.row_bytes = (uint32_t)pl.box_w * 128;
  if (pl.tile_w == 16) { P.sub1_di = 0; P.sub1_dj = 8; P.sub1_off = (8 * 128) >> 4; }       // two 8-wide sub-tiles side by side
  else { P.sub1_di = 16; P.sub1_dj = 0; P.sub1_off = (uint32_t)(16 * pl.box_w * 128) >> 4; }  // two 16-row sub-tiles stacked
  P.cchunks = d->Cin / (128 / esz);
  P.Cin = d->Cin;
  P.tile_w = pl.tile_w; P.tile_h = pl.tile_h;
  P.tiles_w = (d->Wm + pl.tile_w - 1) / pl.tile_w;
  P.tiles_h = (d->Hm + pl.tile_h - 1) / pl.tile_h;
  P.m_tiles = pl.m_tiles; P.m_groups = pl.m_groups;
  P.total_groups = pl.m_groups * pl.n_blocks;
  P.Hm = d->Hm; P.Wm = d->Wm; P.B = d->B;
  P.stats_ws = stats ? stats_ws : nullptr;
  P.affine = affine;
  P.N = d->N;
  P.bnc = bnc;
  P.pooled = pooled; P.ld_pool = ld_pool;
  if (yprev) {
    rc = encode_act_box(&P.y_map, d->dtype, yprev, d->N, d->Wm, d->Hm, d->B, ld_yprev, (long long)d->Wm * ld_yprev,
                        (long long)d->Hm * d->Wm * ld_yprev, 8, 4, false);
    if (rc) return rc;
  }
  if (oc) {
    if (!affine || !tc3_affine_outconv_supported(d, oc->ncls) || !oc->w || !oc->logits) {
      set_error("tc3_fprop: the OutConv epilogue needs the folded BatchNorm coefficients, bf16, N = 64 and n_classes <= 8");
      return UNETB200_E_INVALID;
    }
    P.oc_w = oc->w; P.oc_b = oc->b; P.logits = oc->logits; P.ncls = oc->ncls;
    return tc3_launch<__nv_bfloat16, 64, 3, 3, 2, 2, 3, EPI_AFFINE_OUT>(P, pl.grid, stream);
  }
  if (affine) {
    // BatchNorm-folded inference epilogue: the pair kernel with N block 128 / 64 (the shapes of the path)
    if (pl.BN == 256) { set_error("tc3_fprop: affine epilogue supports N blocks of 64 / 128"); return UNETB200_E_INVALID; }
    if (d->dtype == UNETB200_BF16) {
      if (pl.CG == 2) return pl.BN == 128 ? tc3_launch<__nv_bfloat16, 128, 2, 3, 2, 2, 3, EPI_AFFINE>(P, pl.grid, stream)
                                          : tc3_launch<__nv_bfloat16, 64, 3, 3, 2, 2, 3, EPI_AFFINE>(P, pl.grid, stream);
      return pl.BN == 128 ? tc3_launch<__nv_bfloat16, 128, 2, 6, 2, 1, 1, EPI_AFFINE>(P, pl.grid, stream)
                          : tc3_launch<__nv_bfloat16, 64, 3, 6, 2, 1, 1, EPI_AFFINE>(P, pl.grid, stream);
    }
    if (pl.CG == 2) return pl.BN == 128 ? tc3_launch<float, 128, 2, 3, 2, 2, 3, EPI_AFFINE>(P, pl.grid, stream)
                                        : tc3_launch<float, 64, 3, 3, 2, 2, 3, EPI_AFFINE>(P, pl.grid, stream);
    return pl.BN == 128 ? tc3_launch<float, 128, 2, 6, 2, 1, 1, EPI_AFFINE>(P, pl.grid, stream)
                        : tc3_launch<float, 64, 3, 6, 2, 1, 1, EPI_AFFINE>(P, pl.grid, stream);
  }
  if (stats) {
    cudaError_t e = cudaMemsetAsync(stats_ws, 0, sizeof(float) * (size_t)tc3_stats_workspace(d), stream);
    if (e != cudaSuccess) return cuda_fail(e, "tc3 stats workspace memset");
  }
  if (yprev) {
    if (pl.CG != 2 || pl.BN == 256) { set_error("tc3_fprop: BatchNorm-backward epilogue: CTA pairs, N blocks of 64 / 128"); return UNETB200_E_INVALID; }
    // N block 64 (the full-resolution layers, one A stage = one whole tile): 3 A stages + 2 weight stages measured
    // 0.439 ms against 0.535 ms for 2 + 3 (64 -> 64 channels, B = 16, 512x512; the plain dgrad takes 0.311 ms and the
    // separate reduction pass 0.180 ms) -- UNETB200_BNBWD64_CFG=0 restores 2 + 3 for A/B runs
    static const int cfg64 = getenv("UNETB200_BNBWD64_CFG") ? atoi(getenv("UNETB200_BNBWD64_CFG")) : 1;
    if (d->dtype == UNETB200_BF16)
      rc = pl.BN == 128 ? tc3_launch<__nv_bfloat16, 128, 2, 3, 2, 2, 3, EPI_BNBWD>(P, pl.grid, stream)
           : cfg64 == 1 ? tc3_launch<__nv_bfloat16, 64, 3, 2, 2, 2, 3, EPI_BNBWD>(P, pl.grid, stream)
                        : tc3_launch<__nv_bfloat16, 64, 2, 3, 2, 2, 3, EPI_BNBWD>(P, pl.grid, stream);
    else
      rc = pl.BN == 128 ? tc3_launch<float, 128, 2, 3, 2, 2, 3, EPI_BNBWD>(P, pl.grid, stream)
                        : tc3_launch<float, 64, 3, 2, 2, 2, 3, EPI_BNBWD>(P, pl.grid, stream);
  } else if (d->dtype == UNETB200_BF16) {
    if (pl.CG == 2) {
      if (pl.BN == 256) rc = tc3_launch<__nv_bfloat16, 256, 2, 2, 1, 2, 3>(P, pl.grid, stream);
      else if (pl.BN == 128) rc = tc3_launch<__nv_bfloat16, 128, 2, 3, 2, 2, 3>(P, pl.grid, stream);
      else rc = tc3_launch<__nv_bfloat16, 64, 3, 3, 2, 2, 3>(P, pl.grid, stream);
    } else {
      if (pl.BN == 256) rc = tc3_launch<__nv_bfloat16, 256, 2, 3, 1, 1, 1>(P, pl.grid, stream);
      else if (pl.BN == 128) rc = tc3_launch<__nv_bfloat16, 128, 2, 6, 2, 1, 1>(P, pl.grid, stream);
      else rc = tc3_launch<__nv_bfloat16, 64, 3, 6, 2, 1, 1>(P, pl.grid, stream);
    }
  } else {
    if (pl.CG == 2) {
      if (pl.BN == 128) rc = tc3_launch<float, 128, 2, 3, 2, 2, 3>(P, pl.grid, stream);
      else rc = tc3_launch<float, 64, 3, 3, 2, 2, 3>(P, pl.grid, stream);
    } else {
      if (pl.BN == 128) rc = tc3_launch<float, 128, 2, 6, 2, 1, 1>(P, pl.grid, stream);
      else rc = tc3_launch<float, 64, 3, 6, 2, 1, 1>(P, pl.grid, stream);
    }
  }
  if (rc) return rc;
  if (stats) {
    const int rows = pl.grid * kEpi3Warps;
    tc3_stats_reduce_kernel<<<dim3((2 * pl.BN + 31) / 32, pl.n_blocks), 1024, 0, stream>>>(stats_ws, rows, pl.BN, d->N, stats);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "tc3_stats_reduce");
  }
  return 0;
}


// ------------------------------------------------------------------------------------------
// wgrad as CTA pairs:  dWp[(t,c)][n] = sum_pixels A[pixel][(t,c)] * G[pixel][n]   (both operands MN-major, K = pixels)
// Same decomposition as tc2_wgrad_kernel (a "unit" = (column shift dx, 128-byte channel chunk); its x box
// {128 B, 8, 8 + 2, 1} serves the three row shifts, one accumulator each), but one tcgen05.mma.cta_group::2
// covers FOUR units (M = 256: two per CTA) against one dY tile of 128 output channels of which each CTA loads
// half -- per 64-pixel step a CTA fills 20 KB + 8 KB instead of 20 KB + 16 KB and the MMA runs at its N = 128
// pair rate (64 cycles instead of 87).  Used when the unit count is a multiple of 4 (C_in >= 256).
// ------------------------------------------------------------------------------------------
struct alignas(64) Tc3WParams {
  CUtensorMap a_map;         // x view, box {128 B, 8, 10, 1}
  CUtensorMap o_map;         // dY view, box {128 B, 8, 8, 1}
  int g_dx[3], g_dy0, g_tap[3][3];
  int Cin, nunits;
  int tiles_w, tiles_h;
  int ptiles, ptiles_per_split;
  int N, K;
  float* partials;
};

template <int STAGES>
__global__ void __launch_bounds__(192, 1) tc3_wgrad_kernel(const __grid_constant__ Tc3WParams p) {
  constexpr int EPR = 64, BLOCK_N = 128;
  constexpr uint32_t kGSub = 64 * 128;                 // one 8x8-pixel x 128-byte dY sub-tile (this CTA's 64 channels)
  constexpr uint32_t kABox = 10 * 8 * 128;             // x box
  constexpr uint32_t kStage = 2 * kABox + kGSub;       // 28 KB
  constexpr int UMMA_K = 16, MMAS = 64 / UMMA_K;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStage);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int mt = blockIdx.x >> 1;                      // pair index along M: units 4*mt .. 4*mt + 3
  const int n0 = blockIdx.y * BLOCK_N;
  const int split = blockIdx.z;
  const int pt_begin = split * p.ptiles_per_split;
  int pt_end = pt_begin + p.ptiles_per_split;
  if (pt_end > p.ptiles) pt_end = p.ptiles;
  const int num_k = pt_end - pt_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map);
    tma_prefetch_desc(&p.o_map);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_cg<2>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int pt = pt_begin;
    int tj = pt % p.tiles_w;
    int rest = pt / p.tiles_w;
    int ti = rest % p.tiles_h;
    int b = rest / p.tiles_h;
    int uc[2], ug[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int u = mt * 4 + (int)rank * 2 + h;         // < nunits (nunits % 4 == 0)
      uc[h] = u / 3;
      ug[h] = u - uc[h] * 3;
    }
    uint32_t s = 0, ph = 1;
    for (int kb = 0; kb < num_k; ++kb) {
      mbar_wait(&empty_bar[s], ph);
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * kStage);
        const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
        const int i0 = ti * 8, j0 = tj * 8;
        uint8_t* sa = smem + s * kStage;
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tma_load_4d_cg<2>(sa + h * kABox, &p.a_map, bar, uc[h] * EPR, j0 + p.g_dx[ug[h]], i0 + p.g_dy0, b);
        tma_load_4d_cg<2>(sa + 2 * kABox, &p.o_map, bar, n0 + (int)rank * 64, j0, i0, b);
      }
      if (++s == STAGES) { s = 0; ph ^= 1; }
      if (++tj == p.tiles_w) { tj = 0; if (++ti == p.tiles_h) { ti = 0; ++b; } }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(false, true, true, 256, BLOCK_N);
      // MN-major: LBO = distance between 128-byte-wide sub-tiles, SBO = one swizzle group of pixel rows
      const uint64_t da_t = make_desc(smem_u32(smem), kABox, 1024, kLayoutSW128);
      const uint64_t db_t = make_desc(smem_u32(smem) + 2 * kABox, kGSub, 1024, kLayoutSW128);
      uint32_t s = 0, ph = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t da0 = da_t + s * (kStage >> 4), db0 = db_t + s * (kStage >> 4);
        if (elect_one()) {
#pragma unroll
          for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int k = 0; k < MMAS; ++k)
              umma_cg<2, false>(tmem_base + a * BLOCK_N, da0 + ((a * 1024 + k * UMMA_K * 128) >> 4),
                                db0 + ((k * UMMA_K * 128) >> 4), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_cg<2>(&empty_bar[s]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit_cg<2>(tmem_full);
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                  // D row of this CTA = (unit row / 64, channel row % 64)
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int u = mt * 4 + (int)rank * 2 + row / EPR;
    const int c = u / 3, g = u - c * 3;
    for (int a = 0; a < 3; ++a) {
      const long long k = (long long)p.g_tap[g][a] * p.Cin + c * EPR + (row % EPR);
      float* out = p.partials + (long long)split * p.K * p.N + k * p.N + n0;
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + a * BLOCK_N + ch * 32, v);
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
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_cg<2>(tmem_base, 512);
  }
}

struct Tc3WPlan {
  int nunits, mpairs, ntiles, ptiles, tiles_w, tiles_h;
  int g_dx[3], g_dy0, g_tap[3][3];
};

static bool tc3_wgrad_plan(const unetb200_gconv_t* d, Tc3WPlan* w) {
  static const bool off = getenv("UNETB200_NO_TC3") != nullptr || getenv("UNETB200_NO_TC3W") != nullptr;
  if (off) return false;
  if (d->dtype != UNETB200_BF16) return false;
  if (d->nquad != 1 || d->in_scale != 1 || d->out_scale != 1 || d->ntaps != 9) return false;
  if (d->in_off_y || d->in_off_x) return false;
  if (d->Cin % 64 || d->N % 128) return false;
  if ((d->ld_in * 2) % 16 || (d->ld_out * 2) % 16) return false;
  const int cchunks = d->Cin / 64;
  if ((3 * cchunks) % 4) return false;                   // a pair owns four units
  // groups: column shift dx = -1, 0, 1, each with the row shifts dy = -1, 0, 1
  for (int g = 0; g < 3; ++g) {
    w->g_dx[g] = g - 1;
    for (int a = 0; a < 3; ++a) {
      int found = -1;
      for (int t = 0; t < 9; ++t)
        if (d->tap_dx[t] == g - 1 && d->tap_dy[t] == a - 1) found = t;
      if (found < 0) return false;
      w->g_tap[g][a] = found;
    }
  }
  w->g_dy0 = -1;
  w->nunits = 3 * cchunks;
  w->mpairs = w->nunits / 4;
  w->ntiles = d->N / 128;
  w->tiles_w = (d->Wm + 7) / 8;
  w->tiles_h = (d->Hm + 7) / 8;
  w->ptiles = d->B * w->tiles_w * w->tiles_h;
  return true;
}

int tc3_wgrad_supported(const unetb200_gconv_t* d, const void* x, const void* gy) {
  Tc3WPlan w;
  if (!tc3_wgrad_plan(d, &w)) return 0;
  if ((x && !aligned16(x)) || (gy && !aligned16(gy))) return 0;
  return 1;
}

int tc3_wgrad_splits(const unetb200_gconv_t* d) {
  Tc3WPlan w;
  if (!tc3_wgrad_plan(d, &w)) return 1;
  const long long pairs = (long long)w.mpairs * w.ntiles;
  const long long max_by_k = (w.ptiles + 15) / 16;      // at least 16 pixel tiles (1024 pixels) per split
  return pick_splits(pairs, sm_count() / 2, max_by_k);  // one CTA pair per SM pair: whole waves
}

int tc3_wgrad(const unetb200_gconv_t* d, const GconvDev& g, const void* x, const void* gy, float* partials, int splits,
              cudaStream_t stream) {
  Tc3WPlan w;
  if (!tc3_wgrad_plan(d, &w)) { set_error("tc3_wgrad: unsupported shape"); return UNETB200_E_INVALID; }
  Tc3WParams P;
  memset(&P, 0, sizeof(P));
  int rc = encode_act_box(&P.a_map, d->dtype, x, d->Cin, d->Win, d->Hin, d->B, d->ld_in, (long long)d->Win * d->ld_in,
                          (long long)d->Hin * d->Win * d->ld_in, 8, 10, true);
  if (rc) return rc;
  const char* obase = (const char*)gy + ((long long)d->out_off_y * d->Wout + d->out_off_x) * d->ld_out * 2LL;
  rc = encode_act_box(&P.o_map, d->dtype, obase, d->N, d->Wm, d->Hm, d->B, d->ld_out, (long long)d->Wout * d->ld_out,
                      (long long)d->Hout * d->Wout * d->ld_out, 8, 8, true);
  if (rc) return rc;
  for (int gi = 0; gi < 3; ++gi) {
    P.g_dx[gi] = w.g_dx[gi];
    for (int a = 0; a < 3; ++a) P.g_tap[gi][a] = w.g_tap[gi][a];
  }
  P.g_dy0 = w.g_dy0;
  P.Cin = d->Cin; P.nunits = w.nunits;
  P.tiles_w = w.tiles_w; P.tiles_h = w.tiles_h;
  P.ptiles = w.ptiles;
  P.ptiles_per_split = (w.ptiles + splits - 1) / splits;
  P.N = d->N; P.K = g.K;
  P.partials = partials;
  // pipeline depth: 7 stages fill the SM's shared memory; UNETB200_TC3W_STAGES=5 leaves ~85 KB for co-resident
  // memory-bound kernels of the main stream (A/B switch)
  static const int st_env = getenv("UNETB200_TC3W_STAGES") ? atoi(getenv("UNETB200_TC3W_STAGES")) : 7;
  const bool deep = st_env >= 7;
  const int smem = (deep ? 7 : 5) * (2 * 10240 + 8192) + 1024 + 256;
  if (int rc = deep ? set_max_dynamic_smem(reinterpret_cast<const void*>(&tc3_wgrad_kernel<7>), smem, "tc3_wgrad smem attribute")
                    : set_max_dynamic_smem(reinterpret_cast<const void*>(&tc3_wgrad_kernel<5>), smem, "tc3_wgrad smem attribute"))
    return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)w.mpairs * 2, (unsigned)w.ntiles, (unsigned)splits);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = deep ? cudaLaunchKernelEx(&cfg, tc3_wgrad_kernel<7>, P) : cudaLaunchKernelEx(&cfg, tc3_wgrad_kernel<5>, P);
  if (e != cudaSuccess) return cuda_fail(e, "tc3_wgrad launch");
  return 0;
}

}  // namespace ub
