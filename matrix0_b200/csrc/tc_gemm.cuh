// tcgen05 / TMEM / TMA implicit-GEMM kernel for sm_100a (hand-written PTX wrappers + kernel).
//
//   D[m][n] = sum_k A(m, k) * W[n][k]          A, W bf16 (K-major), D fp32 accumulated in tensor memory
//
// A is either a plain row-major matrix [M][K] (2-D TMA map) or the 3x3 "same" convolution view of an
// NHWC activation tensor [B][8][8][C]: M = B*64 rows, K = 9*C, one (ky,kx) tap per group of C/64
// k-blocks.  For the convolution the A tile of a k-block is ONE 4-D TMA box {64 ch, 8, 8, 2 boards}
// whose (x, y) start coordinates are shifted by the tap offset: the TMA unit zero-fills the
// out-of-board elements, which is exactly the zero padding of the convolution, and writes the
// 128 rows x 128 B in the 128-byte-swizzled K-major layout tcgen05.mma consumes (no im2col buffer).
//
// CTA = 6 warps: warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one elected lane,
// owns the TMEM allocation), warps 2..5 = epilogue (one thread per accumulator row / TMEM lane).
// Persistent over M tiles of 128 rows; smem ring of STAGES x {A 16 KB, W N*128 B}; accumulator
// 128 lanes x N fp32 columns of tensor memory (N <= 320 -> 512-column allocation).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "nn.cuh"

namespace m0 {
namespace tc {

static constexpr int BM = 128;          // rows per tile (2 boards)
static constexpr int BK = 64;           // bf16 elements per k-block = one 128-byte swizzle row
static constexpr int A_TILE_BYTES = BM * BK * 2;
static constexpr int NUM_THREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per TMEM lane quarter)

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// the same box delivered to the same shared-memory offset of every CTA in cta_mask; each destination CTA's mbarrier
// (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mcast(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16 (bf16 inputs, fp32 accumulate), issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// make the mbarrier track completion of all tcgen05 operations issued so far by this thread
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the same, arriving on the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (base_lane + t), columns c..c+31
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (= 1, unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B -> 64)   [46,48) version = 1   [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::f16: c_format F32 (bit 4), a/b format BF16 (bits 7, 10), K-major A and B,
// n_dim = N >> 3 at [17,23), m_dim = M >> 4 at [24,29)
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n, int fp16 = 0) {
  const uint32_t fmt = fp16 ? 0u : 1u;  // F16F32Format: 0 = F16, 1 = BF16
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- kernel ----------------------------------------------------------------------------------------------------------
struct GemmParams {
  int M;            // valid rows of D (rows >= M are computed on zero / stale A rows and not stored)
  int N;            // output columns handled by this launch (<= 320, multiple of 16)
  int n_store;      // columns actually stored, counted over all n_slices (<= N * n_slices; the rest are zero-padded weight rows)
  int n_part;       // columns per tcgen05.mma (N if N <= 256, else N / 2); multiple of 16
  int taps;         // 9 (3x3 convolution) or 1 (plain GEMM / 1x1 convolution)
  int kb_per_tap;   // k-blocks per tap = Cin / 64 (plain GEMM: K / 64)
  int conv;         // 1: A is the 4-D NHWC map, 0: A is the 2-D [M][K] map
  int w_row0;       // first row of W (output-channel offset of this launch)
  int stages;
  int tmem_cols;    // power of two >= N, >= 32
  int fp16;         // operands are IEEE fp16 instead of bf16
  int cluster;      // CTAs per cluster (1, 2 or 4): the W tile of a stage is loaded once per cluster and multicast
  // epilogue: out[m][col0 + n] = act(acc + bias[n]) * scale; either output may be null
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int ldc, col0;
  const float* bias;   // indexed by w_row0 + n when non-null
  int act;
  float scale;
  // fused epilogues (need the full channel range in the tile: w_row0 == 0, N == channels, N % 32 == 0):
  //   gn_gamma != null : out_bf16 = half(act(GroupNorm_16ch(acc) * gamma + beta)); statistics over the 64 rows of each board
  //                      (nn.GroupNorm(C/16, C) + activation that follow conv1 in the pre-activation block, resnet.py:49-50)
  //   pool_part != null: pool_part[(board*2 + half)][n] = sum over the 32 rows of that half board (SE squeeze, resnet.py:61)
  const float* gn_gamma;
  const float* gn_beta;
  float* pool_part;
  int k_splits;      // >= 1: the K range is cut into k_splits equal parts (whole k-blocks); part ks writes its partial sums to
                     // out_f32 + ks * split_stride (plain fp32 output only) -- more CTAs for tall-K, few-tile GEMMs
  long long split_stride;
  int n_slices;      // >= 1: the launch covers n_slices consecutive groups of N output columns (w_row0 / col0 advance by N per slice);
                     // a CTA walks the slices of one M tile back to back, so the A tile is re-read from L2, not from HBM
  int acc_stages;    // 1 or 2 accumulators in tensor memory (2 when two of them fit in 512 columns): the MMAs of work item i + 1 then
                     // overlap the epilogue of work item i -- the K <= 320 GEMMs are bound by their epilogues otherwise
  int acc_stride;    // TMEM column distance between the accumulators
};

// two floats -> one 32-bit word of bf16 or fp16 (low half = first value)
__device__ __forceinline__ uint32_t pack_half2(float a, float b, int fp16) {
  if (fp16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ float tc_act(float x, int act) {
  switch (act) {
    case ACT_RELU: return x > 0.0f ? x : 0.0f;
    case ACT_SILU: return __fdividef(x, 1.0f + __expf(-x));   // 16-bit outputs: MUFU reciprocal instead of the IEEE division
    case ACT_LEAKY: return x > 0.0f ? x : 0.05f * x;
    case ACT_TANH: return tanhf(x);
    case ACT_SIGMOID: return __fdividef(1.0f, 1.0f + __expf(-x));
    default: return x;
  }
}

static constexpr int EPI_STAGE_BYTES = 8 * 32 * 32 * 4 + 2 * 320 * 4 + 4 * 20 * 2 * 4;  // 8 transpose buffers + gamma/beta + GN partial sums

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the 128-byte swizzle needs 1024-byte aligned tiles; the offset is identical in every CTA of a cluster
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int w_tile_bytes = p.N * BK * 2;
  const int stage_bytes = A_TILE_BYTES + w_tile_bytes;
  float* epi_stage = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes);
  float* s_gamma = epi_stage + 8 * 32 * 32;   // [320]
  float* s_beta = s_gamma + 320;              // [320]
  float* s_stats = s_beta + 320;              // [4 quarters][20 groups][sum, sumsq]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes + EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  const int na = p.acc_stages > 1 ? 2 : 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cs = p.cluster;
  const int rank = cs > 1 ? (int)cluster_ctarank() : 0;
  const int num_tiles = (p.M + BM - 1) / BM;
  const int num_groups = (num_tiles + cs - 1) / cs;          // a cluster processes cs consecutive M tiles together
  const int num_clusters = gridDim.x / cs;
  const int cluster_id = blockIdx.x / cs;
  const int kblocks = p.taps * p.kb_per_tap;
  const int n_parts = p.N / p.n_part;
  const int slice_rows = p.n_part / cs;                       // W rows of each part fetched by this CTA for the whole cluster
  const int ns = p.n_slices > 1 ? p.n_slices : 1;
  const int nk = p.k_splits > 1 ? p.k_splits : 1;          // work item = (group of M tiles, column slice, K part)
  const int kb_part = kblocks / nk;
  const uint16_t mask = (uint16_t)((1u << cs) - 1u);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_w);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], (uint32_t)cs);   // every CTA of the cluster must have consumed the stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // peers' barriers are initialised before anything can arrive on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int gi = cluster_id; gi < num_groups * ns * nk; gi += num_clusters) {
        const int ks = gi % nk, gs = gi / nk;
        const int g = gs / ns, w_row0 = p.w_row0 + (gs - g * ns) * p.N;
        const int m0 = (g * cs + rank) * BM;   // may lie past M for the padding tiles of the last group: TMA zero-fills
        for (int kb = ks * kb_part; kb < (ks + 1) * kb_part; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + (size_t)stage * stage_bytes;
          uint8_t* w_dst = a_dst + A_TILE_BYTES;
          mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          const int tap = kb / p.kb_per_tap, kc = kb - tap * p.kb_per_tap;
          if (p.conv) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            tma_load_4d(a_dst, &tma_a, &full_bar[stage], kc * BK, dx, dy, m0 >> 6);
          } else {
            tma_load_2d(a_dst, &tma_a, &full_bar[stage], kb * BK, m0);
          }
          for (int part = 0; part < n_parts; ++part) {
            uint8_t* dst = w_dst + ((size_t)part * p.n_part + (size_t)rank * slice_rows) * BK * 2;
            const int row = w_row0 + part * p.n_part + rank * slice_rows;
            if (cs > 1) tma_load_2d_mcast(dst, &tma_w, &full_bar[stage], kb * BK, row, mask);
            else tma_load_2d(dst, &tma_w, &full_bar[stage], kb * BK, row);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // The whole warp walks the pipeline (all values stay warp-uniform, so descriptors live in uniform registers);
    // one elected lane issues the tcgen05 instructions.
    const uint32_t idesc = make_idesc_bf16(BM, p.n_part, p.fp16);
    const uint32_t part_bytes = (uint32_t)(p.n_part * BK * 2);
    int stage = 0;
    uint32_t phase = 0, it = 0;
    const uint32_t smem_base = smem_u32(smem);
    for (int gi = cluster_id; gi < num_groups * ns * nk; gi += num_clusters, ++it) {
      const uint32_t acc = na == 2 ? (it & 1u) : 0u, acc_phase = na == 2 ? ((it >> 1) & 1u) : (it & 1u);
      const uint32_t tmem_acc = tmem_base + acc * (uint32_t)p.acc_stride;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      for (int kb = 0; kb < kb_part; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)stage * (uint32_t)stage_bytes;
        const uint64_t adesc0 = make_smem_desc(a_addr);
        const uint64_t bdesc0 = make_smem_desc(a_addr + A_TILE_BYTES);
        const uint64_t bdesc1 = make_smem_desc(a_addr + A_TILE_BYTES + part_bytes);
        if (elect_one()) {
          if (n_parts == 2) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
              umma_bf16(tmem_acc, adesc0 + 2 * k, bdesc0 + 2 * k, idesc, accum);
              umma_bf16(tmem_acc + (uint32_t)p.n_part, adesc0 + 2 * k, bdesc1 + 2 * k, idesc, accum);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_acc, adesc0 + 2 * k, bdesc0 + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          // frees the smem stage (in every CTA of the cluster) once these MMAs have read it
          if (cs > 1) umma_commit_mcast(&empty_bar[stage], mask);
          else umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (bias / activation) -> smem transpose -> coalesced global stores =====
    const int quarter = warp & 3;            // tcgen05.ld: warp w may touch lanes 32*(w%4) .. +31
    const int cset = (warp - 2) >> 2;        // the two warps of a quarter take alternate 32-column chunks
    float* stg = epi_stage + (warp - 2) * (32 * 32);
    const bool plain = (p.bias == nullptr) && (p.act == ACT_NONE) && (p.scale == 1.0f);
    const bool fused_gn = p.gn_gamma != nullptr;
    const int epi_tid = ((warp - 2) << 5) | lane;   // 0..255
    if (fused_gn) {
      for (int c = epi_tid; c < p.N; c += 256) { s_gamma[c] = p.gn_gamma[c]; s_beta[c] = p.gn_beta[c]; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    uint32_t it = 0;
    for (int gi = cluster_id; gi < num_groups * ns * nk; gi += num_clusters, ++it) {
      const uint32_t acc = na == 2 ? (it & 1u) : 0u, acc_phase = na == 2 ? ((it >> 1) & 1u) : (it & 1u);
      const int ks = gi % nk, gs = gi / nk;
      const int g = gs / ns, slice = gs - g * ns;
      float* const out_f32 = p.out_f32 ? p.out_f32 + (size_t)ks * p.split_stride : nullptr;
      const int w_row0 = p.w_row0 + slice * p.N, col0 = p.col0 + slice * p.N;
      const int n_store = (p.n_store - slice * p.N) < p.N ? (p.n_store - slice * p.N) : p.N;   // p.n_store counts over all slices
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int tile_row0 = (g * cs + rank) * BM + quarter * 32;
      const uint32_t tmem_row = tmem_base + acc * (uint32_t)p.acc_stride + ((uint32_t)(quarter * 32) << 16);
      if (fused_gn) {
        // pass 1: per (half board = this warp, group of 16 channels) sum and sum of squares
        for (int c0 = cset * 32; c0 < p.N; c0 += 64) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_row + (uint32_t)c0, r);
          tmem_ld_wait();
          float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(r[j]), b = __uint_as_float(r[16 + j]);
            s0 += a; q0 = fmaf(a, a, q0);
            s1 += b; q1 = fmaf(b, b, q1);
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            s0 += __shfl_xor_sync(0xFFFFFFFFu, s0, off); q0 += __shfl_xor_sync(0xFFFFFFFFu, q0, off);
            s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, off); q1 += __shfl_xor_sync(0xFFFFFFFFu, q1, off);
          }
          if (lane == 0) {
            float* st = s_stats + (quarter * 20 + (c0 >> 4)) * 2;
            st[0] = s0; st[1] = q0; st[2] = s1; st[3] = q1;
          }
        }
        // the two warps that hold one board (quarters 2b, 2b+1) exchange their partial sums
        if (quarter < 2) asm volatile("bar.sync 2, 128;" ::: "memory");
        else asm volatile("bar.sync 3, 128;" ::: "memory");
      }
      for (int c0 = cset * 32; c0 < p.N; c0 += 64) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_row + (uint32_t)c0, r);
        tmem_ld_wait();
        if (fused_gn) {
          const int act = p.act;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int grp = (c0 >> 4) + h;
            const float* sa = s_stats + ((quarter & 2) * 20 + grp) * 2;
            const float* sb = s_stats + ((quarter | 1) * 20 + grp) * 2;
            const float mean = (sa[0] + sb[0]) * (1.0f / 1024.0f);
            const float var = fmaxf((sa[1] + sb[1]) * (1.0f / 1024.0f) - mean * mean, 0.0f);
            const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = c0 + h * 16 + j;
              const float gm = s_gamma[c] * rstd;
              const float y = fmaf(__uint_as_float(r[h * 16 + j]) - mean, gm, s_beta[c]);
              r[h * 16 + j] = __float_as_uint(tc_act(y, act));
            }
          }
        } else if (!plain) {
          // bias (one lane per column, broadcast by shuffle), activation chosen once per chunk, scale
          float bl = 0.0f;
          if (p.bias && c0 + lane < n_store) bl = __ldg(p.bias + w_row0 + c0 + lane);
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[j]) + __shfl_sync(0xFFFFFFFFu, bl, j);
          const int act = p.act;
          if (act == ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __fdividef(1.0f, 1.0f + __expf(-x[j]));
          } else if (act == ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __fdividef(x[j], 1.0f + __expf(-x[j]));
          } else if (act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.0f);
          } else if (act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = tc_act(x[j], act);
          }
          const float scale = p.scale;
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(x[j] * scale);
        }
        // lane = row of the 32x32 block; 16-byte chunk q of row i is stored at chunk position q ^ (i & 7)
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        __syncwarp();
        const int ncols = (n_store - c0) < 32 ? (n_store - c0) : 32;
        if (p.pool_part && tile_row0 < p.M && lane < ncols) {
          // column sums of this half board (lane = column): conflict-free reads of the swizzled buffer
          float cs_sum = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) cs_sum += stg[i * 32 + ((((lane >> 2) ^ (i & 7)) << 2) | (lane & 3))];
          p.pool_part[(size_t)(tile_row0 >> 5) * p.N + c0 + lane] = cs_sum;
        }
        if (out_f32) {
          // 8 lanes cover the 128 contiguous bytes of one row: each store instruction writes 4 full lines
          const int q = lane & 7;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = (lane >> 3) + 4 * i;
            const int m = tile_row0 + rr;
            if (m < p.M && 4 * q < ncols) {
              uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 32 + ((q ^ (rr & 7)) << 2));
              *reinterpret_cast<uint4*>(out_f32 + (size_t)m * p.ldc + col0 + c0 + 4 * q) = v;
            }
          }
        }
        if (p.out_bf16) {
          // 4 lanes cover the 64 contiguous bytes of one row (8 half-precision values per lane)
          const int q2 = lane & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = (lane >> 2) + 8 * i;
            const int m = tile_row0 + rr;
            if (m < p.M && 8 * q2 < ncols) {
              float4 lo = *reinterpret_cast<const float4*>(stg + rr * 32 + (((2 * q2) ^ (rr & 7)) << 2));
              float4 hi = *reinterpret_cast<const float4*>(stg + rr * 32 + (((2 * q2 + 1) ^ (rr & 7)) << 2));
              uint4 pk;
              pk.x = pack_half2(lo.x, lo.y, p.fp16); pk.y = pack_half2(lo.z, lo.w, p.fp16);
              pk.z = pack_half2(hi.x, hi.y, p.fp16); pk.w = pack_half2(hi.z, hi.w, p.fp16);
              *reinterpret_cast<uint4*>(p.out_bf16 + (size_t)m * p.ldc + col0 + c0 + 8 * q2) = pk;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // no CTA leaves while peers may still multicast into it or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace tc
}  // namespace m0
