// ChessAttention core (azchess/model/resnet.py:141-174) on tensor cores: one warp per (board, head).
// S = Q K^T / sqrt(d) + rel_bias, clamp +-50, masked softmax (mask == 0 -> -1e4) and unmasked softmax,
// O = (1 - mix) * P_masked V + mix * P_unmasked V.  64 tokens x head_dim 16: the whole head lives in registers as
// mma.sync.m16n8k16 fragments (fp16 or bf16 operands, fp32 accumulate).  Inputs / outputs are the half-precision
// buffers of the tensor-core GEMMs.
//
// The kernel is instruction-bound (2 x 4096 probabilities per head), so the per-element work is kept minimal:
//   * the 64x64 bias of the head is staged once per block, pre-multiplied by log2(e), and a block walks 32 boards;
//   * the attack-pattern mask (resnet.py:105-129) is a 0/1 table staged next to the bias;
//   * ONE exponential per element: softmax is shift invariant, so the masked probabilities are 2^(s - max_all)
//     times the mask (masked-out entries are exactly 0 in the reference too: exp(-1e4 - max) underflows), and both
//     normalisers come out of the tensor cores as P * ones.  Rows whose allowed squares hold < 2^-6 of the mass
//     are rescaled by their own maximum before the 16-bit rounding (warp-uniform branch, rare);
//   * V fragments come from 32-bit loads + movmatrix.trans instead of 16-bit gathers.
// Blocks are ordered head-fastest so that the 20 heads of a board (one 1920-byte row of qkv per token) are read
// by co-scheduled blocks while the lines are still in L2.
//
// Two kernels share the arithmetic (attend_head): attention_tc_staged_kernel (the default) brings the next board's q / k / v
// slice into a per-warp shared-memory buffer with cp.async while the current board is computed; attention_tc_kernel loads the
// fragments straight from global memory at the top of every board (M0_ATT_STAGED=0, kept for A/B).  Their outputs are
// bit-identical (tests/test_nn_gpu.py::test_attention_staged_kernel_is_bit_identical_to_direct_loads).
#include "nn.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace m0 {

template <bool FP16>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  if (FP16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
}
template <bool FP16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if (FP16) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// lane (g, t) holds element pair [g][2t, 2t+1] of an 8x8 b16 tile; returns the pair [g][2t, 2t+1] of its transpose
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ bool attn_mask_tc(int i, int j) {  // resnet.py:105-129
  int dr = (i >> 3) - (j >> 3), dc = (i & 7) - (j & 7);
  int adr = dr < 0 ? -dr : dr, adc = dc < 0 ? -dc : dc;
  return dr == 0 || dc == 0 || adr == adc || (adr == 2 && adc == 1) || (adr == 1 && adc == 2) || (adr <= 1 && adc <= 1);
}

static constexpr int BIAS_LD = 72;          // padded row stride of the staged tables (floats)
static constexpr int ATT_BOARDS_PER_WARP = 4;    // direct-load kernel; the staged kernel takes 1..8 as an argument
static constexpr float LOG2E = 1.4426950408889634f;

// One (board, head): the four 16-row tiles of S = Q K^T, the two softmaxes and O = P V from register-resident fragments
// (kf: K^T as the col-major B operand, vf: V as the B operand of P V, qall: the A fragments of the four row tiles).
template <bool FP16>
__device__ __forceinline__ void attend_head(const uint32_t (&kf)[8][2], const uint32_t (&vf)[4][2][2], const uint32_t (&qall)[4][4],
                                            const float* s_bias, const float* s_mask, uint16_t* __restrict__ out,
                                            int b, int h, int C, int g, int t, float mix) {
  const float blend = 1.0f - mix;
  const float kscale = 0.25f * LOG2E, kclamp = 50.0f * LOG2E;
  // B fragment of an all-ones 16 x 8 tile: P * ones = the row sums of the (rounded) probabilities the PV product uses
  const uint32_t one2 = pack2<FP16>(1.0f, 1.0f);
  const uint32_t ones[2] = {one2, one2};
#pragma unroll 1
  for (int mt = 0; mt < 4; ++mt) {
    const int r0 = 16 * mt + g, r1 = r0 + 8;
    uint32_t qa[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) qa[e] = mt == 0 ? qall[0][e] : mt == 1 ? qall[1][e] : mt == 2 ? qall[2][e] : qall[3][e];
    float su[8][4];
    float mu0 = -INFINITY, mu1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      mma16816<FP16>(acc, qa, kf[nt]);
      const float2 b0 = *reinterpret_cast<const float2*>(s_bias + r0 * BIAS_LD + 8 * nt + 2 * t);
      const float2 b1 = *reinterpret_cast<const float2*>(s_bias + r1 * BIAS_LD + 8 * nt + 2 * t);
      su[nt][0] = fminf(fmaxf(fmaf(acc[0], kscale, b0.x), -kclamp), kclamp);
      su[nt][1] = fminf(fmaxf(fmaf(acc[1], kscale, b0.y), -kclamp), kclamp);
      su[nt][2] = fminf(fmaxf(fmaf(acc[2], kscale, b1.x), -kclamp), kclamp);
      su[nt][3] = fminf(fmaxf(fmaf(acc[3], kscale, b1.y), -kclamp), kclamp);
      mu0 = fmaxf(mu0, fmaxf(su[nt][0], su[nt][1]));
      mu1 = fmaxf(mu1, fmaxf(su[nt][2], su[nt][3]));
    }
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {
      mu0 = fmaxf(mu0, __shfl_xor_sync(0xFFFFFFFFu, mu0, off));
      mu1 = fmaxf(mu1, __shfl_xor_sync(0xFFFFFFFFu, mu1, off));
    }
    // softmax is shift invariant: both distributions use 2^(s - row max); the masked one multiplies by the 0/1 mask
    // (the reference's masked entries are exactly 0 as well: exp(-1e4 - max) underflows)
    float sm[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 m0 = *reinterpret_cast<const float2*>(s_mask + r0 * BIAS_LD + 8 * nt + 2 * t);
      const float2 m1 = *reinterpret_cast<const float2*>(s_mask + r1 * BIAS_LD + 8 * nt + 2 * t);
      su[nt][0] = fast_ex2(su[nt][0] - mu0); su[nt][1] = fast_ex2(su[nt][1] - mu0);
      su[nt][2] = fast_ex2(su[nt][2] - mu1); su[nt][3] = fast_ex2(su[nt][3] - mu1);
      sm[nt][0] = su[nt][0] * m0.x; sm[nt][1] = su[nt][1] * m0.y;
      sm[nt][2] = su[nt][2] * m1.x; sm[nt][3] = su[nt][3] * m1.y;
    }
    float ou[2][4], om[2][4], zu[4] = {0.f, 0.f, 0.f, 0.f}, zm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int nd = 0; nd < 2; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) ou[nd][e] = om[nd][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pu[4], pm[4];
      pu[0] = pack2<FP16>(su[2 * kk][0], su[2 * kk][1]);
      pu[1] = pack2<FP16>(su[2 * kk][2], su[2 * kk][3]);
      pu[2] = pack2<FP16>(su[2 * kk + 1][0], su[2 * kk + 1][1]);
      pu[3] = pack2<FP16>(su[2 * kk + 1][2], su[2 * kk + 1][3]);
      pm[0] = pack2<FP16>(sm[2 * kk][0], sm[2 * kk][1]);
      pm[1] = pack2<FP16>(sm[2 * kk][2], sm[2 * kk][3]);
      pm[2] = pack2<FP16>(sm[2 * kk + 1][0], sm[2 * kk + 1][1]);
      pm[3] = pack2<FP16>(sm[2 * kk + 1][2], sm[2 * kk + 1][3]);
#pragma unroll
      for (int nd = 0; nd < 2; ++nd) {
        mma16816<FP16>(ou[nd], pu, vf[kk][nd]);
        mma16816<FP16>(om[nd], pm, vf[kk][nd]);
      }
      mma16816<FP16>(zu, pu, ones);
      mma16816<FP16>(zm, pm, ones);
    }
    // zu, zm: elements 0 / 2 hold the sums of rows r0 / r1.  A row whose allowed squares carry less than 2^-6 of the unmasked
    // mass would lose precision in the 16-bit probabilities: rescale its masked probabilities by their maximum and redo the product
    if (__any_sync(0xFFFFFFFFu, fminf(zm[0], zm[2]) < 0.015625f)) {
      float x0 = 0.f, x1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        x0 = fmaxf(x0, fmaxf(sm[nt][0], sm[nt][1]));
        x1 = fmaxf(x1, fmaxf(sm[nt][2], sm[nt][3]));
      }
#pragma unroll
      for (int off = 1; off <= 2; off <<= 1) {
        x0 = fmaxf(x0, __shfl_xor_sync(0xFFFFFFFFu, x0, off));
        x1 = fmaxf(x1, __shfl_xor_sync(0xFFFFFFFFu, x1, off));
      }
      const float c0 = x0 > 0.f ? 1.0f / x0 : 0.f, c1 = x1 > 0.f ? 1.0f / x1 : 0.f;
#pragma unroll
      for (int nd = 0; nd < 2; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) om[nd][e] = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) zm[e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pm[4];
        pm[0] = pack2<FP16>(sm[2 * kk][0] * c0, sm[2 * kk][1] * c0);
        pm[1] = pack2<FP16>(sm[2 * kk][2] * c1, sm[2 * kk][3] * c1);
        pm[2] = pack2<FP16>(sm[2 * kk + 1][0] * c0, sm[2 * kk + 1][1] * c0);
        pm[3] = pack2<FP16>(sm[2 * kk + 1][2] * c1, sm[2 * kk + 1][3] * c1);
#pragma unroll
        for (int nd = 0; nd < 2; ++nd) mma16816<FP16>(om[nd], pm, vf[kk][nd]);
        mma16816<FP16>(zm, pm, ones);
      }
    }
    const float iu0 = __fdividef(1.0f, zu[0]), iu1 = __fdividef(1.0f, zu[2]);
    const float im0 = zm[0] > 0.f ? __fdividef(1.0f, zm[0]) : 0.f, im1 = zm[2] > 0.f ? __fdividef(1.0f, zm[2]) : 0.f;
#pragma unroll
    for (int nd = 0; nd < 2; ++nd) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float vu = ou[nd][e] * ((e & 2) ? iu1 : iu0), vm = om[nd][e] * ((e & 2) ? im1 : im0);
        o[e] = (mix > 0.0f && mix < 1.0f) ? (blend * vm + (1.0f - blend) * vu) : (mix >= 1.0f ? vm : vu);  // resnet.py:160-174
      }
      uint16_t* o0 = out + ((size_t)b * 64 + r0) * C + h * 16 + 8 * nd + 2 * t;
      uint16_t* o1 = out + ((size_t)b * 64 + r1) * C + h * 16 + 8 * nd + 2 * t;
      *reinterpret_cast<uint32_t*>(o0) = pack2<FP16>(o[0], o[1]);
      *reinterpret_cast<uint32_t*>(o1) = pack2<FP16>(o[2], o[3]);
    }
  }
}

template <bool FP16>
__global__ void __launch_bounds__(256, 2)
attention_tc_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ rel_bias, uint16_t* __restrict__ out, int B, int C, float mix) {
  __shared__ __align__(16) float s_bias[64 * BIAS_LD];   // rel_bias of the head * log2(e)
  __shared__ __align__(16) float s_mask[64 * BIAS_LD];   // 1 where the attack-pattern mask lets row attend to column, else 0
  const int h = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    s_bias[(i >> 6) * BIAS_LD + (i & 63)] = rel_bias ? rel_bias[(size_t)h * 4096 + i] * LOG2E : 0.0f;
    s_mask[(i >> 6) * BIAS_LD + (i & 63)] = attn_mask_tc(i >> 6, i & 63) ? 1.0f : 0.0f;
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const int ld = 3 * C;
  const int b_first = (blockIdx.y * 8 + warp) * ATT_BOARDS_PER_WARP;
#pragma unroll 1
  for (int bi = 0; bi < ATT_BOARDS_PER_WARP; ++bi) {
    const int b = b_first + bi;
    if (b >= B) break;
    const uint16_t* q_base = qkv + (size_t)b * 64 * ld + h * 16;
    const uint16_t* k_base = q_base + C;
    const uint16_t* v_base = q_base + 2 * C;
    // K^T as the col-major B operand of S = Q K^T: (k = d, n = key)
    uint32_t kf[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const uint16_t* kr = k_base + (size_t)(8 * nt + g) * ld;
      kf[nt][0] = __ldg(reinterpret_cast<const uint32_t*>(kr + 2 * t));
      kf[nt][1] = __ldg(reinterpret_cast<const uint32_t*>(kr + 2 * t + 8));
    }
    // V as the B operand of O = P V: (k = key, n = d): 8x8 tiles V[8j + g][8nd + 2t..] transposed in registers
    uint32_t vf[4][2][2];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int nd = 0; nd < 2; ++nd) {
        const uint32_t lo = __ldg(reinterpret_cast<const uint32_t*>(v_base + (size_t)(16 * kk + g) * ld + 8 * nd + 2 * t));
        const uint32_t hi = __ldg(reinterpret_cast<const uint32_t*>(v_base + (size_t)(16 * kk + 8 + g) * ld + 8 * nd + 2 * t));
        vf[kk][nd][0] = movmatrix_trans(lo);
        vf[kk][nd][1] = movmatrix_trans(hi);
      }
    // Q fragments of all four 16-row tiles up front: the loads overlap the K / V loads instead of stalling every tile
    uint32_t qall[4][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const uint16_t* qr = q_base + (size_t)(16 * mt + g) * ld + 2 * t;
      qall[mt][0] = __ldg(reinterpret_cast<const uint32_t*>(qr));
      qall[mt][1] = __ldg(reinterpret_cast<const uint32_t*>(qr + (size_t)8 * ld));
      qall[mt][2] = __ldg(reinterpret_cast<const uint32_t*>(qr + 8));
      qall[mt][3] = __ldg(reinterpret_cast<const uint32_t*>(qr + (size_t)8 * ld + 8));
    }
    attend_head<FP16>(kf, vf, qall, s_bias, s_mask, out, b, h, C, g, t, mix);
  }
}

// Staged variant (the default): the same arithmetic on the same fragments -- results are bit-identical to attention_tc_kernel --
// but the head's q / k / v slice of the NEXT board (64 tokens x 3 x 32 bytes = 6 KB) is brought into a per-warp shared-memory
// buffer with cp.async while the current board is being computed, so the ~1 us of dependent global-load latency at the top of
// every board (long-scoreboard stalls: 1.7 per issued instruction in profiles/r02_ncu_attention_tc_ln_res_gn_summary.txt)
// overlaps the tensor-core / exponential work.  One buffer per warp is enough: its fragments are in registers before the next
// copy is issued.  Buffer rows are 32 bytes, so the 32 lanes of a fragment load (row g, word t) hit 32 different banks.
static constexpr int ATT_STAGE_BYTES = 3 * 64 * 32;   // q, k, v rows of one (board, head)

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void stage_head(uint32_t buf, const uint16_t* __restrict__ qkv, int b, int h, int C, int lane) {
  // lane pair p = lane / 2 copies the two 16-byte halves of row (16 j + p) of q, k and v: one per-lane base pointer, all other
  // offsets are warp-uniform; shared-memory address = which * 2048 + token * 32 + half * 16 = which * 2048 + j * 512 + lane * 16
  const size_t ld = 3 * (size_t)C;
  const char* base = reinterpret_cast<const char*>(qkv + ((size_t)b * 64 + (lane >> 1)) * ld + h * 16) + (lane & 1) * 16;
  const uint32_t dst = buf + lane * 16;
#pragma unroll
  for (int which = 0; which < 3; ++which)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      cp_async16(dst + which * 2048 + j * 512, base + ((size_t)(16 * j) * ld + (size_t)which * C) * 2);
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <bool FP16>
__global__ void __launch_bounds__(256, 2)
attention_tc_staged_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ rel_bias, uint16_t* __restrict__ out, int B, int C,
                           float mix, int boards_per_warp) {
  extern __shared__ __align__(16) unsigned char att_smem[];
  float* s_bias = reinterpret_cast<float*>(att_smem);                 // rel_bias of the head * log2(e)
  float* s_mask = s_bias + 64 * BIAS_LD;                              // 1 where the attack-pattern mask lets row attend to column
  const int h = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint16_t* stage = reinterpret_cast<const uint16_t*>(att_smem + 2 * 64 * BIAS_LD * sizeof(float) + warp * ATT_STAGE_BYTES);
  const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(stage);
  const int b_first = (blockIdx.y * 8 + warp) * boards_per_warp;
  const int b_end = min(B, b_first + boards_per_warp);
  if (b_first < b_end) stage_head(stage_addr, qkv, b_first, h, C, lane);   // in flight while the tables are being staged
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    s_bias[(i >> 6) * BIAS_LD + (i & 63)] = rel_bias ? rel_bias[(size_t)h * 4096 + i] * LOG2E : 0.0f;
    s_mask[(i >> 6) * BIAS_LD + (i & 63)] = attn_mask_tc(i >> 6, i & 63) ? 1.0f : 0.0f;
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const uint16_t* sq = stage, * sk = stage + 64 * 16, * sv = stage + 2 * 64 * 16;
#pragma unroll 1
  for (int b = b_first; b < b_end; ++b) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    uint32_t kf[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const uint16_t* kr = sk + (8 * nt + g) * 16;
      kf[nt][0] = *reinterpret_cast<const uint32_t*>(kr + 2 * t);
      kf[nt][1] = *reinterpret_cast<const uint32_t*>(kr + 2 * t + 8);
    }
    uint32_t vf[4][2][2];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int nd = 0; nd < 2; ++nd) {
        const uint32_t lo = *reinterpret_cast<const uint32_t*>(sv + (16 * kk + g) * 16 + 8 * nd + 2 * t);
        const uint32_t hi = *reinterpret_cast<const uint32_t*>(sv + (16 * kk + 8 + g) * 16 + 8 * nd + 2 * t);
        vf[kk][nd][0] = movmatrix_trans(lo);
        vf[kk][nd][1] = movmatrix_trans(hi);
      }
    uint32_t qall[4][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const uint16_t* qr = sq + (16 * mt + g) * 16 + 2 * t;
      qall[mt][0] = *reinterpret_cast<const uint32_t*>(qr);
      qall[mt][1] = *reinterpret_cast<const uint32_t*>(qr + 8 * 16);
      qall[mt][2] = *reinterpret_cast<const uint32_t*>(qr + 8);
      qall[mt][3] = *reinterpret_cast<const uint32_t*>(qr + 8 * 16 + 8);
    }
    __syncwarp();                                                        // every lane has its fragments: the buffer is free
    if (b + 1 < b_end) stage_head(stage_addr, qkv, b + 1, h, C, lane);
    attend_head<FP16>(kf, vf, qall, s_bias, s_mask, out, b, h, C, g, t, mix);
  }
}

// qkv: half [B][64][3C] (channel = which*C + head*16 + d), out: half [B][64][C]; head_dim must be 16
int nn_attention_tc(const void* qkv_half, const float* rel_bias, void* out_half, int B, int C, int heads, float mix, cudaStream_t s) {
  if (C != heads * 16) { m0_set_error("attention_tc: head_dim must be 16 (C=%d heads=%d)", C, heads); return M0_ERR_ARG; }
  // M0_ATT_STAGED=0 selects the direct-load kernel (A/B and fallback).  Boards per warp of the staged kernel (M0_ATT_BOARDS overrides):
  // as many as keep every warp slot of the device (two blocks of 8 warps per SM) busy for about four rounds, at most 8 -- 8 at the
  // self-play batch of 4096 boards (393 us against 430 us with 4 and 494 us for the direct-load kernel inside the running forward,
  // profiles/r02_ab_attention_staged.json), 1 at the 96-row batches of a single game, where one round of warps is the latency.
  static const int staged = [] { const char* e = getenv("M0_ATT_STAGED"); return e ? atoi(e) : 1; }();
  static const int bpw_env = [] { const char* e = getenv("M0_ATT_BOARDS"); return e ? atoi(e) : 0; }();
  static const int warp_slots = [] {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return 16 * sms;
  }();
  int bpw = bpw_env;
  if (bpw <= 0) {
    bpw = (int)(((long long)B * heads) / (4LL * warp_slots));
    bpw = bpw < 1 ? 1 : bpw > 8 ? 8 : bpw;
  }
  if (staged) {
    const int smem = 2 * 64 * BIAS_LD * (int)sizeof(float) + 8 * ATT_STAGE_BYTES;
    static unsigned long long granted = 0;   // bit d: device d's kernels have the shared-memory grant (it is per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !((granted >> dev) & 1ull)) {
      cudaError_t e = cudaFuncSetAttribute(attention_tc_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) { m0_set_error("attention_tc: cannot reserve %d bytes of shared memory (%s)", smem, cudaGetErrorString(e)); return M0_ERR_CUDA; }
      if (dev < 64) granted |= 1ull << dev;
    }
    dim3 grid(heads, (B + 8 * bpw - 1) / (8 * bpw));
    if (nn_half_format()) attention_tc_staged_kernel<true><<<grid, 256, smem, s>>>((const uint16_t*)qkv_half, rel_bias, (uint16_t*)out_half, B, C, mix, bpw);
    else attention_tc_staged_kernel<false><<<grid, 256, smem, s>>>((const uint16_t*)qkv_half, rel_bias, (uint16_t*)out_half, B, C, mix, bpw);
    return m0_check_launch("attention_tc_staged");
  }
  const int per_block = 8 * ATT_BOARDS_PER_WARP;
  dim3 grid(heads, (B + per_block - 1) / per_block);
  if (nn_half_format()) attention_tc_kernel<true><<<grid, 256, 0, s>>>((const uint16_t*)qkv_half, rel_bias, (uint16_t*)out_half, B, C, mix);
  else attention_tc_kernel<false><<<grid, 256, 0, s>>>((const uint16_t*)qkv_half, rel_bias, (uint16_t*)out_half, B, C, mix);
  return m0_check_launch("attention_tc");
}

}  // namespace m0
