// ChessAttention core (azchess/model/resnet.py:141-174) on tensor cores: one warp per (board, head).
// S = Q K^T / sqrt(d) + rel_bias, clamp +-50, masked softmax (mask == 0 -> -1e4) and unmasked softmax,
// O = (1 - mix) * P_masked V + mix * P_unmasked V.  64 tokens x head_dim 16: the whole head lives in registers as
// mma.sync.m16n8k16 fragments (fp16 or bf16 operands, fp32 accumulate).  Inputs / outputs are the half-precision
// buffers of the tensor-core GEMMs.
//
// The kernel is instruction-bound (2 x 4096 probabilities per head), so the per-element work is kept minimal:
//   * the 64x64 bias of the head is staged once per block, pre-multiplied by log2(e), and a block walks 32 boards;
//   * the attack-pattern mask (resnet.py:105-129) of the 128 score elements a lane owns is 4 registers of bits,
//     computed once per warp;
//   * ONE exponential per element: softmax is shift invariant, so the masked probabilities reuse
//     2^(s - max_all) scaled by 2^(max_all - max_allowed) per row (masked-out entries are exactly 0 in the
//     reference too: exp(-1e4 - max) underflows).  Rows whose allowed maximum lies more than 80 octaves below the
//     row maximum take a second exponential instead (warp-uniform branch);
//   * V fragments come from 32-bit loads + movmatrix.trans instead of 16-bit gathers.
// Blocks are ordered head-fastest so that the 20 heads of a board (one 1920-byte row of qkv per token) are read
// by co-scheduled blocks while the lines are still in L2.
#include "nn.cuh"
#include <cuda_fp16.h>

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

static constexpr int BIAS_LD = 72;          // padded row stride of the staged bias (floats)
static constexpr int ATT_BOARDS_PER_WARP = 4;
static constexpr float LOG2E = 1.4426950408889634f;

template <bool FP16>
__global__ void __launch_bounds__(256, 2)
attention_tc_kernel(const uint16_t* __restrict__ qkv, const float* __restrict__ rel_bias, uint16_t* __restrict__ out, int B, int C, float mix) {
  __shared__ __align__(16) float s_bias[64 * BIAS_LD];
  const int h = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) s_bias[(i >> 6) * BIAS_LD + (i & 63)] = rel_bias ? rel_bias[(size_t)h * 4096 + i] * LOG2E : 0.0f;
  const int g = lane >> 2, t = lane & 3;
  // allowed[mt] bit (nt*4 + e): element (row = 16mt + g + 8*(e>>1), col = 8nt + 2t + (e&1)) may attend
  uint32_t allowed[4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    uint32_t w = 0;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (attn_mask_tc(16 * mt + g + ((e & 2) ? 8 : 0), 8 * nt + 2 * t + (e & 1))) w |= 1u << (nt * 4 + e);
    allowed[mt] = w;
  }
  __syncthreads();
  const int ld = 3 * C;
  const float blend = 1.0f - mix;
  const float kscale = 0.25f * LOG2E, kclamp = 50.0f * LOG2E;
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
#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      const int r0 = 16 * mt + g, r1 = r0 + 8;
      uint32_t qa[4];
      qa[0] = __ldg(reinterpret_cast<const uint32_t*>(q_base + (size_t)r0 * ld + 2 * t));
      qa[1] = __ldg(reinterpret_cast<const uint32_t*>(q_base + (size_t)r1 * ld + 2 * t));
      qa[2] = __ldg(reinterpret_cast<const uint32_t*>(q_base + (size_t)r0 * ld + 2 * t + 8));
      qa[3] = __ldg(reinterpret_cast<const uint32_t*>(q_base + (size_t)r1 * ld + 2 * t + 8));
      const uint32_t am = mt == 0 ? allowed[0] : mt == 1 ? allowed[1] : mt == 2 ? allowed[2] : allowed[3];
      float su[8][4];
      float mu0 = -INFINITY, mu1 = -INFINITY, mm0 = -INFINITY, mm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816<FP16>(acc, qa, kf[nt]);
        const float2 b0 = *reinterpret_cast<const float2*>(s_bias + r0 * BIAS_LD + 8 * nt + 2 * t);
        const float2 b1 = *reinterpret_cast<const float2*>(s_bias + r1 * BIAS_LD + 8 * nt + 2 * t);
        const float bb[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = fmaf(acc[e], kscale, bb[e]);
          v = fminf(fmaxf(v, -kclamp), kclamp);
          su[nt][e] = v;
          const float vm = (am >> (nt * 4 + e)) & 1u ? v : -INFINITY;
          if (e & 2) { mu1 = fmaxf(mu1, v); mm1 = fmaxf(mm1, vm); }
          else { mu0 = fmaxf(mu0, v); mm0 = fmaxf(mm0, vm); }
        }
      }
#pragma unroll
      for (int off = 1; off <= 2; off <<= 1) {
        mu0 = fmaxf(mu0, __shfl_xor_sync(0xFFFFFFFFu, mu0, off)); mu1 = fmaxf(mu1, __shfl_xor_sync(0xFFFFFFFFu, mu1, off));
        mm0 = fmaxf(mm0, __shfl_xor_sync(0xFFFFFFFFu, mm0, off)); mm1 = fmaxf(mm1, __shfl_xor_sync(0xFFFFFFFFu, mm1, off));
      }
      // every row may attend to itself, so mm is finite
      const float gap0 = mu0 - mm0, gap1 = mu1 - mm1;
      const bool far = __any_sync(0xFFFFFFFFu, fmaxf(gap0, gap1) > 80.0f);
      const float c0 = fast_ex2(fminf(gap0, 80.0f)), c1 = fast_ex2(fminf(gap1, 80.0f));
      float zu0 = 0.f, zu1 = 0.f, zm0 = 0.f, zm1 = 0.f;
      float sm[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float s = su[nt][e];
          const float pu = fast_ex2(s - ((e & 2) ? mu1 : mu0));
          float pm;
          if (far) pm = fast_ex2(s - ((e & 2) ? mm1 : mm0));
          else pm = pu * ((e & 2) ? c1 : c0);
          pm = (am >> (nt * 4 + e)) & 1u ? pm : 0.0f;
          su[nt][e] = pu;
          sm[nt][e] = pm;
          if (e & 2) { zu1 += pu; zm1 += pm; } else { zu0 += pu; zm0 += pm; }
        }
#pragma unroll
      for (int off = 1; off <= 2; off <<= 1) {
        zu0 += __shfl_xor_sync(0xFFFFFFFFu, zu0, off); zu1 += __shfl_xor_sync(0xFFFFFFFFu, zu1, off);
        zm0 += __shfl_xor_sync(0xFFFFFFFFu, zm0, off); zm1 += __shfl_xor_sync(0xFFFFFFFFu, zm1, off);
      }
      float ou[2][4], om[2][4];
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
      }
      const float iu0 = __fdividef(1.0f, zu0), iu1 = __fdividef(1.0f, zu1), im0 = __fdividef(1.0f, zm0), im1 = __fdividef(1.0f, zm1);
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
}

// qkv: half [B][64][3C] (channel = which*C + head*16 + d), out: half [B][64][C]; head_dim must be 16
int nn_attention_tc(const void* qkv_half, const float* rel_bias, void* out_half, int B, int C, int heads, float mix, cudaStream_t s) {
  if (C != heads * 16) { m0_set_error("attention_tc: head_dim must be 16 (C=%d heads=%d)", C, heads); return M0_ERR_ARG; }
  const int per_block = 8 * ATT_BOARDS_PER_WARP;
  dim3 grid(heads, (B + per_block - 1) / per_block);
  if (nn_half_format()) attention_tc_kernel<true><<<grid, 256, 0, s>>>((const uint16_t*)qkv_half, rel_bias, (uint16_t*)out_half, B, C, mix);
  else attention_tc_kernel<false><<<grid, 256, 0, s>>>((const uint16_t*)qkv_half, rel_bias, (uint16_t*)out_half, B, C, mix);
  return m0_check_launch("attention_tc");
}

}  // namespace m0
