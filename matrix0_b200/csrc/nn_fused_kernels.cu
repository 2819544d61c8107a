// Memory-bound glue kernels of the tensor-core forward, fused so that the residual stream is read and
// written once per residual block.
#include "nn.cuh"
#include <cuda_fp16.h>

namespace m0 {

__device__ __forceinline__ float fk_act(float x, int act) {
  switch (act) {
    case ACT_RELU: return x > 0.0f ? x : 0.0f;
    case ACT_SILU: return x / (1.0f + __expf(-x));
    default: return x;
  }
}
__device__ __forceinline__ __nv_bfloat16 fk_half(float x, int fp16) {
  if (fp16) {
    __half h = __float2half_rn(x);
    return *reinterpret_cast<__nv_bfloat16*>(&h);
  }
  return __float2bfloat16(x);
}

// x_new = x + conv_out * gate   (SE excitation + residual add, resnet.py:68-80)
// a_out = half(act(GroupNorm(x_new)))   (bn1 + activation of the NEXT pre-activation block, resnet.py:46-47), or half(x_new) when
// gamma is null, or nothing when a_out is null.
// One block per board.  Thread (rg, q) owns 4 consecutive channels (one float4) of 16 rows: x and conv_out are read once with
// 16-byte loads, x_new / a_out written once.  GroupNorm statistics: 4 threads (= 16 channels) x 4 row groups per group.
__global__ void __launch_bounds__(320, 2)
se_apply_gn_kernel(const float* __restrict__ conv_out, const float* __restrict__ gate, float* __restrict__ x,
                   const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ a_out, int C, int act, int fp16) {
  __shared__ float s_part[4][80][2];
  const int b = blockIdx.x;
  const int nq = C >> 2;                       // float4 columns (<= 80)
  const int q = threadIdx.x % nq, rg = threadIdx.x / nq;   // rg < 4 when blockDim == 4 * nq
  const size_t base = (size_t)b * 64 * C + (size_t)rg * 16 * C + 4 * q;
  float4 e = make_float4(1.f, 1.f, 1.f, 1.f);
  if (gate) e = *reinterpret_cast<const float4*>(gate + (size_t)b * C + 4 * q);
  float4 v[16];
  if (conv_out) {
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = __ldg(reinterpret_cast<const float4*>(conv_out + base + (size_t)r * C));
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float4 xv = *reinterpret_cast<const float4*>(x + base + (size_t)r * C);
      v[r].x = fmaf(v[r].x, e.x, xv.x); v[r].y = fmaf(v[r].y, e.y, xv.y);
      v[r].z = fmaf(v[r].z, e.z, xv.z); v[r].w = fmaf(v[r].w, e.w, xv.w);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) *reinterpret_cast<float4*>(x + base + (size_t)r * C) = v[r];
  } else {   // plain GroupNorm of x: nothing to add, nothing to write back
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = *reinterpret_cast<const float4*>(x + base + (size_t)r * C);
  }
  if (!a_out) return;
  __nv_bfloat16* ao = a_out + base;
  if (!gamma) {  // the next consumer is the attention qkv GEMM: it takes the raw residual stream in half precision
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      __nv_bfloat16 h4[4] = {fk_half(v[r].x, fp16), fk_half(v[r].y, fp16), fk_half(v[r].z, fp16), fk_half(v[r].w, fp16)};
      *reinterpret_cast<uint2*>(ao + (size_t)r * C) = *reinterpret_cast<uint2*>(h4);
    }
    return;
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 16; ++r) s += (v[r].x + v[r].y) + (v[r].z + v[r].w);
  s += __shfl_xor_sync(0xFFFFFFFFu, s, 1);
  s += __shfl_xor_sync(0xFFFFFFFFu, s, 2);   // 4 adjacent threads = one group of 16 channels (nq % 4 == 0, warps hold whole groups)
  s_part[rg][q][0] = s;
  __syncthreads();
  const float mean = (s_part[0][q][0] + s_part[1][q][0] + s_part[2][q][0] + s_part[3][q][0]) * (1.0f / 1024.0f);
  float d2 = 0.f;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    float a0 = v[r].x - mean, a1 = v[r].y - mean, a2 = v[r].z - mean, a3 = v[r].w - mean;
    d2 = fmaf(a0, a0, d2); d2 = fmaf(a1, a1, d2); d2 = fmaf(a2, a2, d2); d2 = fmaf(a3, a3, d2);
  }
  d2 += __shfl_xor_sync(0xFFFFFFFFu, d2, 1);
  d2 += __shfl_xor_sync(0xFFFFFFFFu, d2, 2);
  s_part[rg][q][1] = d2;
  __syncthreads();
  const float var = (s_part[0][q][1] + s_part[1][q][1] + s_part[2][q][1] + s_part[3][q][1]) * (1.0f / 1024.0f);
  const float rstd = rsqrtf(var + 1e-5f);
  const float4 gm = *reinterpret_cast<const float4*>(gamma + 4 * q), bt = *reinterpret_cast<const float4*>(beta + 4 * q);
  const float g0 = gm.x * rstd, g1 = gm.y * rstd, g2 = gm.z * rstd, g3 = gm.w * rstd;
  const float b0 = bt.x - mean * g0, b1 = bt.y - mean * g1, b2 = bt.z - mean * g2, b3 = bt.w - mean * g3;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    __nv_bfloat16 h4[4] = {fk_half(fk_act(fmaf(v[r].x, g0, b0), act), fp16), fk_half(fk_act(fmaf(v[r].y, g1, b1), act), fp16),
                           fk_half(fk_act(fmaf(v[r].z, g2, b2), act), fp16), fk_half(fk_act(fmaf(v[r].w, g3, b3), act), fp16)};
    *reinterpret_cast<uint2*>(ao + (size_t)r * C) = *reinterpret_cast<uint2*>(h4);
  }
}

// NCHW float32 planes [B][P][8][8] -> NHWC half [B][64][64] with channels P..63 zero (stem input of the tensor-core path)
__global__ void planes_to_nhwc_half_kernel(const float* __restrict__ planes, __nv_bfloat16* __restrict__ out, int B, int P, int fp16) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*64*64
  if (i >= (size_t)B * 64 * 64) return;
  const int c = (int)(i & 63);
  const int sq = (int)((i >> 6) & 63);
  const size_t b = i >> 12;
  out[i] = c < P ? fk_half(planes[(b * P + c) * 64 + sq], fp16) : fk_half(0.0f, fp16);
}

// SE excitation for 16 boards per block (resnet.py:61-64): s = avgpool, h = act(W1 s + b1), gate = sigmoid(W2 h + b2).
// pool: half-board column sums [B][2][C]; w1t [C][hid] and w2t [hid][C] are transposed so that threads read them coalesced.
__global__ void __launch_bounds__(320)
se_gate_kernel(const float* __restrict__ pool, const float* __restrict__ w1t, const float* __restrict__ b1, const float* __restrict__ w2t,
               const float* __restrict__ b2, float* __restrict__ gate, int B, int C, int hid, int act) {
  extern __shared__ float sm[];
  float* s_pool = sm;              // [16][C]
  float* s_hid = sm + 16 * C;      // [16][hid]
  const int b0 = blockIdx.x * 16, t = threadIdx.x;
  const int nb = min(16, B - b0);
  for (int i = t; i < nb * C; i += blockDim.x) {
    const int j = i / C, c = i - j * C;
    const float* pp = pool + (size_t)(b0 + j) * 2 * C;
    s_pool[j * C + c] = (pp[c] + pp[C + c]) * (1.0f / 64.0f);
  }
  __syncthreads();
  for (int o = t; o < nb * hid; o += blockDim.x) {
    const int j = o / hid, u = o - j * hid;
    float h = b1[u];
    const float* sp = s_pool + j * C;
    for (int c = 0; c < C; ++c) h = fmaf(w1t[c * hid + u], sp[c], h);
    s_hid[j * hid + u] = fk_act(h, act);
  }
  __syncthreads();
  for (int c = t; c < C; c += blockDim.x) {
    float z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = b2[c];
    for (int u = 0; u < hid; ++u) {
      const float w = w2t[u * C + c];
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = fmaf(w, s_hid[j * hid + u], z[j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nb) gate[(size_t)(b0 + j) * C + c] = 1.0f / (1.0f + __expf(-z[j]));
  }
}

int nn_se_gate(const float* pool, const float* w1t, const float* b1, const float* w2t, const float* b2, float* gate, int B, int C, int hid, int act,
               cudaStream_t s) {
  se_gate_kernel<<<(B + 15) / 16, 320, (size_t)16 * (C + hid) * sizeof(float), s>>>(pool, w1t, b1, w2t, b2, gate, B, C, hid, act);
  return m0_check_launch("se_gate");
}

int nn_se_apply_gn(const float* conv_out, const float* gate, float* x, const float* gamma, const float* beta, __nv_bfloat16* a_out, int B,
                   int C, int act, cudaStream_t s) {
  if (C > 320 || C % 32 != 0) { m0_set_error("se_apply_gn: unsupported channel count %d", C); return M0_ERR_ARG; }
  if (C % 16 != 0) { m0_set_error("se_apply_gn: channels must be a multiple of 16"); return M0_ERR_ARG; }
  se_apply_gn_kernel<<<B, C, 0, s>>>(conv_out, gate, x, gamma, beta, a_out, C, act, nn_half_format());
  return m0_check_launch("se_apply_gn");
}
int nn_planes_to_nhwc_half(const float* planes, __nv_bfloat16* out, int B, int P, cudaStream_t s) {
  const size_t total = (size_t)B * 64 * 64;
  planes_to_nhwc_half_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(planes, out, B, P, nn_half_format());
  return m0_check_launch("planes_to_nhwc_half");
}

}  // namespace m0
