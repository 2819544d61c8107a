// SSL target maps for a batch of positions (azchess/ssl_algorithms.py create_enhanced_ssl_targets, called per played ply by
// selfplay_worker, internal.py:460-466): one thread computes the 17 bit masks of a position (ssl_core.cuh), the block then writes
// the float32 maps with 16-byte coalesced stores.
#include "m0_common.cuh"
#include "ssl_core.cuh"

namespace m0 {

static constexpr int SSL_THREADS = 128;
static constexpr int SSL_WORDS = 17;   // 13 piece masks, threat, fork, control +, control -

__device__ __forceinline__ float4 bits_to_f4(u32 bits) {
  return make_float4((bits & 1) ? 1.0f : 0.0f, (bits & 2) ? 1.0f : 0.0f, (bits & 4) ? 1.0f : 0.0f, (bits & 8) ? 1.0f : 0.0f);
}

__global__ void __launch_bounds__(SSL_THREADS)
ssl_targets_kernel(const u64* __restrict__ pos, int n, float* __restrict__ piece, float* __restrict__ threat, float* __restrict__ pin,
                   float* __restrict__ fork, float* __restrict__ control) {
  __shared__ u64 s_mask[SSL_THREADS * SSL_WORDS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base = blockIdx.x * SSL_THREADS;
  const int nb = min(SSL_THREADS, n - base);
  if (tid < nb) {
    Position p = load_position(pos + (size_t)(base + tid) * POSITION_WORDS);
    SslMasks m;
    ssl_masks(p, m);
    u64* row = s_mask + tid * SSL_WORDS;
#pragma unroll
    for (int k = 0; k < 13; ++k) row[k] = m.piece[k];
    row[13] = m.threat;
    row[14] = m.fork;
    row[15] = m.ctrl_pos;
    row[16] = m.ctrl_neg;
  }
  __syncthreads();
  // each warp walks its 32 positions; a map of 64 floats = 16 float4 chunks (chunk c = squares 4c .. 4c+3 in plane order)
  for (int j = 0; j < 32; ++j) {
    const int t = warp * 32 + j;
    if (t >= nb) break;
    const u64* row = s_mask + t * SSL_WORDS;
    const size_t g = (size_t)(base + t);
    if (piece) {
      float4* o = reinterpret_cast<float4*>(piece + g * 13 * 64);
      for (int c = lane; c < 13 * 16; c += 32) st_global_cs_f4(o + c, bits_to_f4((u32)(row[c >> 4] >> ((c & 15) * 4)) & 15u));
    }
    if (lane < 16) {
      const int sh = lane * 4;
      if (threat) st_global_cs_f4(reinterpret_cast<float4*>(threat + g * 64) + lane, bits_to_f4((u32)(row[13] >> sh) & 15u));
      if (pin) st_global_cs_f4(reinterpret_cast<float4*>(pin + g * 64) + lane, make_float4(0.f, 0.f, 0.f, 0.f));
      if (fork) st_global_cs_f4(reinterpret_cast<float4*>(fork + g * 64) + lane, bits_to_f4((u32)(row[14] >> sh) & 15u));
      if (control) {
        const float4 a = bits_to_f4((u32)(row[15] >> sh) & 15u), b = bits_to_f4((u32)(row[16] >> sh) & 15u);
        st_global_cs_f4(reinterpret_cast<float4*>(control + g * 64) + lane, make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w));
      }
    }
  }
}

}  // namespace m0

using namespace m0;

extern "C" int m0_ssl_targets(const uint64_t* d_pos, int n, float* d_piece, float* d_threat, float* d_pin, float* d_fork, float* d_control,
                              void* stream) {
  if (n < 0 || (n > 0 && !d_pos)) { m0_set_error("m0_ssl_targets: invalid argument"); return M0_ERR_ARG; }
  if (n == 0) return M0_OK;
  ssl_targets_kernel<<<(n + SSL_THREADS - 1) / SSL_THREADS, SSL_THREADS, 0, (cudaStream_t)stream>>>(d_pos, n, d_piece, d_threat, d_pin, d_fork,
                                                                                                  d_control);
  return m0_check_launch("m0_ssl_targets");
}
