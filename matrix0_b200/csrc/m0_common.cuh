// Shared helpers for the matrix0_b200 CUDA translation units: error reporting for the C ABI,
// packed-position load/store and cache-hinted vector memory operations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "chess_core.cuh"

// ---- C-ABI error reporting (include/matrix0_b200.h: m0_last_error) --------------------------------
extern "C" const char* m0_last_error(void);
void m0_set_error(const char* fmt, ...);
// returns 0 when the preceding launch was accepted, a negative M0_ERR_* code otherwise
int m0_check_launch(const char* what);
int m0_check_cuda(cudaError_t e, const char* what);

#define M0_OK 0
#define M0_ERR_CUDA (-1)
#define M0_ERR_ARG (-2)
#define M0_ERR_STATE (-3)
#define M0_ERR_CAPACITY (-4)

#define M0_CUDA_TRY(expr)                                   \
  do {                                                      \
    int _rc = m0_check_cuda((expr), #expr);                 \
    if (_rc != 0) return _rc;                               \
  } while (0)

namespace m0 {

// Entry points that own device memory (engine / network handles) run on THEIR device and leave the caller's current device as they
// found it: the host side is PyTorch, which tracks the current device itself (a handle destroyed by the garbage collector inside a
// `with torch.cuda.device(1)` block must not move the thread to device 0 behind its back).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) err = cudaSetDevice(device);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

M0_HD Position load_position(const u64* w) {
  Position p;
  p.pawns = w[0]; p.knights = w[1]; p.bishops = w[2]; p.rooks = w[3]; p.queens = w[4]; p.kings = w[5];
  p.occ_w = w[6]; p.occ_b = w[7]; p.state = w[8];
  return p;
}
M0_HD void store_position(u64* w, const Position& p) {
  w[0] = p.pawns; w[1] = p.knights; w[2] = p.bishops; w[3] = p.rooks; w[4] = p.queens; w[5] = p.kings;
  w[6] = p.occ_w; w[7] = p.occ_b; w[8] = p.state;
}

#if defined(__CUDACC__)
// streaming (evict-first) 16-byte store: outputs are written once and not re-read by this kernel
__device__ __forceinline__ void st_global_cs_f4(float4* ptr, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ u64 ld_global_nc_u64(const u64* ptr) {
  u64 v;
  asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(ptr));
  return v;
}
// One warp writes the 19x8x8 float32 planes of one position: 304 float4 chunks, 16 per plane.
__device__ __forceinline__ void warp_write_planes(const Position& p, float* __restrict__ out, int lane) {
  float4* o4 = reinterpret_cast<float4*>(out);
#pragma unroll 2
  for (int c = lane; c < 19 * 16; c += 32) {
    int plane = c >> 4;
    float4 v;
    if (plane < 12) {
      int row = (c & 15) >> 1, col0 = (c & 1) * 4;
      u32 bits = (u32)(piece_plane_bb(p, plane) >> ((7 - row) * 8 + col0)) & 15u;
      v.x = (bits & 1) ? 1.0f : 0.0f;
      v.y = (bits & 2) ? 1.0f : 0.0f;
      v.z = (bits & 4) ? 1.0f : 0.0f;
      v.w = (bits & 8) ? 1.0f : 0.0f;
    } else {
      float f = const_plane_value(p, plane);
      v = make_float4(f, f, f, f);
    }
    st_global_cs_f4(o4 + c, v);
  }
}
#endif

}  // namespace m0
