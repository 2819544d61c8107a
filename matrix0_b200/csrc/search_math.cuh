// Bit-exact arithmetic of the reference search (azchess/mcts.py) restated as single-rounding IEEE
// operations.  The reference computes PUCT scores and backups in Python floats (IEEE double, one
// rounding per operator) and priors in numpy/torch float32; to reproduce its visit counts exactly
// no operation here may be contracted into an FMA, hence the explicit *_rn intrinsics on the device
// (the host build used by tests/hostcheck is compiled with -ffp-contract=off).
#pragma once
#include "chess_core.cuh"

namespace m0 {

#if defined(__CUDA_ARCH__)
M0_HD double d_add(double a, double b) { return __dadd_rn(a, b); }
M0_HD double d_sub(double a, double b) { return __dsub_rn(a, b); }
M0_HD double d_mul(double a, double b) { return __dmul_rn(a, b); }
M0_HD double d_div(double a, double b) { return __ddiv_rn(a, b); }
M0_HD double d_sqrt(double a) { return __dsqrt_rn(a); }
M0_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
M0_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
M0_HD float f_div(float a, float b) { return __fdiv_rn(a, b); }
M0_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
#else
}  // namespace m0
#include <math.h>
namespace m0 {
M0_HD double d_add(double a, double b) { return a + b; }
M0_HD double d_sub(double a, double b) { return a - b; }
M0_HD double d_mul(double a, double b) { return a * b; }
M0_HD double d_div(double a, double b) { return a / b; }
M0_HD double d_sqrt(double a) { return sqrt(a); }
M0_HD float f_add(float a, float b) { return a + b; }
M0_HD float f_sub(float a, float b) { return a - b; }
M0_HD float f_div(float a, float b) { return a / b; }
M0_HD float f_mul(float a, float b) { return a * b; }
#endif

// azchess/mcts.py:878-881:  u = eff_cpuct * child.prior * (math.sqrt(parent_visits) / (1.0 + child.n));
// score = q + u   (left-to-right products, then the parenthesised quotient)
M0_HD double puct_score(double q, double cpuct, double prior, double sqrt_parent_visits, int child_n) {
  double u = d_mul(d_mul(cpuct, prior), d_div(sqrt_parent_visits, d_add(1.0, (double)child_n)));
  return d_add(q, u);
}

// azchess/mcts.py:946-953 repeated `times` times for one path entry that occurs once in the path:
//   n += 1; w += v; q = w / n   (w is accumulated by repeated addition, not n*v -- SURVEY Q7)
M0_HD void backup_repeated(int& n, double& w, double& q, double v, int times) {
  for (int i = 0; i < times; ++i) w = d_add(w, v);
  n += times;
  q = d_div(w, (double)n);
}

// numpy's float32 add.reduce (FLOAT_pairwise_sum, numpy/_core/src/umath/loops_utils.h.src) over a
// contiguous array -- what `lp.sum()` at azchess/mcts.py:206 evaluates.
M0_HD float np_pairwise_block_f32(const float* a, int n) {  // n <= 128 (PW_BLOCKSIZE)
  if (n < 8) {
    float res = 0.0f;
    for (int i = 0; i < n; ++i) res = f_add(res, a[i]);
    return res;
  }
  float r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
  int i;
  for (i = 8; i < n - (n % 8); i += 8) {
    r0 = f_add(r0, a[i + 0]); r1 = f_add(r1, a[i + 1]); r2 = f_add(r2, a[i + 2]); r3 = f_add(r3, a[i + 3]);
    r4 = f_add(r4, a[i + 4]); r5 = f_add(r5, a[i + 5]); r6 = f_add(r6, a[i + 6]); r7 = f_add(r7, a[i + 7]);
  }
  float res = f_add(f_add(f_add(r0, r1), f_add(r2, r3)), f_add(f_add(r4, r5), f_add(r6, r7)));
  for (; i < n; ++i) res = f_add(res, a[i]);
  return res;
}
template <int DEPTH>
M0_HD float np_pairwise_sum_rec(const float* a, int n) {
  if (n <= 128) return np_pairwise_block_f32(a, n);
  int n2 = n / 2;
  n2 -= n2 % 8;
  return f_add(np_pairwise_sum_rec<DEPTH - 1>(a, n2), np_pairwise_sum_rec<DEPTH - 1>(a + n2, n - n2));
}
template <>
M0_HD float np_pairwise_sum_rec<0>(const float* a, int n) { return np_pairwise_block_f32(a, n); }
// three halvings: exact while every part is <= 128 after them -- the larger half of n is n - (n/2 - (n/2) % 8) <= n/2 + 8, so any
// n <= 900 (and n = 1024) qualifies; legal-move lists have n <= 256 (tests/test_hostcheck.py checks every length up to 300)
M0_HD float np_pairwise_sum_f32(const float* a, int n) { return np_pairwise_sum_rec<3>(a, n); }
// six halvings: any n <= 7000 (and n = 8192) by the same bound -- the whole 4672-entry policy vector
M0_HD float np_pairwise_sum_f32_big(const float* a, int n) { return np_pairwise_sum_rec<6>(a, n); }

// the same reduction for float64 arrays (DOUBLE_pairwise_sum): `dist.sum()` at azchess/mcts.py:184 after the float64 noise was added
M0_HD double np_pairwise_block_f64(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = d_add(res, a[i]);
    return res;
  }
  double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
  int i;
  for (i = 8; i < n - (n % 8); i += 8) {
    r0 = d_add(r0, a[i + 0]); r1 = d_add(r1, a[i + 1]); r2 = d_add(r2, a[i + 2]); r3 = d_add(r3, a[i + 3]);
    r4 = d_add(r4, a[i + 4]); r5 = d_add(r5, a[i + 5]); r6 = d_add(r6, a[i + 6]); r7 = d_add(r7, a[i + 7]);
  }
  double res = d_add(d_add(d_add(r0, r1), d_add(r2, r3)), d_add(d_add(r4, r5), d_add(r6, r7)));
  for (; i < n; ++i) res = d_add(res, a[i]);
  return res;
}
template <int DEPTH>
M0_HD double np_pairwise_sum_rec_f64(const double* a, int n) {
  if (n <= 128) return np_pairwise_block_f64(a, n);
  int n2 = n / 2;
  n2 -= n2 % 8;
  return d_add(np_pairwise_sum_rec_f64<DEPTH - 1>(a, n2), np_pairwise_sum_rec_f64<DEPTH - 1>(a + n2, n - n2));
}
template <>
M0_HD double np_pairwise_sum_rec_f64<0>(const double* a, int n) { return np_pairwise_block_f64(a, n); }
M0_HD double np_pairwise_sum_f64(const double* a, int n) { return np_pairwise_sum_rec_f64<3>(a, n); }
M0_HD double np_pairwise_sum_f64_big(const double* a, int n) { return np_pairwise_sum_rec_f64<6>(a, n); }

// azchess/mcts.py:948: v = max(-1.0, min(1.0, float(value)))  with Python's min/max NaN behaviour
M0_HD double py_clip_unit(double x) {
  double v = (x < 1.0) ? x : 1.0;
  return (v > -1.0) ? v : -1.0;
}

}  // namespace m0
