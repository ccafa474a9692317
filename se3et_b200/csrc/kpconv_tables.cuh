// Octahedral-group tables of KPConvInterSO3 for kanchor = 6, quotient_factor = 4, K = 15 (blocks_epn.py:228-332;
// SURVEY 8c golden constants, compared with the module's kidx_rot / ridx_rot buffers by the host before every launch)
// and the 16-row basis the tensor-core gather kernels use.
//
//   kidx_tab(k, r): weight-sharing class (0..5) kernel point k falls into after rotation by anchor r
//   ridx_tab(a, r): weight anchor slot a' that input anchor a feeds for output anchor r (a permutation of a per r)
//
// The contraction 'kpac,karcd->prd' with W_eff = W[kidx_rot, ridx_rot] (blocks_epn.py:503-506) is evaluated as
//   out[(p, r)][d] = sum_{kc, a', c} A'[(p, r)][(kc, a', c)] * W[kc][a'][c][d],
//   A'[(p, r)][(kc, ridx[a][r], c)] = sum_n ( sum_{k : kidx[k][r] == kc} w[p][n][k] ) * x[idx[p][n]][a][c].
// The 36 kernel-point subsets {k : kidx[k][r] == kc} are only 16 distinct sets:
//   rows 0-5   the single vertices k = 0..5                (class 0 of one r, class 2 of the opposite r)
//   rows 6-8   the three equatorial 4-vertex sets          (class 1 of an antipodal pair of r)
//   rows 9-14  the six 4-face half spaces                  (class 3 of r, class 4 of the opposite r)
//   row  15    the centre point k = 14                     (class 5 of every r)
// so per point ONE 16-row weight matrix W16[row][n] = sum_{k in set(row)} w[n][k] times the gathered features gives
// every A' entry as a plain copy of a product element: no adds after the tensor-core product.
#pragma once
#include <stdint.h>

namespace se3et {

constexpr int kA = 6;    // anchors
constexpr int kKP = 15;  // kernel points
constexpr int kKC = 6;   // weight-sharing classes

__host__ __device__ constexpr int kidx_tab(int k, int r) {
  constexpr int t[15][6] = {{0, 1, 1, 1, 1, 2}, {1, 0, 1, 2, 1, 1}, {1, 1, 0, 1, 2, 1}, {1, 2, 1, 0, 1, 1},
                            {1, 1, 2, 1, 0, 1}, {2, 1, 1, 1, 1, 0}, {3, 3, 3, 4, 4, 4}, {3, 4, 3, 3, 4, 4},
                            {3, 4, 4, 3, 3, 4}, {3, 3, 4, 4, 3, 4}, {4, 3, 3, 4, 4, 3}, {4, 4, 3, 3, 4, 3},
                            {4, 4, 4, 3, 3, 3}, {4, 3, 4, 4, 3, 3}, {5, 5, 5, 5, 5, 5}};
  return t[k][r];
}
__host__ __device__ constexpr int ridx_tab(int a, int r) {
  constexpr int t[6][6] = {{0, 3, 3, 3, 3, 5}, {1, 0, 4, 5, 2, 1}, {2, 2, 0, 4, 5, 4},
                           {3, 5, 2, 0, 4, 3}, {4, 4, 5, 2, 0, 2}, {5, 1, 1, 1, 1, 0}};
  return t[a][r];
}

// kernel points whose influence weights are summed for (output anchor r, class kc): 15-bit mask
__host__ __device__ constexpr uint32_t class_mask(int r, int kc) {
  uint32_t m = 0;
  for (int k = 0; k < kKP; ++k)
    if (kidx_tab(k, r) == kc) m |= 1u << k;
  return m;
}
__host__ __device__ constexpr uint32_t basis_mask(int row) {
  return row < 6 ? (1u << row) : row < 9 ? class_mask(row - 6, 1) : row < 15 ? class_mask(row - 9, 3) : (1u << 14);
}
// basis row that equals the subset of (r, kc); -1 if the tables were not closed (checked below)
__host__ __device__ constexpr int basis_row(int r, int kc) {
  for (int row = 0; row < 16; ++row)
    if (basis_mask(row) == class_mask(r, kc)) return row;
  return -1;
}
__host__ __device__ constexpr bool basis_complete() {
  for (int r = 0; r < kA; ++r)
    for (int kc = 0; kc < kKC; ++kc)
      if (basis_row(r, kc) < 0) return false;
  return true;
}
static_assert(basis_complete(), "the 36 (anchor, class) kernel-point subsets must reduce to the 16-row basis");

// t-th (r, kc) pair (t = 0..5) that copies basis row `row`, packed r * 8 + kc; -1 past the end
__host__ __device__ constexpr int basis_target(int row, int t) {
  int seen = 0;
  for (int r = 0; r < kA; ++r)
    for (int kc = 0; kc < kKC; ++kc)
      if (basis_row(r, kc) == row) {
        if (seen == t) return r * 8 + kc;
        ++seen;
      }
  return -1;
}
// rows 0..14 are copied to exactly two (r, kc) pairs, the centre row to six
__host__ __device__ constexpr bool basis_fanout_ok() {
  for (int row = 0; row < 15; ++row)
    if (basis_target(row, 1) < 0 || basis_target(row, 2) >= 0) return false;
  return basis_target(15, 5) >= 0;
}
static_assert(basis_fanout_ok(), "unexpected fan-out of the 16-row basis");

// column r of ridx_tab packed 3 bits per input anchor a: (word >> 3a) & 7 == ridx_tab(a, r)
__host__ __device__ constexpr uint32_t ridx_col_packed(int r) {
  uint32_t w = 0;
  for (int a = 0; a < kA; ++a) w |= (uint32_t)ridx_tab(a, r) << (3 * a);
  return w;
}

}  // namespace se3et
