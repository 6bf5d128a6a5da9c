// util.hpp -- small dense helpers and the geometric predicates that classify elements.
//
// Reference: include/util.hpp:58-163.  The predicates decide elem_type from floating-point geometry
// at element centroids (src/micropp.cpp:339-544), and elem_type must be bit-identical to the
// reference's, so every expression below keeps the reference's operation order (sequential sums
// starting from 0, norm = sqrt of that sum, the cylinder distance through sqrt(1 - cos^2) which is
// NaN on the axis and therefore "outside").  Host-only: this is set-up code, not the hot path.
#pragma once

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

constexpr int mypow(int v, int e) { return (e == 0) ? 1 : v * mypow(v, e - 1); }

inline uint64_t devest(const std::vector<uint64_t> &in, const uint64_t mean) {
  uint64_t acc = 0;
  for (const auto &x : in) {
    const uint64_t d = (x - mean);
    acc += d * d;
  }
  return sqrt(acc / in.size());
}

inline void print_vec(const double *vec, int n, const char file_name[]) {
  FILE *f = fopen(file_name, "w");
  for (int i = 0; i < n; ++i) fprintf(f, "[%lf]\n", vec[i]);
  fclose(f);
}

// y = m x for a dense n x n matrix (row sums accumulate left to right from 0).
template <typename T, int n>
inline void mvp(const T m[n][n], const T x[n], T *y) {
  for (int r = 0; r < n; ++r) {
    T acc = 0.0;
    for (int c = 0; c < n; ++c) acc += m[r][c] * x[c];
    y[r] = acc;
  }
}

template <typename T, int n>
inline double norm(const T v[n]) {
  T acc = 0;
  for (int i = 0; i < n; ++i) acc += v[i] * v[i];
  return sqrt((double)acc);
}

template <typename T, int n>
inline T dot_prod(const T a[n], const T b[n]) {
  T acc = 0;
  for (int i = 0; i < n; ++i) acc += a[i] * b[i];
  return acc;
}

// strictly inside the ball
inline bool point_inside_sphere(const double center[3], const double radius, const double point[3]) {
  const double d[3] = {point[0] - center[0], point[1] - center[1], point[2] - center[2]};
  return norm<double, 3>(d) < radius;
}

// inside or on an infinite cylinder of axis `dir` through `center`
inline bool point_inside_cilinder_inf(const double dir[3], const double center[3], const double radius,
                                      const double point[3]) {
  const double d[3] = {point[0] - center[0], point[1] - center[1], point[2] - center[2]};
  const double along = dot_prod<double, 3>(dir, d);
  const double len_dir = norm<double, 3>(dir);
  const double len_d = norm<double, 3>(d);
  const double c = along / (len_dir * len_d);
  const double s = sqrt(1 - c * c);
  const double dist = len_d * s;
  return dist <= radius;
}

// cofactor inverse; returns the determinant
inline double invert_3x3(const double m[3][3], double inv[3][3]) {
  const double det = m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) -
                     m[0][1] * (m[1][0] * m[2][2] - m[2][0] * m[1][2]) +
                     m[0][2] * (m[1][0] * m[2][1] - m[2][0] * m[1][1]);

  inv[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) / det;
  inv[0][1] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) / det;
  inv[0][2] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) / det;
  inv[1][0] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) / det;
  inv[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) / det;
  inv[1][2] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) / det;
  inv[2][0] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) / det;
  inv[2][1] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) / det;
  inv[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) / det;
  return det;
}
