// pcl_eigen33.cuh — float32 closed-form symmetric 3x3 eigen routines with the operation order of
// pcl/common/impl/eigen.hpp (computeRoots2, computeRoots, eigen33), shared by the normal estimation
// kernels and the RANSAC sample-distance threshold of the grouping stage.
#pragma once

#include "common.cuh"

__device__ __forceinline__ void compute_roots2(float b, float c, float roots[3]) {
  roots[0] = 0.0f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  const float sd = sqrtf(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

// pcl::computeRoots (common/impl/eigen.hpp), float instantiation
__device__ inline void compute_roots(const float m[9], float roots[3]) {
  const float c0 = m[0] * m[4] * m[8] + 2.0f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] -
                   m[8] * m[1] * m[1];
  const float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  const float c2 = m[0] + m[4] + m[8];
  if (fabsf(c0) < 1.1920928955078125e-07f) {
    compute_roots2(c2, c1, roots);
  } else {
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = sqrtf(3.0f);
    const float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    const float rho = sqrtf(-a_over_3);
    const float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
    const float cos_theta = cosf(theta);
    const float sin_theta = sinf(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    float t;
    if (roots[0] >= roots[1]) {
      t = roots[0];
      roots[0] = roots[1];
      roots[1] = t;
    }
    if (roots[1] >= roots[2]) {
      t = roots[1];
      roots[1] = roots[2];
      roots[2] = t;
      if (roots[0] >= roots[1]) {
        t = roots[0];
        roots[0] = roots[1];
        roots[1] = t;
      }
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
  }
}

__device__ __forceinline__ void cross3(const float *a, const float *b, float *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}


// pcl::eigen33(mat, evals): eigenvalues only, ascending
__device__ inline void eigen33_values(const float cov[9], float evals[3]) {
  float scale = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(cov[i]));
  if (scale <= 1.17549435e-38f) scale = 1.0f;
  float m[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) m[i] = cov[i] / scale;
  compute_roots(m, evals);
  for (int i = 0; i < 3; ++i) evals[i] *= scale;
}
