// linalg3.cuh — 3x3 double-precision helpers for the device code: symmetric eigen-decomposition
// (stands in for Eigen::SelfAdjointEigenSolver<Matrix3d> inside PCL's SHOT local reference frame)
// and the rigid Umeyama fit (pcl::umeyama(src, dst, false) as called by
// SampleConsensusModelRegistration::estimateRigidTransformationSVD).
#pragma once

#include "common.cuh"

// Cyclic Jacobi.  A: row-major symmetric.  w: ascending eigenvalues, V: eigenvectors in columns
// (row-major 3x3).
__host__ __device__ inline void eigh3_f64(const double A_in[9], double w[3], double V[9]) {
  double a[3][3] = {{A_in[0], A_in[1], A_in[2]}, {A_in[3], A_in[4], A_in[5]}, {A_in[6], A_in[7], A_in[8]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  const double d[3] = {a[0][0], a[1][1], a[2][2]};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (d[idx[j]] > d[idx[j + 1]]) {
        const int t = idx[j];
        idx[j] = idx[j + 1];
        idx[j + 1] = t;
      }
  for (int j = 0; j < 3; ++j) {
    w[j] = d[idx[j]];
    for (int k = 0; k < 3; ++k) V[k * 3 + j] = v[k][idx[j]];
  }
}

__host__ __device__ __forceinline__ void cross3d(const double *a, const double *b, double *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// Rigid (no scale) least-squares transform dst ~ R src + t of n 3-D points (row-major n x 3).
// T: row-major 4x4.  R = U diag(1, 1, det(U) det(V)) V^T from the SVD of the cross-covariance,
// obtained through the eigen-decomposition of S^T S.
__host__ __device__ inline void umeyama_from_moments(const double S[9], const double ms[3], const double md[3],
                                                     double T[16]);

__host__ __device__ inline void umeyama3(const double *src, const double *dst, int n, double T[16]) {
  double ms[3] = {0, 0, 0}, md[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) {
      ms[a] += src[i * 3 + a];
      md[a] += dst[i * 3 + a];
    }
  for (int a = 0; a < 3; ++a) {
    ms[a] /= n;
    md[a] /= n;
  }
  double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) S[r * 3 + c] += (dst[i * 3 + r] - md[r]) * (src[i * 3 + c] - ms[c]);
  for (int i = 0; i < 9; ++i) S[i] /= n;
  umeyama_from_moments(S, ms, md, T);
}

// Rigid transform from the moments of two corresponding point sets: S = cross-covariance
// (dst - md)(src - ms)^T / n (row-major 3x3), ms / md = means.  Shared by the 3-point RANSAC models and
// by ICP's TransformationEstimationSVD.
__host__ __device__ inline void umeyama_from_moments(const double S[9], const double ms[3], const double md[3],
                                                     double T[16]) {
  double StS[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += S[k * 3 + r] * S[k * 3 + c];
      StS[r * 3 + c] = s;
    }
  double w[3], V[9];
  eigh3_f64(StS, w, V);
  double v[3][3], u[3][3], sv[3];
  for (int j = 0; j < 3; ++j) {
    const int col = 2 - j;
    sv[j] = sqrt(fmax(w[col], 0.0));
    for (int k = 0; k < 3; ++k) v[j][k] = V[k * 3 + col];
  }
  for (int i = 0; i < 16; ++i) T[i] = 0;
  T[15] = 1;
  if (sv[0] <= 0.0) {
    T[0] = T[5] = T[10] = 1;
    for (int a = 0; a < 3; ++a) T[a * 4 + 3] = md[a] - ms[a];
    return;
  }
  const double tol = 1e-12 * fmax(sv[0], 1e-300);
  for (int r = 0; r < 3; ++r) u[0][r] = S[r * 3 + 0] * v[0][0] + S[r * 3 + 1] * v[0][1] + S[r * 3 + 2] * v[0][2];
  const double n0 = sqrt(u[0][0] * u[0][0] + u[0][1] * u[0][1] + u[0][2] * u[0][2]);
  for (int k = 0; k < 3; ++k) u[0][k] /= n0;
  if (sv[1] > tol) {
    for (int r = 0; r < 3; ++r) u[1][r] = S[r * 3 + 0] * v[1][0] + S[r * 3 + 1] * v[1][1] + S[r * 3 + 2] * v[1][2];
    const double d = u[1][0] * u[0][0] + u[1][1] * u[0][1] + u[1][2] * u[0][2];
    for (int k = 0; k < 3; ++k) u[1][k] -= d * u[0][k];
    const double n1 = sqrt(u[1][0] * u[1][0] + u[1][1] * u[1][1] + u[1][2] * u[1][2]);
    for (int k = 0; k < 3; ++k) u[1][k] /= n1;
  } else {
    int m = 0;
    if (fabs(u[0][1]) < fabs(u[0][m])) m = 1;
    if (fabs(u[0][2]) < fabs(u[0][m])) m = 2;
    double e[3] = {0, 0, 0};
    e[m] = 1;
    cross3d(u[0], e, u[1]);
    const double n1 = sqrt(u[1][0] * u[1][0] + u[1][1] * u[1][1] + u[1][2] * u[1][2]);
    for (int k = 0; k < 3; ++k) u[1][k] /= n1;
  }
  cross3d(u[0], u[1], u[2]);
  double c12[3];
  cross3d(v[0], v[1], c12);
  const double detV = c12[0] * v[2][0] + c12[1] * v[2][1] + c12[2] * v[2][2];
  const double dsign = detV >= 0 ? 1.0 : -1.0;
  for (int r = 0; r < 3; ++r) {
    double Rr[3];
    for (int c = 0; c < 3; ++c) Rr[c] = u[0][r] * v[0][c] + u[1][r] * v[1][c] + dsign * u[2][r] * v[2][c];
    for (int c = 0; c < 3; ++c) T[r * 4 + c] = Rr[c];
    T[r * 4 + 3] = md[r] - (Rr[0] * ms[0] + Rr[1] * ms[1] + Rr[2] * ms[2]);
  }
}
