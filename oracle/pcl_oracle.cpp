/*
 * pcl_oracle.cpp — CPU parity oracle (TEST INFRASTRUCTURE, see pcl_oracle.h).
 *
 * PARITY UNPINNED: restated from the published PCL 1.8.x / FLANN 1.8 algorithms (SURVEY.md
 * Appendix A); the reference repository carries no golden vectors for this path.
 *
 * Build: g++ -O2 -fopenmp -ffp-contract=off  (no FMA contraction: PCL's float32 sums are plain
 * mul/add sequences and the CUDA path reproduces them with __fmul_rn/__fadd_rn).
 *
 * OpenMP is applied exactly where PCL's *OMP classes parallelise (per point / per keypoint loops of
 * NormalEstimationOMP, SHOTLocalReferenceFrameEstimationOMP, SHOTEstimationOMP, FPFHEstimationOMP).
 * The matching loop (user code, SHOT.cpp:409-423) and GeometricConsistencyGrouping are serial.
 */
#include "pcl_oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <random>
#include <set>
#include <unordered_map>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const float kNaNf = std::numeric_limits<float>::quiet_NaN();

/* Sensitivity variants (oracle/sensitivity.py, tests): alternatives for the details of upstream PCL that this
 * restatement could not pin (SURVEY.md Appendix A "[?]" items), plus +-epsilon perturbations of the discrete
 * decisions, used to tell which outputs are stable.  0 = the restatement as documented. */
unsigned g_variant = 0;
enum {
  V_DOT4 = 1,         /* Eigen 4-lane dot products: (x x' + z z') + (y y' + w w') instead of the sequential sum */
  V_UMEYAMA_F32 = 2,  /* rigid fit from float32 moments (PCL's Matrix4f path) instead of float64 */
  V_FPFH_SKIP = 4,    /* degenerate FPFH pairs do not vote (Appendix A) instead of voting with f = 0 */
  V_FPFH_BIN_UP = 8,  /* FPFH f1 (atan2f) bin coordinate + 1e-6 / - 1e-6 before the floor */
  V_FPFH_BIN_DOWN = 16,
  V_ROOT_UP = 32,     /* eigen33: theta of the closed-form roots * (1 +- 2^-21) (libm atan2f / cosf / sinf ulps) */
  V_ROOT_DOWN = 64,
  V_SHOT_BIN_UP = 128,  /* SHOT: cosine-bin coordinate + 1e-6 / - 1e-6 before floor(x + 0.5) */
  V_SHOT_BIN_DOWN = 256,
  V_BOARD_ANGLE_UP = 512, /* BOARD: direction angle of a support point * (1 +- 2^-21) before the sector index (acosf ulps) */
  V_BOARD_ANGLE_DOWN = 1024
};

/* float dot product of two 3-vectors the way the call site evaluates it */
inline float vdot3f(const float *a, const float *b) {
  if (g_variant & V_DOT4) return (a[0] * b[0] + a[2] * b[2]) + (a[1] * b[1] + 0.0f);
  float s = a[0] * b[0];
  s += a[1] * b[1];
  s += a[2] * b[2];
  return s;
}
inline double dot3d(const double *a, const double *b) {
  if (g_variant & V_DOT4) return (a[0] * b[0] + a[2] * b[2]) + (a[1] * b[1] + 0.0);
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

inline bool finite3(const float *p) {
  return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]);
}

/* FLANN L2_Simple<float> over 3 dims (flann/algorithms/dist.h): diff = a[i]-b[i]; result += diff*diff. */
inline float sqdist3(const float *a, const float *b) {
  float r = 0.0f;
  float d0 = a[0] - b[0];
  r += d0 * d0;
  float d1 = a[1] - b[1];
  r += d1 * d1;
  float d2 = a[2] - b[2];
  r += d2 * d2;
  return r;
}

struct DistIdx {
  float d2;
  int idx;
  /* flann/util/result_set.h DistanceIndex::operator< : (dist, index) lexicographic. */
  bool operator<(const DistIdx &o) const { return (d2 < o.d2) || (d2 == o.d2 && idx < o.idx); }
};

/* ------------------------------------------------------------------------------------------------
 * Exact neighbour search.  PCL uses a FLANN kd-tree (pcl/kdtree/impl/kdtree_flann.hpp); the search
 * is exact (checks=-1, eps=0), so any exact structure yields the same (d2, index)-sorted answers.
 * A uniform grid keeps the oracle fast enough for the 1 M-point configurations.
 * Non-finite surface rows are dropped exactly like KdTreeFLANN::convertCloudToArray does (the
 * returned indices still refer to the original cloud, via index_mapping_).
 * ---------------------------------------------------------------------------------------------- */
struct CpuGrid {
  int n = 0;
  double lo[3] = {0, 0, 0};
  double h = 1.0;
  int dim[3] = {1, 1, 1};
  std::vector<int> cell_start;   /* ncell + 1 */
  std::vector<int> order;        /* original indices, cell-major, ascending index inside a cell */
  std::vector<float> pts;        /* 3 floats per sorted point */
  int n_valid = 0;

  inline int cell_coord(double v, int a) const {
    int c = (int)std::floor((v - lo[a]) / h);
    if (c < 0) c = 0;
    if (c >= dim[a]) c = dim[a] - 1;
    return c;
  }

  void build(const float *surf, int n_, int stride, double cell) {
    n = n_;
    double hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    lo[0] = lo[1] = lo[2] = DBL_MAX;
    n_valid = 0;
    for (int i = 0; i < n; ++i) {
      const float *p = surf + (size_t)i * stride;
      if (!finite3(p)) continue;
      ++n_valid;
      for (int a = 0; a < 3; ++a) {
        lo[a] = std::min(lo[a], (double)p[a]);
        hi[a] = std::max(hi[a], (double)p[a]);
      }
    }
    if (n_valid == 0) {
      lo[0] = lo[1] = lo[2] = 0;
      hi[0] = hi[1] = hi[2] = 0;
    }
    h = cell;
    /* keep the dense cell array bounded: at most 2^26 cells, 1024 per axis */
    for (;;) {
      double cells = 1;
      bool ok = true;
      for (int a = 0; a < 3; ++a) {
        double d = std::floor((hi[a] - lo[a]) / h) + 1;
        if (d > 1024) ok = false;
        cells *= d;
      }
      if (ok && cells <= (double)(1 << 26)) break;
      h *= 1.25;
    }
    for (int a = 0; a < 3; ++a) dim[a] = (int)std::floor((hi[a] - lo[a]) / h) + 1;
    size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
    cell_start.assign(ncell + 1, 0);
    std::vector<int> cell_of(n, -1);
    for (int i = 0; i < n; ++i) {
      const float *p = surf + (size_t)i * stride;
      if (!finite3(p)) continue;
      int c = cell_coord(p[0], 0) + dim[0] * (cell_coord(p[1], 1) + dim[1] * cell_coord(p[2], 2));
      cell_of[i] = c;
      cell_start[c + 1]++;
    }
    for (size_t c = 0; c < ncell; ++c) cell_start[c + 1] += cell_start[c];
    order.resize(n_valid);
    pts.resize((size_t)n_valid * 3);
    std::vector<int> cursor(cell_start.begin(), cell_start.end() - 1);
    for (int i = 0; i < n; ++i) {
      int c = cell_of[i];
      if (c < 0) continue;
      int s = cursor[c]++;
      order[s] = i;
      const float *p = surf + (size_t)i * stride;
      pts[(size_t)s * 3 + 0] = p[0];
      pts[(size_t)s * 3 + 1] = p[1];
      pts[(size_t)s * 3 + 2] = p[2];
    }
  }

  /* KdTreeFLANN::radiusSearch: r2 = (float)(radius*radius); accept d2 < r2 (RadiusResultSet::addPoint);
   * results sorted (sorted_ = true). */
  void radius(const float *q, double radius_, std::vector<DistIdx> &out) const {
    out.clear();
    if (n_valid == 0 || !finite3(q)) return;
    const float r2 = (float)(radius_ * radius_);
    int c0[3], c1[3];
    for (int a = 0; a < 3; ++a) {
      c0[a] = cell_coord((double)q[a] - radius_ - 1e-9, a);
      c1[a] = cell_coord((double)q[a] + radius_ + 1e-9, a);
    }
    for (int z = c0[2]; z <= c1[2]; ++z)
      for (int y = c0[1]; y <= c1[1]; ++y) {
        size_t row = (size_t)dim[0] * (y + (size_t)dim[1] * z);
        int s = cell_start[row + c0[0]], e = cell_start[row + c1[0] + 1];
        for (int j = s; j < e; ++j) {
          float d2 = sqdist3(q, &pts[(size_t)j * 3]);
          if (d2 < r2) out.push_back({d2, order[j]});
        }
      }
    std::sort(out.begin(), out.end());
  }

  /* KdTreeFLANN::nearestKSearch: exact k nearest, k clamped to the number of points; order (d2, idx). */
  void knn(const float *q, int k, std::vector<DistIdx> &heap) const {
    heap.clear();
    if (n_valid == 0 || !finite3(q)) return;
    if (k > n_valid) k = n_valid;
    int c[3];
    for (int a = 0; a < 3; ++a) c[a] = cell_coord(q[a], a);
    int maxR = std::max(dim[0], std::max(dim[1], dim[2]));
    for (int R = 0; R <= maxR; ++R) {
      /* scan the shell at Chebyshev distance R */
      int z0 = std::max(c[2] - R, 0), z1 = std::min(c[2] + R, dim[2] - 1);
      int y0 = std::max(c[1] - R, 0), y1 = std::min(c[1] + R, dim[1] - 1);
      int x0 = std::max(c[0] - R, 0), x1 = std::min(c[0] + R, dim[0] - 1);
      for (int z = z0; z <= z1; ++z)
        for (int y = y0; y <= y1; ++y) {
          bool face = (std::abs(z - c[2]) == R) || (std::abs(y - c[1]) == R);
          size_t row = (size_t)dim[0] * (y + (size_t)dim[1] * z);
          auto scan = [&](int xa, int xb) {
            int s = cell_start[row + xa], e = cell_start[row + xb + 1];
            for (int j = s; j < e; ++j) {
              DistIdx di{sqdist3(q, &pts[(size_t)j * 3]), order[j]};
              if ((int)heap.size() < k) {
                heap.push_back(di);
                std::push_heap(heap.begin(), heap.end());
              } else if (di < heap.front()) {
                std::pop_heap(heap.begin(), heap.end());
                heap.back() = di;
                std::push_heap(heap.begin(), heap.end());
              }
            }
          };
          if (face) {
            scan(x0, x1);
          } else {
            if (c[0] - R >= 0) scan(c[0] - R, c[0] - R);
            if (c[0] + R <= dim[0] - 1 && R > 0) scan(c[0] + R, c[0] + R);
          }
        }
      if ((int)heap.size() == k) {
        /* everything outside the scanned block is at least `cert` away */
        double cert = DBL_MAX;
        for (int a = 0; a < 3; ++a) {
          if (c[a] - R > 0) cert = std::min(cert, (double)q[a] - (lo[a] + (double)(c[a] - R) * h));
          if (c[a] + R < dim[a] - 1) cert = std::min(cert, (lo[a] + (double)(c[a] + R + 1) * h) - (double)q[a]);
        }
        if (cert == DBL_MAX) break; /* whole grid scanned */
        if (cert > 0) {
          double c2 = cert * cert * (1.0 - 1e-5);
          if ((double)heap.front().d2 < c2) break;
        }
      }
    }
    std::sort_heap(heap.begin(), heap.end());
  }
};

double auto_cell_for_knn(const float *surf, int n, int stride, int k) {
  /* choose a cell so that non-empty cells hold about k/2 points (surface-like data) */
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  int nv = 0;
  for (int i = 0; i < n; ++i) {
    const float *p = surf + (size_t)i * stride;
    if (!finite3(p)) continue;
    ++nv;
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], (double)p[a]);
      hi[a] = std::max(hi[a], (double)p[a]);
    }
  }
  if (nv == 0) return 1.0;
  double ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
  std::sort(ext, ext + 3);
  double area = std::max(ext[2] * ext[1], 1e-12);
  double target = std::max(2.0, 0.5 * k);
  double h = std::sqrt(area * target / nv);
  CpuGrid g;
  for (int it = 0; it < 3; ++it) {
    g.build(surf, n, stride, h);
    size_t ne = 0;
    for (size_t c = 0; c + 1 < g.cell_start.size(); ++c) ne += (g.cell_start[c + 1] > g.cell_start[c]);
    double occ = (double)nv / std::max<size_t>(ne, 1);
    double ratio = target / occ;
    if (ratio > 0.7 && ratio < 1.4) break;
    h = g.h * std::sqrt(ratio);
  }
  return h;
}

double cell_for_radius(double radius) { return radius > 0 ? radius : 1.0; }

/* ------------------------------------------------------------------------------------------------
 * pcl/common/impl/eigen.hpp : computeRoots2 / computeRoots / eigen33 (float instantiation).
 * ---------------------------------------------------------------------------------------------- */
inline void compute_roots2(float b, float c, float roots[3]) {
  roots[0] = 0.0f;
  float d = (float)(b * b - 4.0 * c);
  if (d < 0.0) d = 0.0;
  float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

inline void compute_roots(const float m[9], float roots[3]) {
  /* m is row-major symmetric 3x3 */
  float c0 = m[0] * m[4] * m[8] + 2.0f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] -
             m[8] * m[1] * m[1];
  float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  float c2 = m[0] + m[4] + m[8];
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {
    compute_roots2(c2, c1, roots);
  } else {
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = std::sqrt(-a_over_3);
    float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
    if (g_variant & V_ROOT_UP) theta *= 1.0f + 4.76837158e-7f;
    if (g_variant & V_ROOT_DOWN) theta *= 1.0f - 4.76837158e-7f;
    float cos_theta = std::cos(theta);
    float sin_theta = std::sin(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
      std::swap(roots[1], roots[2]);
      if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
  }
}

inline void cross3f(const float *a, const float *b, float *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

/* pcl::eigen33(mat, eigenvalue, eigenvector): smallest eigenvalue and its eigenvector. */
void eigen33_smallest(const float cov[9], float &eigenvalue, float evec[3]) {
  float scale = 0.0f;
  for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(cov[i]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float m[9];
  for (int i = 0; i < 9; ++i) m[i] = cov[i] / scale;
  float roots[3];
  compute_roots(m, roots);
  eigenvalue = roots[0] * scale;
  m[0] -= roots[0];
  m[4] -= roots[0];
  m[8] -= roots[0];
  float v1[3], v2[3], v3[3];
  cross3f(&m[0], &m[3], v1);
  cross3f(&m[0], &m[6], v2);
  cross3f(&m[3], &m[6], v3);
  float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float *v;
  float l;
  if (l1 >= l2 && l1 >= l3) {
    v = v1;
    l = l1;
  } else if (l2 >= l1 && l2 >= l3) {
    v = v2;
    l = l2;
  } else {
    v = v3;
    l = l3;
  }
  float s = std::sqrt(l);
  evec[0] = v[0] / s;
  evec[1] = v[1] / s;
  evec[2] = v[2] / s;
}

/* pcl::eigen33(mat, evals): eigenvalues only (used by computeSampleDistanceThreshold). */
void eigen33_values(const float cov[9], float evals[3]) {
  float scale = 0.0f;
  for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(cov[i]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float m[9];
  for (int i = 0; i < 9; ++i) m[i] = cov[i] / scale;
  compute_roots(m, evals);
  for (int i = 0; i < 3; ++i) evals[i] *= scale;
}

/* pcl/common/impl/centroid.hpp computeMeanAndCovarianceMatrix (PCL 1.8: single pass, 9 float
 * accumulators, no mean shift), over points given in list order. */
template <class GetPoint>
inline void mean_and_cov(int count, GetPoint get, float cov[9], float centroid[3]) {
  float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < count; ++i) {
    const float *p = get(i);
    accu[0] += p[0] * p[0];
    accu[1] += p[0] * p[1];
    accu[2] += p[0] * p[2];
    accu[3] += p[1] * p[1];
    accu[4] += p[1] * p[2];
    accu[5] += p[2] * p[2];
    accu[6] += p[0];
    accu[7] += p[1];
    accu[8] += p[2];
  }
  float fn = (float)count;
  for (int i = 0; i < 9; ++i) accu[i] /= fn;
  centroid[0] = accu[6];
  centroid[1] = accu[7];
  centroid[2] = accu[8];
  cov[0] = accu[0] - accu[6] * accu[6];
  cov[1] = accu[1] - accu[6] * accu[7];
  cov[2] = accu[2] - accu[6] * accu[8];
  cov[4] = accu[3] - accu[7] * accu[7];
  cov[5] = accu[4] - accu[7] * accu[8];
  cov[8] = accu[5] - accu[8] * accu[8];
  cov[3] = cov[1];
  cov[6] = cov[2];
  cov[7] = cov[5];
}

/* ------------------------------------------------------------------------------------------------
 * Symmetric 3x3 eigen-solver in double (stand-in for Eigen::SelfAdjointEigenSolver<Matrix3d>, which
 * PCL's SHOT LRF uses; any accurate solver gives the same eigenvectors up to sign, and the sign is
 * fixed afterwards by the disambiguation votes).  Cyclic Jacobi; ascending eigenvalues; eigenvectors
 * in the columns of V (row-major 3x3).
 * ---------------------------------------------------------------------------------------------- */
void eigh3_f64(const double A_in[9], double w[3], double V[9]) {
  double a[3][3] = {{A_in[0], A_in[1], A_in[2]}, {A_in[3], A_in[4], A_in[5]}, {A_in[6], A_in[7], A_in[8]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double apq = a[p][q];
        if (apq == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { /* A <- A * J */
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) { /* A <- J^T * A */
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  double d[3] = {a[0][0], a[1][1], a[2][2]};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (d[idx[j]] > d[idx[j + 1]]) std::swap(idx[j], idx[j + 1]);
  for (int j = 0; j < 3; ++j) {
    w[j] = d[idx[j]];
    for (int k = 0; k < 3; ++k) V[k * 3 + j] = v[k][idx[j]];
  }
}

/* pcl::umeyama(src, dst, with_scaling=false) on double 3xN (what
 * SampleConsensusModelRegistration::estimateRigidTransformationSVD calls).  SVD of the 3x3
 * cross-covariance via the eigen-decomposition of S^T S; rotation R = U diag(1,1,det(U)det(V)) V^T. */
void umeyama3(const double *src, const double *dst, int n, double T[16]) {
  double ms[3] = {0, 0, 0}, md[3] = {0, 0, 0};
  double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; /* sigma = dst_demean * src_demean^T / n */
  if (g_variant & V_UMEYAMA_F32) {
    /* float32 means and cross-covariance (what a Matrix<float, 3, Dynamic> umeyama accumulates); the
     * decomposition below stays float64, standing in for a backward-stable float SVD */
    float fs[3] = {0, 0, 0}, fd[3] = {0, 0, 0}, FS[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) {
        fs[a] += (float)src[i * 3 + a];
        fd[a] += (float)dst[i * 3 + a];
      }
    for (int a = 0; a < 3; ++a) {
      fs[a] /= (float)n;
      fd[a] /= (float)n;
    }
    for (int i = 0; i < n; ++i)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) FS[r * 3 + c] += ((float)dst[i * 3 + r] - fd[r]) * ((float)src[i * 3 + c] - fs[c]);
    for (int i = 0; i < 9; ++i) S[i] = FS[i] / (float)n;
    for (int a = 0; a < 3; ++a) {
      ms[a] = fs[a];
      md[a] = fd[a];
    }
  } else {
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) {
        ms[a] += src[i * 3 + a];
        md[a] += dst[i * 3 + a];
      }
    for (int a = 0; a < 3; ++a) {
      ms[a] /= n;
      md[a] /= n;
    }
    for (int i = 0; i < n; ++i)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) S[r * 3 + c] += (dst[i * 3 + r] - md[r]) * (src[i * 3 + c] - ms[c]);
    for (int i = 0; i < 9; ++i) S[i] /= n;
  }
  double StS[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += S[k * 3 + r] * S[k * 3 + c];
      StS[r * 3 + c] = s;
    }
  double w[3], V[9];
  eigh3_f64(StS, w, V);
  /* descending singular values: column order 2,1,0 */
  double v[3][3], u[3][3], sv[3];
  for (int j = 0; j < 3; ++j) {
    int src_col = 2 - j;
    sv[j] = std::sqrt(std::max(w[src_col], 0.0));
    for (int k = 0; k < 3; ++k) v[j][k] = V[k * 3 + src_col];
  }
  auto matvec = [&](const double *x, double *y) {
    for (int r = 0; r < 3; ++r) y[r] = S[r * 3 + 0] * x[0] + S[r * 3 + 1] * x[1] + S[r * 3 + 2] * x[2];
  };
  auto norm3 = [](const double *x) { return std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); };
  auto cross = [](const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
  };
  const double tol = 1e-12 * std::max(sv[0], 1e-300);
  if (sv[0] <= 0.0) { /* all points coincide: identity rotation */
    for (int i = 0; i < 16; ++i) T[i] = 0;
    T[0] = T[5] = T[10] = T[15] = 1;
    for (int a = 0; a < 3; ++a) T[a * 4 + 3] = md[a] - ms[a];
    return;
  }
  matvec(v[0], u[0]);
  double n0 = norm3(u[0]);
  for (int k = 0; k < 3; ++k) u[0][k] /= n0;
  if (sv[1] > tol) {
    matvec(v[1], u[1]);
    /* re-orthogonalise against u0 for numerical safety */
    double d = u[1][0] * u[0][0] + u[1][1] * u[0][1] + u[1][2] * u[0][2];
    for (int k = 0; k < 3; ++k) u[1][k] -= d * u[0][k];
    double n1 = norm3(u[1]);
    for (int k = 0; k < 3; ++k) u[1][k] /= n1;
  } else { /* rank 1: any unit vector orthogonal to u0 (deterministic choice) */
    int m = 0;
    if (std::fabs(u[0][1]) < std::fabs(u[0][m])) m = 1;
    if (std::fabs(u[0][2]) < std::fabs(u[0][m])) m = 2;
    double e[3] = {0, 0, 0};
    e[m] = 1;
    cross(u[0], e, u[1]);
    double n1 = norm3(u[1]);
    for (int k = 0; k < 3; ++k) u[1][k] /= n1;
  }
  cross(u[0], u[1], u[2]); /* det(U) = +1 by construction */
  double c12[3];
  cross(v[0], v[1], c12);
  double detV = c12[0] * v[2][0] + c12[1] * v[2][1] + c12[2] * v[2][2];
  double dsign = detV >= 0 ? 1.0 : -1.0;
  double R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = u[0][r] * v[0][c] + u[1][r] * v[1][c] + dsign * u[2][r] * v[2][c];
  for (int i = 0; i < 16; ++i) T[i] = 0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[r * 4 + c] = R[r * 3 + c];
    T[r * 4 + 3] = md[r] - (R[r * 3 + 0] * ms[0] + R[r * 3 + 1] * ms[1] + R[r * 3 + 2] * ms[2]);
  }
  if (g_variant & V_UMEYAMA_F32)
    for (int r = 0; r < 3; ++r) {
      const float fR[3] = {(float)R[r * 3 + 0], (float)R[r * 3 + 1], (float)R[r * 3 + 2]};
      T[r * 4 + 3] = (float)md[r] - ((fR[0] * (float)ms[0] + fR[1] * (float)ms[1]) + fR[2] * (float)ms[2]);
    }
  T[15] = 1;
}

/* ------------------------------------------------------------------------------------------------
 * SHOT local reference frame — pcl/features/impl/shot_lrf.hpp getLocalRF.
 * nb: sorted neighbours of the keypoint.  Returns false (NaN frame) when < 5 valid neighbours.
 * ---------------------------------------------------------------------------------------------- */
bool shot_local_rf(const float *surf, int sstride, const float *central, const std::vector<DistIdx> &nb,
                   double radius, float rf[9]) {
  std::vector<double> vij;
  vij.reserve(nb.size() * 3);
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double sum = 0.0;
  int valid = 0;
  for (size_t i = 0; i < nb.size(); ++i) {
    const float *pt = surf + (size_t)nb[i].idx * sstride;
    if (pt[0] == central[0] && pt[1] == central[1] && pt[2] == central[2]) continue;
    double v[3] = {(double)(pt[0] - central[0]), (double)(pt[1] - central[1]), (double)(pt[2] - central[2])};
    double distance = radius - std::sqrt((double)nb[i].d2);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) cov[r * 3 + c] += distance * (v[r] * v[c]);
    sum += distance;
    vij.push_back(v[0]);
    vij.push_back(v[1]);
    vij.push_back(v[2]);
    ++valid;
  }
  if (valid < 5) {
    for (int i = 0; i < 9; ++i) rf[i] = kNaNf;
    return false;
  }
  for (int i = 0; i < 9; ++i) cov[i] /= sum;
  double w[3], V[9];
  eigh3_f64(cov, w, V);
  if (!std::isfinite(w[0]) || !std::isfinite(w[1]) || !std::isfinite(w[2])) {
    for (int i = 0; i < 9; ++i) rf[i] = kNaNf;
    return false;
  }
  double v1[3] = {V[0 * 3 + 2], V[1 * 3 + 2], V[2 * 3 + 2]}; /* largest eigenvalue -> x */
  double v3[3] = {V[0 * 3 + 0], V[1 * 3 + 0], V[2 * 3 + 0]}; /* smallest eigenvalue -> z */
  int plusNormal = 0, plusTangent = 0;
  for (int ne = 0; ne < valid; ++ne) {
    const double *v = &vij[(size_t)ne * 3];
    double dp = dot3d(v, v1);
    if (dp >= 0) plusTangent++;
    dp = dot3d(v, v3);
    if (dp >= 0) plusNormal++;
  }
  auto disambiguate = [&](int plus, double *axis) {
    plus = 2 * plus - valid;
    if (plus == 0) {
      const int points = 5;
      int medianIndex = valid / 2;
      for (int i = -points / 2; i <= points / 2; i++) {
        const double *v = &vij[(size_t)(medianIndex - i) * 3];
        if (dot3d(v, axis) > 0) plus++;
      }
      if (plus < points / 2 + 1)
        for (int k = 0; k < 3; ++k) axis[k] *= -1;
    } else if (plus < 0) {
      for (int k = 0; k < 3; ++k) axis[k] *= -1;
    }
  };
  disambiguate(plusTangent, v1);
  disambiguate(plusNormal, v3);
  float x[3] = {(float)v1[0], (float)v1[1], (float)v1[2]};
  float z[3] = {(float)v3[0], (float)v3[1], (float)v3[2]};
  float y[3];
  cross3f(z, x, y);
  for (int k = 0; k < 3; ++k) {
    rf[k] = x[k];
    rf[3 + k] = y[k];
    rf[6 + k] = z[k];
  }
  return true;
}

/* ------------------------------------------------------------------------------------------------
 * SHOT352 — pcl/features/impl/shot.hpp: createBinDistanceShape, interpolateSingleChannel,
 * normalizeHistogram, computePointSHOT.
 * ---------------------------------------------------------------------------------------------- */
const double PST_PI = 3.1415926535897932384626433832795;
const double PST_RAD_45 = 0.78539816339744830961566084581988;
const double PST_RAD_90 = 1.5707963267948966192313216916398;
const double PST_RAD_135 = 2.3561944901923449288469825374596;
const double PST_RAD_PI_7_8 = 2.7488935718910690836548129603691;

void point_shot352(const float *surf, const float *normals, int sstride, const float *central,
                   const float rf[9], const std::vector<DistIdx> &nb, double radius, float *shot) {
  const int nr_bins = 10, stride_b = 11, max_sectors = 32, desc_len = 352;
  if (nb.size() < 5) {
    for (int d = 0; d < desc_len; ++d) shot[d] = kNaNf;
    return;
  }
  const double radius3_4 = (radius * 3) / 4, radius1_4 = radius / 4, radius1_2 = radius / 2;
  for (int d = 0; d < desc_len; ++d) shot[d] = 0.0f;
  const float *fx = rf, *fy = rf + 3, *fz = rf + 6;
  for (size_t i = 0; i < nb.size(); ++i) {
    const float *nrm = normals + (size_t)nb[i].idx * 4;
    if (!std::isfinite(nrm[0]) || !std::isfinite(nrm[1]) || !std::isfinite(nrm[2])) continue;
    /* createBinDistanceShape: float dot, widened */
    float dotf = vdot3f(nrm, fz);
    double cosineDesc = dotf;
    if (cosineDesc > 1.0) cosineDesc = 1.0;
    if (cosineDesc < -1.0) cosineDesc = -1.0;
    double binDistance = ((1.0 + cosineDesc) * nr_bins) / 2;

    const float *pt = surf + (size_t)nb[i].idx * sstride;
    float delta[3] = {pt[0] - central[0], pt[1] - central[1], pt[2] - central[2]};
    double distance = std::sqrt((double)nb[i].d2);
    if (std::fabs(distance - 0.0) < 1E-15) continue;
    auto dotd = [&](const float *f) { return (double)vdot3f(delta, f); };
    double xInFeatRef = dotd(fx), yInFeatRef = dotd(fy), zInFeatRef = dotd(fz);
    if (std::fabs(yInFeatRef) < 1E-30) yInFeatRef = 0;
    if (std::fabs(xInFeatRef) < 1E-30) xInFeatRef = 0;
    if (std::fabs(zInFeatRef) < 1E-30) zInFeatRef = 0;

    unsigned char bit4 = ((yInFeatRef > 0) || ((yInFeatRef == 0.0) && (xInFeatRef < 0))) ? 1 : 0;
    unsigned char bit3 =
        (unsigned char)(((xInFeatRef > 0) || ((xInFeatRef == 0.0) && (yInFeatRef > 0))) ? !bit4 : bit4);
    int desc_index = (bit4 << 3) + (bit3 << 2);
    desc_index = desc_index << 1;
    if ((xInFeatRef * yInFeatRef > 0) || (xInFeatRef == 0.0))
      desc_index += (std::fabs(xInFeatRef) >= std::fabs(yInFeatRef)) ? 0 : 4;
    else
      desc_index += (std::fabs(xInFeatRef) > std::fabs(yInFeatRef)) ? 4 : 0;
    desc_index += zInFeatRef > 0 ? 1 : 0;
    desc_index += (distance > radius1_2) ? 2 : 0;

    if (g_variant & V_SHOT_BIN_UP) binDistance += 1e-6;
    if (g_variant & V_SHOT_BIN_DOWN) binDistance -= 1e-6;
    int step_index = (int)std::floor(binDistance + 0.5);
    int volume_index = desc_index * stride_b;

    binDistance -= step_index;
    double intWeight = (1 - std::fabs(binDistance));
    if (binDistance > 0)
      shot[volume_index + ((step_index + 1) % nr_bins)] += (float)binDistance;
    else
      shot[volume_index + ((step_index - 1 + nr_bins) % nr_bins)] += -(float)binDistance;

    if (distance > radius1_2) {
      double radiusDistance = (distance - radius3_4) / radius1_2;
      if (distance > radius3_4)
        intWeight += 1 - radiusDistance;
      else {
        intWeight += 1 + radiusDistance;
        shot[(desc_index - 2) * stride_b + step_index] -= (float)radiusDistance;
      }
    } else {
      double radiusDistance = (distance - radius1_4) / radius1_2;
      if (distance < radius1_4)
        intWeight += 1 + radiusDistance;
      else {
        intWeight += 1 - radiusDistance;
        shot[(desc_index + 2) * stride_b + step_index] += (float)radiusDistance;
      }
    }

    double inclinationCos = zInFeatRef / distance;
    if (inclinationCos < -1.0) inclinationCos = -1.0;
    if (inclinationCos > 1.0) inclinationCos = 1.0;
    double inclination = std::acos(inclinationCos);
    if (inclination > PST_RAD_90 || (std::fabs(inclination - PST_RAD_90) < 1e-30 && zInFeatRef <= 0)) {
      double inclinationDistance = (inclination - PST_RAD_135) / PST_RAD_90;
      if (inclination > PST_RAD_135)
        intWeight += 1 - inclinationDistance;
      else {
        intWeight += 1 + inclinationDistance;
        shot[(desc_index + 1) * stride_b + step_index] -= (float)inclinationDistance;
      }
    } else {
      double inclinationDistance = (inclination - PST_RAD_45) / PST_RAD_90;
      if (inclination < PST_RAD_45)
        intWeight += 1 + inclinationDistance;
      else {
        intWeight += 1 - inclinationDistance;
        shot[(desc_index - 1) * stride_b + step_index] += (float)inclinationDistance;
      }
    }

    if (yInFeatRef != 0.0 || xInFeatRef != 0.0) {
      double azimuth = std::atan2(yInFeatRef, xInFeatRef);
      int sel = desc_index >> 2;
      double angularSectorSpan = PST_RAD_45;
      double angularSectorStart = -PST_RAD_PI_7_8;
      double azimuthDistance = (azimuth - (angularSectorStart + angularSectorSpan * sel)) / angularSectorSpan;
      azimuthDistance = std::max(-0.5, std::min(azimuthDistance, 0.5));
      if (azimuthDistance > 0) {
        intWeight += 1 - azimuthDistance;
        int interp_index = (desc_index + 4) % max_sectors;
        shot[interp_index * stride_b + step_index] += (float)azimuthDistance;
      } else {
        int interp_index = (desc_index - 4 + max_sectors) % max_sectors;
        intWeight += 1 + azimuthDistance;
        shot[interp_index * stride_b + step_index] -= (float)azimuthDistance;
      }
    }
    shot[volume_index + step_index] += (float)intWeight;
  }
  /* normalizeHistogram */
  double acc_norm = 0;
  for (int j = 0; j < desc_len; ++j) acc_norm += shot[j] * shot[j];
  acc_norm = std::sqrt(acc_norm);
  for (int j = 0; j < desc_len; ++j) shot[j] /= (float)acc_norm;
  (void)PST_PI;
}

/* ------------------------------------------------------------------------------------------------
 * pcl::computePairFeatures — pcl/features/src/pfh.cpp (float).
 * ---------------------------------------------------------------------------------------------- */
inline bool pair_features(const float *p1, const float *n1, const float *p2, const float *n2, float &f1,
                          float &f2, float &f3, float &f4) {
  float dp[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  {
    float s = dp[0] * dp[0];
    s += dp[1] * dp[1];
    s += dp[2] * dp[2];
    f4 = std::sqrt(s);
  }
  if (f4 == 0.0f) {
    f1 = f2 = f3 = f4 = 0.0f;
    return false;
  }
  float n1c[3] = {n1[0], n1[1], n1[2]}, n2c[3] = {n2[0], n2[1], n2[2]};
  auto dot = [](const float *a, const float *b) {
    float s = a[0] * b[0];
    s += a[1] * b[1];
    s += a[2] * b[2];
    return s;
  };
  float angle1 = dot(n1c, dp) / f4;
  float angle2 = dot(n2c, dp) / f4;
  if (std::acos((double)std::fabs(angle1)) > std::acos((double)std::fabs(angle2))) {
    for (int k = 0; k < 3; ++k) {
      n1c[k] = n2[k];
      n2c[k] = n1[k];
      dp[k] *= -1;
    }
    f3 = -angle2;
  } else
    f3 = angle1;
  float v[3];
  cross3f(dp, n1c, v);
  float v_norm;
  {
    float s = v[0] * v[0];
    s += v[1] * v[1];
    s += v[2] * v[2];
    v_norm = std::sqrt(s);
  }
  if (v_norm == 0.0f) {
    f1 = f2 = f3 = f4 = 0.0f;
    return false;
  }
  for (int k = 0; k < 3; ++k) v[k] /= v_norm;
  float w[3];
  cross3f(n1c, v, w);
  f2 = dot(v, n2c);
  f1 = std::atan2(dot(w, n2c), dot(n1c, n2c));
  return true;
}

inline int clamp_bin(double x, int nbins) {
  double f = std::floor(x);
  int h;
  if (!(f == f))
    h = INT_MIN; /* static_cast<int>(NaN) on x86 */
  else if (f >= 2147483648.0 || f < -2147483648.0)
    h = INT_MIN;
  else
    h = (int)f;
  if (h < 0) h = 0;
  if (h >= nbins) h = nbins - 1;
  return h;
}

/* FPFHEstimation::computePointSPFHSignature — pcl/features/impl/fpfh.hpp.  hist: 33 floats.
 * Note: FPFHEstimation::computePairFeatures returns true unconditionally, so a degenerate pair
 * (zero distance between distinct indices, or dp parallel to the normal) still votes with
 * f1 = f2 = f3 = 0. */
void point_spfh(const float *surf, const float *normals, int sstride, int p_idx, const std::vector<DistIdx> &nb,
                float *hist) {
  const int nb1 = 11;
  const float d_pi = 1.0f / (2.0f * (float)M_PI);
  for (int i = 0; i < 33; ++i) hist[i] = 0.0f;
  float hist_incr = 100.0f / (float)((long)nb.size() - 1);
  const float *p = surf + (size_t)p_idx * sstride;
  const float *np = normals + (size_t)p_idx * 4;
  for (size_t i = 0; i < nb.size(); ++i) {
    if (nb[i].idx == p_idx) continue;
    float f1, f2, f3, f4;
    const bool okp =
        pair_features(p, np, surf + (size_t)nb[i].idx * sstride, normals + (size_t)nb[i].idx * 4, f1, f2, f3, f4);
    if (!okp && (g_variant & V_FPFH_SKIP)) continue;
    /* f1 comes out of atan2f (libm implementations differ by an ulp or two: 1e-6 in bin units covers 2 ulp at
     * |f1| = pi); f2, f3 are plain float products and sums, identical everywhere */
    const double be = (g_variant & V_FPFH_BIN_UP) ? 1e-6 : (g_variant & V_FPFH_BIN_DOWN) ? -1e-6 : 0.0;
    int h = clamp_bin(nb1 * ((f1 + M_PI) * d_pi) + be, nb1);
    hist[h] += hist_incr;
    h = clamp_bin(nb1 * ((f2 + 1.0) * 0.5), nb1);
    hist[11 + h] += hist_incr;
    h = clamp_bin(nb1 * ((f3 + 1.0) * 0.5), nb1);
    hist[22 + h] += hist_incr;
  }
}

/* FPFHEstimation::weightPointSPFHSignature. spfh_row(idx) gives the 33-float SPFH of a surface point. */
template <class RowOf>
void weight_spfh(const std::vector<DistIdx> &nb, RowOf spfh_row, float *out) {
  double sum_f[3] = {0, 0, 0};
  for (int i = 0; i < 33; ++i) out[i] = 0.0f;
  for (size_t i = 0; i < nb.size(); ++i) {
    if (nb[i].d2 == 0) continue;
    float weight = 1.0f / nb[i].d2;
    const float *row = spfh_row(nb[i].idx);
    for (int f = 0; f < 3; ++f)
      for (int b = 0; b < 11; ++b) {
        float val = row[f * 11 + b] * weight;
        sum_f[f] += val;
        out[f * 11 + b] += val;
      }
  }
  for (int f = 0; f < 3; ++f) {
    if (sum_f[f] != 0) sum_f[f] = 100.0 / sum_f[f];
    for (int b = 0; b < 11; ++b) out[f * 11 + b] *= (float)sum_f[f];
  }
}

/* FLANN L2_Simple over D dims. */
inline float sqdistD(const float *a, const float *b, int D) {
  float r = 0.0f;
  for (int i = 0; i < D; ++i) {
    float diff = a[i] - b[i];
    r += diff * diff;
  }
  return r;
}

int match_impl(const float *model, int Km, const float *scene, int Ks, int D, int mode, float thr,
               orc_corr *out, bool parallel) {
  /* KdTreeFLANN::setInputCloud drops rows with any non-finite value (index mapping kept) */
  std::vector<int> valid;
  valid.reserve(Km);
  for (int j = 0; j < Km; ++j) {
    bool ok = true;
    for (int d = 0; d < D; ++d)
      if (!std::isfinite(model[(size_t)j * D + d])) {
        ok = false;
        break;
      }
    if (ok) valid.push_back(j);
  }
  std::vector<orc_corr> tmp(Ks);
  std::vector<unsigned char> keep(Ks, 0);
  auto body = [&](int i) {
    const float *s = scene + (size_t)i * D;
    if (!std::isfinite(s[0])) return;
    DistIdx best{std::numeric_limits<float>::infinity(), INT_MAX}, second = best;
    int found = 0;
    for (int j : valid) {
      DistIdx c{sqdistD(s, model + (size_t)j * D, D), j};
      if (c < best) {
        second = best;
        best = c;
      } else if (c < second)
        second = c;
      ++found;
    }
    if (found == 0) return;
    if (mode == 1) {
      if (best.d2 < thr) {
        tmp[i] = {best.idx, i, best.d2};
        keep[i] = 1;
      }
    } else {
      if (found < 2) return; /* reference reads neigh_sqr_dists[1] of a 1-element answer: undefined; rejected here */
      double tau = best.d2 / second.d2; /* float division, widened (SHOT_demo.cpp:523) */
      if (tau <= 1) {
        tmp[i] = {best.idx, i, best.d2};
        keep[i] = 1;
      }
    }
  };
  if (parallel) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < Ks; ++i) body(i);
  } else {
    for (int i = 0; i < Ks; ++i) body(i);
  }
  int c = 0;
  for (int i = 0; i < Ks; ++i)
    if (keep[i]) out[c++] = tmp[i];
  return c;
}

/* ------------------------------------------------------------------------------------------------
 * RANSAC pose: registration/correspondence_rejection_sample_consensus.hpp →
 * sample_consensus/ransac.hpp on SampleConsensusModelRegistration.
 * source = model keypoints (input_), target = scene keypoints.
 * ---------------------------------------------------------------------------------------------- */
struct RansacResult {
  bool ok;
  float T[16];
  std::vector<int> inliers; /* source (model) indices, one per inlier correspondence */
};

RansacResult ransac_registration(const float *model_kp, int mstride, const float *scene_kp, int sstride,
                                 const std::vector<orc_corr> &corrs, double threshold, int max_iterations) {
  RansacResult res;
  res.ok = false;
  const int n = (int)corrs.size();
  std::vector<int> indices(n), indices_tgt(n);
  for (int i = 0; i < n; ++i) {
    indices[i] = corrs[i].index_query;
    indices_tgt[i] = corrs[i].index_match;
  }
  /* computeOriginalIndexMapping: std::map, last assignment wins for duplicated source indices */
  std::map<int, int> correspondences;
  for (int i = 0; i < n; ++i) correspondences[indices[i]] = indices_tgt[i];
  /* computeSampleDistanceThreshold(cloud, indices) */
  float cov[9], cen[3];
  mean_and_cov(n, [&](int i) { return model_kp + (size_t)indices[i] * mstride; }, cov, cen);
  float ev[3];
  eigen33_values(cov, ev);
  double sample_dist_thresh = ((double)(std::sqrt(ev[0]) + std::sqrt(ev[1]) + std::sqrt(ev[2]))) / 3.0;
  sample_dist_thresh *= sample_dist_thresh;

  std::mt19937 rng(12345u); /* boost::mt19937 rng_alg_, seeded 12345 when random == false */
  auto rnd = [&]() -> int { return (int)(rng() >> 1); }; /* boost::uniform_int<>(0, INT_MAX) */
  std::vector<int> shuffled(indices);
  const int sample_size = 3;
  const int max_sample_checks = 1000;

  int iterations = 0;
  int n_best = -INT_MAX;
  double k = 1.0;
  const double log_probability = std::log(1.0 - 0.99);
  const double one_over_indices = 1.0 / (double)n;
  unsigned skipped = 0;
  const unsigned max_skip = (unsigned)max_iterations * 10;
  std::vector<int> selection, best_model;
  float best_T[16];
  const double thresh2 = threshold * threshold;

  auto apply_count = [&](const float T[16], std::vector<int> *inl) {
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
      const float *s = model_kp + (size_t)indices[i] * mstride;
      const float *t = scene_kp + (size_t)indices_tgt[i] * sstride;
      float e[3];
      for (int r = 0; r < 3; ++r) {
        /* Matrix4f * Vector4f accumulated column by column */
        float v = T[r * 4 + 0] * s[0];
        v += T[r * 4 + 1] * s[1];
        v += T[r * 4 + 2] * s[2];
        v += T[r * 4 + 3];
        e[r] = v - t[r];
      }
      float d = e[0] * e[0];
      d += e[1] * e[1];
      d += e[2] * e[2];
      if ((double)d < thresh2) {
        ++cnt;
        if (inl) inl->push_back(indices[i]);
      }
    }
    return cnt;
  };

  while (iterations < k && skipped < max_skip) {
    /* SampleConsensusModel::getSamples */
    selection.clear();
    if (n < sample_size) break;
    bool good = false;
    for (int iter = 0; iter < max_sample_checks; ++iter) {
      /* drawIndexSample */
      for (int i = 0; i < sample_size; ++i) std::swap(shuffled[i], shuffled[i + (rnd() % (n - i))]);
      selection.assign(shuffled.begin(), shuffled.begin() + sample_size);
      /* isSampleGood */
      const float *p0 = model_kp + (size_t)selection[0] * mstride;
      const float *p1 = model_kp + (size_t)selection[1] * mstride;
      const float *p2 = model_kp + (size_t)selection[2] * mstride;
      auto sq = [](const float *a, const float *b) {
        float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        return dx * dx + dy * dy + dz * dz;
      };
      if ((double)sq(p1, p0) > sample_dist_thresh && (double)sq(p2, p0) > sample_dist_thresh &&
          (double)sq(p2, p1) > sample_dist_thresh) {
        good = true;
        break;
      }
    }
    if (!good) {
      selection.clear();
      break; /* "No samples could be selected" */
    }
    /* computeModelCoefficients → estimateRigidTransformationSVD (double umeyama, cast to float) */
    double src[9], dst[9];
    for (int i = 0; i < 3; ++i) {
      const float *s = model_kp + (size_t)selection[i] * mstride;
      const float *t = scene_kp + (size_t)correspondences[selection[i]] * sstride;
      for (int a = 0; a < 3; ++a) {
        src[i * 3 + a] = s[a];
        dst[i * 3 + a] = t[a];
      }
    }
    double Td[16];
    umeyama3(src, dst, 3, Td);
    float T[16];
    for (int i = 0; i < 16; ++i) T[i] = (float)Td[i];
    int cnt = apply_count(T, nullptr);
    if (cnt > n_best) {
      n_best = cnt;
      best_model = selection;
      std::memcpy(best_T, T, sizeof(T));
      double w = (double)n_best * one_over_indices;
      double p_no_outliers = 1.0 - std::pow(w, (double)sample_size);
      p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
      p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
      k = log_probability / std::log(p_no_outliers);
    }
    ++iterations;
    if (iterations > max_iterations) break;
  }
  if (best_model.empty()) return res;
  res.ok = true;
  std::memcpy(res.T, best_T, sizeof(best_T));
  apply_count(best_T, &res.inliers);
  return res;
}


/* RANSAC on one consensus / voter set and the filtered correspondences, appended to the outputs the way
 * GeometricConsistencyGrouping::recognize and Hough3DGrouping::recognize do. */
void emit_instance(const float *model_kp, int mstride, const float *scene_kp, int sstride,
                   const std::vector<orc_corr> &temp, double threshold, float *transforms, int max_inst,
                   int *inst_offsets, orc_corr *inst_corrs, int corr_cap, int &n_inst, int &written) {
  RansacResult rr = ransac_registration(model_kp, mstride, scene_kp, sstride, temp, threshold, 10000);
  std::vector<orc_corr> filtered;
  float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  if (!rr.ok || rr.inliers.size() < 3) {
    filtered = temp; /* identity + unfiltered set */
  } else {
    /* index_to_correspondence keyed by index_query: last one wins */
    std::unordered_map<int, int> index_to_corr;
    for (int t = 0; t < (int)temp.size(); ++t) index_to_corr[temp[t].index_query] = t;
    for (int inl : rr.inliers) filtered.push_back(temp[index_to_corr[inl]]);
    std::memcpy(T, rr.T, sizeof(T));
  }
  if (n_inst < max_inst) {
    if (transforms) std::memcpy(transforms + (size_t)n_inst * 16, T, sizeof(T));
    if (inst_corrs && written + (int)filtered.size() <= corr_cap) {
      std::memcpy(inst_corrs + written, filtered.data(), filtered.size() * sizeof(orc_corr));
      written += (int)filtered.size();
    }
    if (inst_offsets) inst_offsets[n_inst + 1] = written;
  }
  ++n_inst;
}
}  // namespace

/* ================================================================================================
 * C API
 * ============================================================================================== */
extern "C" {

void orc_set_variant(unsigned v) { g_variant = v; }
unsigned orc_get_variant(void) { return g_variant; }

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int64_t orc_radius_search(const float *surf, int n, int sstride, const float *q, int nq, int qstride,
                          double radius, int64_t *offsets, int *idx, float *d2, int64_t cap) {
  CpuGrid g;
  g.build(surf, n, sstride, cell_for_radius(radius));
  std::vector<std::vector<DistIdx>> all(nq);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < nq; ++i) g.radius(q + (size_t)i * qstride, radius, all[i]);
  int64_t total = 0;
  for (int i = 0; i < nq; ++i) {
    offsets[i] = total;
    total += (int64_t)all[i].size();
  }
  offsets[nq] = total;
  if (idx && d2 && cap >= total) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < nq; ++i) {
      int64_t o = offsets[i];
      for (size_t j = 0; j < all[i].size(); ++j) {
        idx[o + (int64_t)j] = all[i][j].idx;
        d2[o + (int64_t)j] = all[i][j].d2;
      }
    }
  }
  return total;
}

int orc_knn_search(const float *surf, int n, int sstride, const float *q, int nq, int qstride, int k, int *idx,
                   float *d2) {
  CpuGrid g;
  g.build(surf, n, sstride, auto_cell_for_knn(surf, n, sstride, k));
  int kk = std::min(k, g.n_valid);
#pragma omp parallel
  {
    std::vector<DistIdx> nb;
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < nq; ++i) {
      g.knn(q + (size_t)i * qstride, k, nb);
      for (int j = 0; j < k; ++j) {
        bool have = j < (int)nb.size();
        idx[(size_t)i * k + j] = have ? nb[j].idx : -1;
        d2[(size_t)i * k + j] = have ? nb[j].d2 : std::numeric_limits<float>::infinity();
      }
    }
  }
  return kk;
}

int orc_normals(const float *surf, int n, int sstride, const float *q, int nq, int qstride, int k, double radius,
                const float *vp, float *out) {
  /* Feature::initCompute: exactly one of k / radius */
  if ((k != 0) == (radius != 0.0)) return -1;
  CpuGrid g;
  g.build(surf, n, sstride, k ? auto_cell_for_knn(surf, n, sstride, k) : cell_for_radius(radius));
  float vpx = vp ? vp[0] : 0.f, vpy = vp ? vp[1] : 0.f, vpz = vp ? vp[2] : 0.f;
#pragma omp parallel
  {
    std::vector<DistIdx> nb;
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < nq; ++i) {
      const float *p = q + (size_t)i * qstride;
      float *o = out + (size_t)i * 4;
      if (k)
        g.knn(p, k, nb);
      else
        g.radius(p, radius, nb);
      if (nb.empty() || nb.size() < 3) { /* search failure, or computePointNormal: < 3 neighbours */
        o[0] = o[1] = o[2] = o[3] = kNaNf;
        continue;
      }
      float cov[9], cen[3];
      mean_and_cov((int)nb.size(), [&](int j) { return surf + (size_t)nb[j].idx * sstride; }, cov, cen);
      /* solvePlaneParameters */
      float ev, vec[3];
      eigen33_smallest(cov, ev, vec);
      float eig_sum = cov[0] + cov[4] + cov[8];
      float curvature = (eig_sum != 0) ? std::fabs(ev / eig_sum) : 0.0f;
      /* flipNormalTowardsViewpoint */
      float dx = vpx - p[0], dy = vpy - p[1], dz = vpz - p[2];
      float cos_theta = (dx * vec[0] + dy * vec[1] + dz * vec[2]);
      if (cos_theta < 0) {
        vec[0] *= -1;
        vec[1] *= -1;
        vec[2] *= -1;
      }
      o[0] = vec[0];
      o[1] = vec[1];
      o[2] = vec[2];
      o[3] = curvature;
    }
  }
  return 0;
}

int orc_shot_lrf(const float *surf, int n, int sstride, const float *kp, int K, int kstride, double radius,
                 float *out) {
  CpuGrid g;
  g.build(surf, n, sstride, cell_for_radius(radius));
#pragma omp parallel
  {
    std::vector<DistIdx> nb;
#pragma omp for schedule(dynamic, 32)
    for (int i = 0; i < K; ++i) {
      const float *c = kp + (size_t)i * kstride;
      g.radius(c, radius, nb);
      shot_local_rf(surf, sstride, c, nb, radius, out + (size_t)i * 9);
    }
  }
  return 0;
}

int orc_shot352(const float *surf, const float *normals, int n, int sstride, const float *kp, int K, int kstride,
                double radius, float *desc, float *rf_out) {
  CpuGrid g;
  g.build(surf, n, sstride, cell_for_radius(radius));
  /* SHOTEstimationOMP::initCompute: LRF pass over all keypoints first (its own radius search) ... */
  std::vector<float> frames((size_t)K * 9);
#pragma omp parallel
  {
    std::vector<DistIdx> nb;
#pragma omp for schedule(dynamic, 32)
    for (int i = 0; i < K; ++i) {
      const float *c = kp + (size_t)i * kstride;
      g.radius(c, radius, nb);
      shot_local_rf(surf, sstride, c, nb, radius, &frames[(size_t)i * 9]);
    }
  }
  /* ... then computeFeature (second radius search per keypoint) */
#pragma omp parallel
  {
    std::vector<DistIdx> nb;
#pragma omp for schedule(dynamic, 32)
    for (int i = 0; i < K; ++i) {
      const float *c = kp + (size_t)i * kstride;
      const float *rf = &frames[(size_t)i * 9];
      float *d = desc + (size_t)i * 352;
      float *ro = rf_out + (size_t)i * 9;
      bool lrf_nan = !std::isfinite(rf[0]) || !std::isfinite(rf[3]) || !std::isfinite(rf[6]);
      bool bad = !finite3(c) || lrf_nan;
      if (!bad) {
        g.radius(c, radius, nb);
        bad = nb.empty();
      }
      if (bad) {
        for (int k = 0; k < 352; ++k) d[k] = kNaNf;
        for (int k = 0; k < 9; ++k) ro[k] = kNaNf;
        continue;
      }
      point_shot352(surf, normals, sstride, c, rf, nb, radius, d);
      for (int k = 0; k < 9; ++k) ro[k] = rf[k];
    }
  }
  return 0;
}

int orc_fpfh33(const float *surf, const float *normals, int n, int sstride, const float *q, int nq, int qstride,
               double radius, float *out) {
  CpuGrid g;
  g.build(surf, n, sstride, cell_for_radius(radius));
  const bool every_point = (q == nullptr);
  if (every_point) {
    q = surf;
    nq = n;
    qstride = sstride;
  }
  /* computeSPFHSignatures: which surface points need an SPFH */
  std::vector<unsigned char> need(n, every_point ? 1 : 0);
  if (!every_point) {
#pragma omp parallel
    {
      std::vector<DistIdx> nb;
#pragma omp for schedule(dynamic, 64)
      for (int i = 0; i < nq; ++i) {
        g.radius(q + (size_t)i * qstride, radius, nb);
        for (auto &e : nb) need[e.idx] = 1; /* benign race: all writers store 1 */
      }
    }
  }
  std::vector<float> spfh((size_t)n * 33, 0.0f);
#pragma omp parallel
  {
    std::vector<DistIdx> nb;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
      if (!need[i]) continue;
      g.radius(surf + (size_t)i * sstride, radius, nb);
      if (nb.empty()) continue; /* row stays zero */
      point_spfh(surf, normals, sstride, i, nb, &spfh[(size_t)i * 33]);
    }
  }
#pragma omp parallel
  {
    std::vector<DistIdx> nb;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < nq; ++i) {
      float *o = out + (size_t)i * 33;
      g.radius(q + (size_t)i * qstride, radius, nb);
      if (nb.empty()) {
        for (int k = 0; k < 33; ++k) o[k] = kNaNf;
        continue;
      }
      weight_spfh(nb, [&](int idx) { return &spfh[(size_t)idx * 33]; }, o);
    }
  }
  return 0;
}

int orc_match(const float *model, int Km, const float *scene, int Ks, int D, int mode, float thr, orc_corr *out) {
  return match_impl(model, Km, scene, Ks, D, mode, thr, out, false);
}

int orc_match_omp(const float *model, int Km, const float *scene, int Ks, int D, int mode, float thr,
                  orc_corr *out) {
  return match_impl(model, Km, scene, Ks, D, mode, thr, out, true);
}

int orc_gc_recognize(const float *model_kp, int Km, int mstride, const float *scene_kp, int Ks, int sstride,
                     const orc_corr *corrs_in, int C, double gc_size, int gc_threshold, float *transforms,
                     int max_inst, int *inst_offsets, orc_corr *inst_corrs, int corr_cap) {
  (void)Km;
  (void)Ks;
  /* std::sort by distance (unstable in PCL); ties broken by original position for determinism */
  std::vector<int> perm(C);
  for (int i = 0; i < C; ++i) perm[i] = i;
  std::sort(perm.begin(), perm.end(), [&](int a, int b) {
    return corrs_in[a].distance < corrs_in[b].distance ||
           (corrs_in[a].distance == corrs_in[b].distance && a < b);
  });
  std::vector<orc_corr> corrs(C);
  for (int i = 0; i < C; ++i) corrs[i] = corrs_in[perm[i]];
  std::vector<unsigned char> taken(C, 0);
  std::vector<int> consensus;
  int n_inst = 0;
  int written = 0;
  if (inst_offsets && max_inst >= 0) inst_offsets[0] = 0;
  auto norm3f = [](const float *a, const float *b) {
    float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    float s = d0 * d0;
    s += d1 * d1;
    s += d2 * d2;
    return std::sqrt(s);
  };
  for (int i = 0; i < C; ++i) {
    if (taken[i]) continue;
    consensus.clear();
    consensus.push_back(i);
    for (int j = 0; j < C; ++j) {
      if (j == i || taken[j]) continue;
      bool good = true;
      const float *sj = scene_kp + (size_t)corrs[j].index_match * sstride;
      const float *mj = model_kp + (size_t)corrs[j].index_query * mstride;
      for (size_t kk = 0; kk < consensus.size(); ++kk) {
        const orc_corr &ck = corrs[consensus[kk]];
        const float *sk = scene_kp + (size_t)ck.index_match * sstride;
        const float *mk = model_kp + (size_t)ck.index_query * mstride;
        double distance = std::fabs(norm3f(sk, sj) - norm3f(mk, mj));
        if (distance > gc_size) {
          good = false;
          break;
        }
      }
      if (good) consensus.push_back(j);
    }
    if ((int)consensus.size() > gc_threshold) {
      std::vector<orc_corr> temp;
      for (int c : consensus) {
        temp.push_back(corrs[c]);
        taken[c] = 1;
      }
      emit_instance(model_kp, mstride, scene_kp, sstride, temp, gc_size, transforms, max_inst, inst_offsets, inst_corrs,
                    corr_cap, n_inst, written);
    }
  }
  return n_inst;
}

void orc_eigen33_smallest(const float cov9[9], float *eigenvalue, float evec3[3]) {
  eigen33_smallest(cov9, *eigenvalue, evec3);
}
void orc_eigh3_f64(const double a9[9], double evals3[3], double evecs9[9]) { eigh3_f64(a9, evals3, evecs9); }
/* ------------------------------------------------------------------------------------------------
 * Keypoint extraction.
 * pcl::UniformSampling<PointT>::applyFilter (pcl 1.8 filters/impl/uniform_sampling.hpp; SHOT.cpp:314-323):
 *   inverse_leaf = 1 / leaf (float); ijk = floor(p * inverse_leaf); leaf index = (ijk - min_b) . divb_mul;
 *   per leaf keep the point with the smaller (p - (float)ijk).squaredNorm() on the homogeneous 4-vectors
 *   (w: 1 - 0), strict <, i.e. the first point wins ties.  Output order: hash-map order in PCL, defined
 *   here as ascending leaf index.
 * pcl::VoxelGrid<PointT>::applyFilter (filters/impl/voxel_grid.hpp; SHOT_demo.cpp:413-417): same lattice;
 *   centroid of every occupied voxel, ascending voxel index.  PCL sums in float32 in the order of an
 *   unstable sort; the centroid is evaluated in float64 here (order independent) and rounded once.
 * Non-finite points are skipped.  Returns the number of keypoints, or -1 when the lattice would need
 * more than 2^26 leaves ("Leaf size is too small for the input dataset").
 * ---------------------------------------------------------------------------------------------- */
namespace {
struct LatticeO {
  float inv[3];
  long long mn[3], dim[3];
};
bool make_lattice_o(const float *xyz, int n, int stride, const float leaf[3], LatticeO &L) {
  for (int a = 0; a < 3; ++a) L.inv[a] = 1.0f / leaf[a];
  long long mx[3] = {LLONG_MIN, LLONG_MIN, LLONG_MIN};
  for (int a = 0; a < 3; ++a) L.mn[a] = LLONG_MAX;
  bool any = false;
  for (int i = 0; i < n; ++i) {
    const float *p = xyz + (size_t)i * stride;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    any = true;
    for (int a = 0; a < 3; ++a) {
      const long long c = (long long)std::floor(p[a] * L.inv[a]);
      L.mn[a] = std::min(L.mn[a], c);
      mx[a] = std::max(mx[a], c);
    }
  }
  if (!any) {
    for (int a = 0; a < 3; ++a) L.dim[a] = 0;
    return true;
  }
  for (int a = 0; a < 3; ++a) L.dim[a] = mx[a] - L.mn[a] + 1;
  return L.dim[0] * L.dim[1] * L.dim[2] <= (1ll << 26);
}
}  // namespace

int orc_uniform_sampling(const float *xyz, int n, int stride, double leaf, float *out_xyz, int *out_index) {
  LatticeO L;
  const float lf[3] = {(float)leaf, (float)leaf, (float)leaf};
  if (!make_lattice_o(xyz, n, stride, lf, L)) return -1;
  const long long nleaf = L.dim[0] * L.dim[1] * L.dim[2];
  std::vector<int> keep((size_t)nleaf, -1);
  std::vector<float> best((size_t)nleaf, 0.f);
  for (int i = 0; i < n; ++i) {
    const float *p = xyz + (size_t)i * stride;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    long long c[3];
    for (int a = 0; a < 3; ++a) c[a] = (long long)std::floor(p[a] * L.inv[a]);
    const long long idx = (c[0] - L.mn[0]) + L.dim[0] * ((c[1] - L.mn[1]) + L.dim[1] * (c[2] - L.mn[2]));
    const float d0 = p[0] - (float)c[0], d1 = p[1] - (float)c[1], d2 = p[2] - (float)c[2];
    float diff = d0 * d0;
    diff += d1 * d1;
    diff += d2 * d2;
    diff += 1.0f;
    if (keep[(size_t)idx] < 0 || diff < best[(size_t)idx]) {
      keep[(size_t)idx] = i;
      best[(size_t)idx] = diff;
    }
  }
  int m = 0;
  for (long long l = 0; l < nleaf; ++l)
    if (keep[(size_t)l] >= 0) {
      const float *p = xyz + (size_t)keep[(size_t)l] * stride;
      out_xyz[(size_t)m * 3 + 0] = p[0], out_xyz[(size_t)m * 3 + 1] = p[1], out_xyz[(size_t)m * 3 + 2] = p[2];
      if (out_index) out_index[m] = keep[(size_t)l];
      ++m;
    }
  return m;
}

int orc_voxel_grid(const float *xyz, int n, int stride, float lx, float ly, float lz, float *out_xyz) {
  LatticeO L;
  const float lf[3] = {lx, ly, lz};
  if (!make_lattice_o(xyz, n, stride, lf, L)) return -1;
  const long long nleaf = L.dim[0] * L.dim[1] * L.dim[2];
  std::vector<double> sums((size_t)nleaf * 3, 0.0);
  std::vector<int> cnt((size_t)nleaf, 0);
  for (int i = 0; i < n; ++i) {
    const float *p = xyz + (size_t)i * stride;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    long long c[3];
    for (int a = 0; a < 3; ++a) c[a] = (long long)std::floor(p[a] * L.inv[a]);
    const long long idx = (c[0] - L.mn[0]) + L.dim[0] * ((c[1] - L.mn[1]) + L.dim[1] * (c[2] - L.mn[2]));
    for (int a = 0; a < 3; ++a) sums[(size_t)idx * 3 + a] += (double)p[a];
    ++cnt[(size_t)idx];
  }
  int m = 0;
  for (long long l = 0; l < nleaf; ++l)
    if (cnt[(size_t)l] > 0) {
      const double inv = 1.0 / (double)cnt[(size_t)l];
      for (int a = 0; a < 3; ++a) out_xyz[(size_t)m * 3 + a] = (float)(sums[(size_t)l * 3 + a] * inv);
      ++m;
    }
  return m;
}

/* ------------------------------------------------------------------------------------------------
 * pcl::Hough3DGrouping::recognize (pcl 1.8 recognition/impl/cg/hough_3d.hpp + recognition/hough_3d.cpp
 * HoughSpace3D), as the reference configures it (SHOT.cpp:433-470): reference frames given (setInputRf /
 * setSceneRf), setUseInterpolation(false), setUseDistanceWeight(true).
 *   train: centroid = float running sum of the model keypoints / n; model_vote_i = (x.d, y.d, z.d) with
 *          d = centroid - keypoint_i and x, y, z the axes of keypoint i's frame.
 *   houghVoting: scene_vote = (float) rf_x[a] * mv.x + rf_y[a] * mv.y + rf_z[a] * mv.z + scene_point[a],
 *          widened to double; space = [min, max] per axis, bin_count = ceil((max - min) / bin); weight =
 *          1 - distance / max_distance where max_distance is tracked only with interpolation (it stays
 *          -FLT_MAX), i.e. 1.0; HoughSpace3D::vote drops a vote whose bin coordinate is out of range.
 *   findMaxima: bins with value >= threshold (negative threshold: that fraction of the largest bin), in
 *          ascending bin index, voters in voting order; then RANSAC with inlier threshold = bin size.
 * Correspondences whose frames are not finite are skipped (PCL would index out of range with them).
 * rf: K x 9 floats (x, y, z axes).
 * ---------------------------------------------------------------------------------------------- */
int orc_hough3d_recognize(const float *model_kp, const float *model_rf, int Km, int mstride, const float *scene_kp,
                          const float *scene_rf, int Ks, int sstride, const orc_corr *corrs, int C, double bin_size,
                          double threshold, float *transforms, int max_inst, int *inst_offsets, orc_corr *inst_corrs,
                          int corr_cap) {
  (void)Ks;
  if (inst_offsets) inst_offsets[0] = 0;
  if (C <= 0 || Km <= 0) return 0;
  float centroid[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < Km; ++i)
    for (int a = 0; a < 3; ++a) centroid[a] += model_kp[(size_t)i * mstride + a];
  for (int a = 0; a < 3; ++a) centroid[a] /= (float)Km;
  std::vector<float> mv((size_t)Km * 3);
  for (int i = 0; i < Km; ++i) {
    const float *p = model_kp + (size_t)i * mstride, *r = model_rf + (size_t)i * 9;
    const float d[3] = {centroid[0] - p[0], centroid[1] - p[1], centroid[2] - p[2]};
    for (int a = 0; a < 3; ++a) {
      float v = r[a * 3 + 0] * d[0];
      v += r[a * 3 + 1] * d[1];
      v += r[a * 3 + 2] * d[2];
      mv[(size_t)i * 3 + a] = v;
    }
  }
  std::vector<double> votes((size_t)C * 3);
  std::vector<char> valid(C, 0);
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (int i = 0; i < C; ++i) {
    const float *m = &mv[(size_t)corrs[i].index_query * 3];
    const float *r = scene_rf + (size_t)corrs[i].index_match * 9;
    const float *sp = scene_kp + (size_t)corrs[i].index_match * sstride;
    bool fin = true;
    for (int a = 0; a < 3; ++a) {
      float t = r[0 * 3 + a] * m[0];
      t += r[1 * 3 + a] * m[1];
      t += r[2 * 3 + a] * m[2];
      t += sp[a];
      votes[(size_t)i * 3 + a] = (double)t;
      fin = fin && std::isfinite(t);
    }
    valid[i] = fin;
    if (fin)
      for (int a = 0; a < 3; ++a) {
        mn[a] = std::min(mn[a], votes[(size_t)i * 3 + a]);
        mx[a] = std::max(mx[a], votes[(size_t)i * 3 + a]);
      }
  }
  if (mn[0] > mx[0]) return 0;
  long long cnt[3], total = 1;
  for (int a = 0; a < 3; ++a) {
    cnt[a] = (long long)std::ceil((mx[a] - mn[a]) / bin_size);
    total *= std::max(cnt[a], 0ll);
  }
  if (total == 0 || total > (1ll << 27)) return total == 0 ? 0 : -1;
  std::vector<int> space((size_t)total, 0);
  std::vector<std::vector<int>> voters((size_t)total);
  for (int i = 0; i < C; ++i) {
    if (!valid[i]) continue;
    long long index = 0, mul = 1;
    bool in = true;
    for (int a = 0; a < 3; ++a) {
      const int ci = (int)std::floor((votes[(size_t)i * 3 + a] - mn[a]) / bin_size);
      if (ci < 0 || ci >= cnt[a]) {
        in = false;
        break;
      }
      index += mul * ci;
      mul *= cnt[a];
    }
    if (!in) continue;
    space[(size_t)index] += 1; /* weight 1.0 */
    voters[(size_t)index].push_back(i);
  }
  double thr = threshold;
  if (thr < 0) {
    const double hmax = (double)*std::max_element(space.begin(), space.end());
    thr = (thr >= -1.0) ? -thr * hmax : hmax;
  }
  int n_inst = 0, written = 0;
  for (long long b = 0; b < total; ++b) {
    if (space[(size_t)b] <= 0 || (double)space[(size_t)b] < thr) continue;
    std::vector<orc_corr> temp;
    for (int v : voters[(size_t)b]) temp.push_back(corrs[v]);
    emit_instance(model_kp, mstride, scene_kp, sstride, temp, bin_size, transforms, max_inst, inst_offsets, inst_corrs,
                  corr_cap, n_inst, written);
  }
  return n_inst;
}

/* pcl::IterativeClosestPoint<PointT, PointT, float>::align + getFitnessScore (pcl 1.8 registration/impl/icp.hpp
 * computeTransformation, default_convergence_criteria.hpp hasConverged, registration.hpp getFitnessScore) as
 * the reference configures it: setMaximumIterations only (SHOT.cpp:177-192, SHOT_demo.cpp:604-663).
 * Correspondences: nearest target point of every source point (brute force, L2_Simple float, lowest index on
 * ties), kept when d2 <= max_dist^2.  The rigid fit is pcl::umeyama without scaling; PCL evaluates it in
 * float32 through Eigen's JacobiSVD, here it is the float64 umeyama3 cast to float (agreement to float32
 * rounding, not bit for bit — tests use a tolerance). */
int orc_icp_align(const float *source, int ns, int sstride, const float *target, int nt, int tstride, int max_iterations,
                  double max_corr_dist, double transformation_epsilon, double euclidean_fitness_epsilon,
                  const float *guess, float *final_T, float *aligned, double *fitness, int *converged, int *iterations) {
  float fin[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  if (guess) memcpy(fin, guess, sizeof(fin));
  memcpy(final_T, fin, sizeof(fin));
  if (converged) *converged = 0;
  if (iterations) *iterations = 0;
  if (fitness) *fitness = DBL_MAX;
  std::vector<float> src, cur, tgt;
  for (int i = 0; i < ns; ++i) {
    const float *p = source + (size_t)i * sstride;
    if (std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2])) src.insert(src.end(), p, p + 3);
  }
  for (int i = 0; i < nt; ++i) {
    const float *p = target + (size_t)i * tstride;
    if (std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2])) tgt.insert(tgt.end(), p, p + 3);
  }
  const int n = (int)src.size() / 3, m = (int)tgt.size() / 3;
  if (n == 0 || m == 0) return 0;
  auto apply = [](const float *T, std::vector<float> &pts) {
    for (size_t i = 0; i < pts.size() / 3; ++i) {
      const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
      for (int r = 0; r < 3; ++r) {
        float v = T[r * 4 + 0] * x;
        v += T[r * 4 + 1] * y;
        v += T[r * 4 + 2] * z;
        v += T[r * 4 + 3];
        pts[3 * i + r] = v;
      }
    }
  };
  std::vector<int> nn((size_t)n);
  std::vector<float> nd((size_t)n);
  auto nearest = [&](const std::vector<float> &pts) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      int best = -1;
      float bd = 0.f;
      for (int j = 0; j < m; ++j) {
        const float d = sqdist3(&pts[3 * (size_t)i], &tgt[3 * (size_t)j]);
        if (best < 0 || d < bd) {
          best = j;
          bd = d;
        }
      }
      nn[i] = best;
      nd[i] = bd;
    }
  };
  cur = src;
  apply(fin, cur);
  const double md = (max_corr_dist > 0.0) ? max_corr_dist : std::sqrt(DBL_MAX);
  const double max_d2 = md * md;
  const double rotation_threshold = 1.0 - transformation_epsilon, translation_threshold = transformation_epsilon;
  double prev_mse = DBL_MAX;
  int it = 0;
  bool conv = false;
  std::vector<double> a, b;
  for (;;) {
    nearest(cur);
    a.clear();
    b.clear();
    double sum_d = 0.0;
    for (int i = 0; i < n; ++i) {
      if ((double)nd[i] > max_d2) continue;
      for (int k = 0; k < 3; ++k) {
        a.push_back(cur[3 * (size_t)i + k]);
        b.push_back(tgt[3 * (size_t)nn[i] + k]);
      }
      sum_d += (double)nd[i];
    }
    const int cnt = (int)a.size() / 3;
    if (cnt < 3) break; /* NO_CORRESPONDENCES: converged_ stays false */
    double Td[16];
    umeyama3(a.data(), b.data(), cnt, Td);
    float T[16];
    for (int k = 0; k < 16; ++k) T[k] = (float)Td[k];
    apply(T, cur);
    float nf[16];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        float v = T[r * 4 + 0] * fin[0 * 4 + c];
        v += T[r * 4 + 1] * fin[1 * 4 + c];
        v += T[r * 4 + 2] * fin[2 * 4 + c];
        v += T[r * 4 + 3] * fin[3 * 4 + c];
        nf[r * 4 + c] = v;
      }
    memcpy(fin, nf, sizeof(fin));
    ++it;
    if (it >= max_iterations) {
      conv = true;
      break;
    }
    const double cos_angle = 0.5 * ((double)T[0] + (double)T[5] + (double)T[10] - 1.0);
    const double tr2 = (double)T[3] * T[3] + (double)T[7] * T[7] + (double)T[11] * T[11];
    if (cos_angle >= rotation_threshold && tr2 <= translation_threshold) {
      conv = true;
      break;
    }
    const double mse = sum_d / cnt;
    if (std::fabs(mse - prev_mse) < 1e-12) {
      conv = true;
      break;
    }
    if (std::fabs(mse - prev_mse) / prev_mse < euclidean_fitness_epsilon) {
      conv = true;
      break;
    }
    prev_mse = mse;
  }
  memcpy(final_T, fin, sizeof(fin));
  if (converged) *converged = conv ? 1 : 0;
  if (iterations) *iterations = it;
  cur = src;
  apply(fin, cur);
  if (aligned) { /* rows of non-finite source points stay NaN */
    int k = 0;
    for (int i = 0; i < ns; ++i) {
      const float *p = source + (size_t)i * sstride;
      const bool ok = std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]);
      for (int c = 0; c < 3; ++c) aligned[3 * (size_t)i + c] = ok ? cur[3 * (size_t)k + c] : NAN;
      if (ok) ++k;
    }
  }
  if (fitness) {
    nearest(cur);
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += (double)nd[i];
    *fitness = s / n;
  }
  return 0;
}

/* glibc rand() (TYPE_3 additive feedback generator, random_r.c): r[i] = r[i-3] + r[i-31], seeded with the
 * Lehmer sequence 16807 * r mod (2^31 - 1), the first 310 outputs discarded, output = r >> 1.  BOARD draws its
 * "random orthogonal axis" from the process-global rand(); the restatement carries the stream explicitly. */
struct GlibcRand {
  uint32_t r[34];
  int k;
  std::vector<uint32_t> hist; /* sliding window of the last 31 values */
  explicit GlibcRand(unsigned seed) { reseed(seed); }
  void reseed(unsigned seed) {
    if (seed == 0) seed = 1;
    int32_t v[34];
    v[0] = (int32_t)seed;
    for (int i = 1; i < 31; ++i) {
      int64_t hi = v[i - 1] / 127773, lo = v[i - 1] % 127773;
      int64_t word = 16807 * lo - 2836 * hi;
      if (word < 0) word += 2147483647;
      v[i] = (int32_t)word;
    }
    hist.assign(v, v + 31);
    for (int i = 31; i < 34; ++i) hist.push_back(hist[i - 31]);
    for (int i = 34; i < 344; ++i) hist.push_back(hist[i - 31] + hist[i - 3]);
    hist.erase(hist.begin(), hist.end() - 31);
  }
  int next() {
    const uint32_t o = hist[0] + hist[28]; /* r[i-31] + r[i-3] */
    hist.erase(hist.begin());
    hist.push_back(o);
    return (int)(o >> 1);
  }
};

namespace {
inline float dot3f(const float *a, const float *b) { /* Eigen 3-vector dot: ((a0 b0 + a1 b1) + a2 b2) */
  float s = a[0] * b[0];
  s += a[1] * b[1];
  s += a[2] * b[2];
  return s;
}
inline void normalize3f(float *v) { /* Eigen normalize(): v /= sqrt(squaredNorm) */
  const float n = std::sqrt(dot3f(v, v));
  v[0] /= n;
  v[1] /= n;
  v[2] /= n;
}
/* board.hpp directedOrthogonalAxis + projectPointOnPlane */
inline void directed_orthogonal_axis(const float *axis, const float *origin, const float *point, float *out) {
  float xo[3] = {point[0] - origin[0], point[1] - origin[1], point[2] - origin[2]};
  const float t = dot3f(axis, xo);
  float proj[3] = {point[0] - t * axis[0], point[1] - t * axis[1], point[2] - t * axis[2]};
  out[0] = proj[0] - origin[0];
  out[1] = proj[1] - origin[1];
  out[2] = proj[2] - origin[2];
  normalize3f(out);
}
/* board.hpp getAngleBetweenUnitVectors */
inline float angle_between_unit(const float *v1, const float *v2, const float *axis) {
  float o[3];
  cross3f(v1, v2, o);
  float a = std::acos(std::max(-1.0f, std::min(1.0f, dot3f(v1, v2))));
  return dot3f(o, axis) < 0.f ? (2 * (float)M_PI - a) : a;
}
constexpr int BOARD_MAX_SECTORS = 64;

/* board.hpp computePointLRF.  nb: the keypoint's support in (d2, index) order.  rnd: the two rand() values
 * drawn for this keypoint (find_holes only).  out: x, y, z axes.  Returns false when the frame is NaN. */
bool board_point_lrf(const float *surf, int sstride, const float *normals, const float *c, const std::vector<DistIdx> &nb,
                     const std::vector<DistIdx> &nbt, const orc_board_params &bp, const int rnd[2], float *out) {
  const int n = (int)nb.size();
  auto set_nan = [&]() {
    for (int k = 0; k < 9; ++k) out[k] = kNaNf;
    return false;
  };
  if (n < 6) return set_nan();
  /* planeFitting: centroid, centred points, right singular vector of the smallest singular value.  PCL runs
   * Eigen's float JacobiSVD; here: float64 centroid and scatter, eigenvector of the smallest eigenvalue. */
  double mean[3] = {0, 0, 0};
  for (const DistIdx &e : nb)
    for (int a = 0; a < 3; ++a) mean[a] += surf[(size_t)e.idx * sstride + a];
  for (int a = 0; a < 3; ++a) mean[a] /= n;
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (const DistIdx &e : nb) {
    const double d[3] = {surf[(size_t)e.idx * sstride] - mean[0], surf[(size_t)e.idx * sstride + 1] - mean[1],
                         surf[(size_t)e.idx * sstride + 2] - mean[2]};
    cov[0] += d[0] * d[0];
    cov[1] += d[0] * d[1];
    cov[2] += d[0] * d[2];
    cov[4] += d[1] * d[1];
    cov[5] += d[1] * d[2];
    cov[8] += d[2] * d[2];
  }
  cov[3] = cov[1];
  cov[6] = cov[2];
  cov[7] = cov[5];
  double w[3], V[9];
  eigh3_f64(cov, w, V);
  float z[3] = {(float)V[0], (float)V[3], (float)V[6]}; /* column 0: smallest eigenvalue */
  /* normalDisambiguation: sign by the mean of the finite support normals */
  double nm[3] = {0, 0, 0};
  for (const DistIdx &e : nb) {
    const float *q = normals + (size_t)e.idx * 4;
    if (std::isfinite(q[0]) && std::isfinite(q[1]) && std::isfinite(q[2]))
      for (int a = 0; a < 3; ++a) nm[a] += q[a];
  }
  if (nm[0] != 0.0 || nm[1] != 0.0 || nm[2] != 0.0) {
    if ((double)z[0] * nm[0] + (double)z[1] * nm[1] + (double)z[2] * nm[2] < 0.0) {
      z[0] = -z[0];
      z[1] = -z[1];
      z[2] = -z[2];
    }
  }
  for (int a = 0; a < 3; ++a) out[6 + a] = z[a];
  const float tangent_radius = bp.tangent_radius;
  const std::vector<DistIdx> *sup = &nbt; /* support of the x axis: the tangent-radius search when it differs */
  const float radius2 = tangent_radius * tangent_radius;
  const float margin_distance2 = bp.margin_thresh * bp.margin_thresh * radius2;
  const int S = bp.check_margin_array_size;
  float min_normal_cos = FLT_MAX;
  int min_normal_index = -1;
  bool margin_point_found = false;
  bool check[BOARD_MAX_SECTORS];
  float min_angle[BOARD_MAX_SECTORS], max_angle[BOARD_MAX_SECTORS], min_angle_normal[BOARD_MAX_SECTORS],
      max_angle_normal[BOARD_MAX_SECTORS];
  float x_axis[3] = {0, 0, 0}, y_axis[3];
  float max_boundary_angle = 0.f;
  if (bp.find_holes) {
    /* randomOrthogonalAxis */
    const float r0 = ((float)rnd[0] / (float)RAND_MAX) * 2.0f - 1.0f;
    const float r1 = ((float)rnd[1] / (float)RAND_MAX) * 2.0f - 1.0f;
    auto near0 = [](float v) { return std::fabs(v - 0.0f) < 1e-8f; };
    if (!near0(z[2])) {
      x_axis[0] = r0;
      x_axis[1] = r1;
      x_axis[2] = -(z[0] * x_axis[0] + z[1] * x_axis[1]) / z[2];
    } else if (!near0(z[1])) {
      x_axis[0] = r0;
      x_axis[2] = r1;
      x_axis[1] = -(z[0] * x_axis[0] + z[2] * x_axis[2]) / z[1];
    } else if (!near0(z[0])) {
      x_axis[1] = r0;
      x_axis[2] = r1;
      x_axis[0] = -(z[1] * x_axis[1] + z[2] * x_axis[2]) / z[0];
    }
    normalize3f(x_axis);
    for (int i = 0; i < S; ++i) {
      check[i] = false;
      min_angle[i] = FLT_MAX;
      max_angle[i] = -FLT_MAX;
      min_angle_normal[i] = -1.0f;
      max_angle_normal[i] = -1.0f;
    }
    max_boundary_angle = (2 * (float)M_PI) / (float)S;
  }
  for (const DistIdx &e : *sup) {
    if (e.d2 <= margin_distance2) continue;
    margin_point_found = true;
    const float *nrm = normals + (size_t)e.idx * 4;
    const float normal_cos = dot3f(z, nrm);
    if (normal_cos < min_normal_cos) {
      min_normal_index = e.idx;
      min_normal_cos = normal_cos;
    }
    if (bp.find_holes) {
      float ind[3];
      directed_orthogonal_axis(z, c, surf + (size_t)e.idx * sstride, ind);
      float angle = angle_between_unit(x_axis, ind, z);
      if (g_variant & V_BOARD_ANGLE_UP) angle *= 1.0f + 4.76837158e-7f;
      if (g_variant & V_BOARD_ANGLE_DOWN) angle *= 1.0f - 4.76837158e-7f;
      const int b = std::min((int)std::floor(angle / max_boundary_angle), S - 1);
      if (b >= 0) { /* a NaN angle (neighbour on the axis) indexes out of range in PCL; skipped here */
        check[b] = true;
        if (angle < min_angle[b]) {
          min_angle[b] = angle;
          min_angle_normal[b] = normal_cos;
        }
        if (angle > max_angle[b]) {
          max_angle[b] = angle;
          max_angle_normal[b] = normal_cos;
        }
      }
    }
  }
  auto finish = [&](void) {
    cross3f(z, x_axis, y_axis);
    for (int a = 0; a < 3; ++a) {
      out[a] = x_axis[a];
      out[3 + a] = y_axis[a];
    }
    return true;
  };
  auto toward_min_normal = [&]() {
    if (min_normal_index == -1) return false;
    directed_orthogonal_axis(z, c, surf + (size_t)min_normal_index * sstride, x_axis);
    return true;
  };
  if (!margin_point_found) {
    for (const DistIdx &e : *sup) {
      if (e.d2 > margin_distance2) continue;
      const float normal_cos = dot3f(z, normals + (size_t)e.idx * 4);
      if (normal_cos < min_normal_cos) {
        min_normal_index = e.idx;
        min_normal_cos = normal_cos;
      }
    }
    if (!toward_min_normal()) return set_nan();
    return finish();
  }
  bool hole_present = false;
  if (bp.find_holes)
    for (int i = 0; i < S; ++i)
      if (!check[i]) {
        hole_present = true;
        break;
      }
  if (!bp.find_holes || !hole_present) {
    if (!toward_min_normal()) return set_nan();
    return finish();
  }
  /* at least one empty sector: the widest plausible hole gives the x direction */
  float angle = 0.f;
  int first_no_border = -1;
  if (check[S - 1]) {
    first_no_border = 0;
  } else {
    for (int i = 0; i < S; ++i)
      if (check[i]) {
        first_no_border = i;
        break;
      }
  }
  float max_hole_prob = -FLT_MAX;
  if (first_no_border >= 0) /* all sectors empty cannot happen here (margin_point_found), PCL would loop forever */
    for (int ch = first_no_border; ch < S; ++ch) {
      if (check[ch]) continue;
      const int hole_first = ch;
      int hole_end = hole_first + 1;
      while (!check[hole_end % S]) ++hole_end;
      if (hole_end - hole_first > 0) {
        const int previous_hole = (((hole_first - 1) < 0) ? (hole_first - 1) + S : (hole_first - 1)) % S;
        const int following_hole = hole_end % S;
        float normal_begin = max_angle_normal[previous_hole];
        float normal_end = min_angle_normal[following_hole];
        normal_begin -= min_normal_cos;
        normal_end -= min_normal_cos;
        normal_begin = normal_begin / (1.0f - min_normal_cos);
        normal_end = normal_end / (1.0f - min_normal_cos);
        normal_begin = 1.0f - normal_begin;
        normal_end = 1.0f - normal_end;
        float hole_width;
        if (following_hole < previous_hole)
          hole_width = min_angle[following_hole] + 2 * (float)M_PI - max_angle[previous_hole];
        else
          hole_width = min_angle[following_hole] - max_angle[previous_hole];
        const float hole_prob = hole_width / (2 * (float)M_PI);
        const float steep_prob = (normal_end + normal_begin) / 2.0f;
        if (hole_prob > bp.hole_size_prob_thresh && steep_prob > bp.steep_thresh && hole_prob > max_hole_prob) {
          max_hole_prob = hole_prob;
          const float angle_weight = ((normal_end - normal_begin) + 1.0f) / 2.0f;
          if (following_hole < previous_hole)
            angle = max_angle[previous_hole] +
                    (min_angle[following_hole] + 2 * (float)M_PI - max_angle[previous_hole]) * angle_weight;
          else
            angle = max_angle[previous_hole] + (min_angle[following_hole] - max_angle[previous_hole]) * angle_weight;
        }
      }
      if (hole_end >= S) break;
      ch = hole_end - 1;
    }
  if (max_hole_prob > -FLT_MAX) {
    /* x_axis = Eigen::AngleAxisf(angle, z) * x_axis  (AngleAxis::toRotationMatrix, then matrix * vector) */
    const float sn = std::sin(angle), cs = std::cos(angle);
    const float sa[3] = {sn * z[0], sn * z[1], sn * z[2]};
    const float ca[3] = {(1.f - cs) * z[0], (1.f - cs) * z[1], (1.f - cs) * z[2]};
    float R[9];
    float tmp = ca[0] * z[1];
    R[1] = tmp - sa[2];
    R[3] = tmp + sa[2];
    tmp = ca[0] * z[2];
    R[2] = tmp + sa[1];
    R[6] = tmp - sa[1];
    tmp = ca[1] * z[2];
    R[5] = tmp - sa[0];
    R[7] = tmp + sa[0];
    R[0] = ca[0] * z[0] + cs;
    R[4] = ca[1] * z[1] + cs;
    R[8] = ca[2] * z[2] + cs;
    float nx[3];
    for (int r = 0; r < 3; ++r) {
      float v = R[r * 3 + 0] * x_axis[0];
      v += R[r * 3 + 1] * x_axis[1];
      v += R[r * 3 + 2] * x_axis[2];
      nx[r] = v;
    }
    for (int a = 0; a < 3; ++a) x_axis[a] = nx[a];
  } else {
    if (!toward_min_normal()) return set_nan();
  }
  return finish();
}
}  // namespace

/* pcl::BOARDLocalReferenceFrameEstimation::compute (SHOT.cpp:441-453: setFindHoles(true), setRadiusSearch(rf_rad_),
 * keypoints as input, full cloud as search surface, the cloud's normals).  The keypoints are processed serially in
 * PCL, each one with at least 6 support points drawing two rand() values: the stream here is glibc's generator
 * seeded with *rand_seed_state == srand(seed); rand_skip values already consumed are skipped first.  Returns the
 * number of rand() values consumed (so that a second call can continue the stream like PCL's second compute()). */
int orc_board_lrf(const float *surf, const float *normals, int n, int sstride, const float *kp, int K, int kstride,
                  double radius, const orc_board_params *bp, unsigned rand_seed, int rand_skip, float *out) {
  CpuGrid g;
  g.build(surf, n, sstride, cell_for_radius(radius));
  std::vector<std::vector<DistIdx>> nbs((size_t)K);
#pragma omp parallel for schedule(dynamic, 32)
  for (int i = 0; i < K; ++i) g.radius(kp + (size_t)i * kstride, radius, nbs[(size_t)i]);
  GlibcRand rng(rand_seed);
  for (int i = 0; i < rand_skip; ++i) rng.next();
  /* "if (tangent_radius_ != 0.0f && search_parameter_ != tangent_radius_)": second search for the x axis */
  const bool second = bp->tangent_radius != 0.0f && radius != (double)bp->tangent_radius;
  std::vector<std::vector<DistIdx>> nbt(second ? (size_t)K : 0);
  CpuGrid gt;
  if (second) {
    gt.build(surf, n, sstride, cell_for_radius(bp->tangent_radius));
#pragma omp parallel for schedule(dynamic, 32)
    for (int i = 0; i < K; ++i) gt.radius(kp + (size_t)i * kstride, (double)bp->tangent_radius, nbt[(size_t)i]);
  }
  std::vector<int> rnd((size_t)K * 2, 0);
  int consumed = 0;
  for (int i = 0; i < K; ++i)
    if (bp->find_holes && nbs[(size_t)i].size() >= 6) {
      rnd[2 * (size_t)i] = rng.next();
      rnd[2 * (size_t)i + 1] = rng.next();
      consumed += 2;
    }
  orc_board_params p = *bp;
  if (p.check_margin_array_size < 1) p.check_margin_array_size = 1;
  if (p.check_margin_array_size > BOARD_MAX_SECTORS) p.check_margin_array_size = BOARD_MAX_SECTORS;
#pragma omp parallel for schedule(dynamic, 32)
  for (int i = 0; i < K; ++i)
    board_point_lrf(surf, sstride, normals, kp + (size_t)i * kstride, nbs[(size_t)i],
                    second ? nbt[(size_t)i] : nbs[(size_t)i], p, &rnd[2 * (size_t)i], out + (size_t)i * 9);
  return consumed;
}

int orc_glibc_rand_nth(unsigned seed, int nth) {
  GlibcRand rng(seed);
  int v = 0;
  for (int i = 0; i < nth; ++i) v = rng.next();
  return v;
}

/* pcl::removeNaNFromPointCloud (common/impl/filter.hpp; SHOT.cpp:298-299): rows with finite x, y, z, in order. */
int orc_remove_nan(const float *xyz, int n, int stride, float *out_xyz, int *out_index) {
  int k = 0;
  for (int i = 0; i < n; ++i) {
    const float *p = xyz + (size_t)i * stride;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    for (int a = 0; a < 3; ++a) out_xyz[3 * (size_t)k + a] = p[a];
    if (out_index) out_index[k] = i;
    ++k;
  }
  return k;
}

/* pcl::transformPointCloud(in, out, Matrix4f) (common/impl/transforms.hpp): x' = ((t00 x + t01 y) + t02 z) + t03 in
 * float32, row by row; in a non-dense cloud rows with a non-finite coordinate are left as they are. */
void orc_transform_points(const float *xyz, int n, int stride, const float *T, float *out_xyz) {
  for (int i = 0; i < n; ++i) {
    const float *p = xyz + (size_t)i * stride;
    float *o = out_xyz + 3 * (size_t)i;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) {
      o[0] = p[0], o[1] = p[1], o[2] = p[2];
      continue;
    }
    for (int r = 0; r < 3; ++r) {
      float v = T[r * 4 + 0] * p[0];
      v += T[r * 4 + 1] * p[1];
      v += T[r * 4 + 2] * p[2];
      v += T[r * 4 + 3];
      o[r] = v;
    }
  }
}

void orc_umeyama3(const double *src, const double *dst, int n, double T16[16]) { umeyama3(src, dst, n, T16); }
uint32_t orc_mt19937_nth(uint32_t seed, int nth) {
  std::mt19937 rng(seed);
  uint32_t v = 0;
  for (int i = 0; i < nth; ++i) v = rng();
  return v;
}

} /* extern "C" */
