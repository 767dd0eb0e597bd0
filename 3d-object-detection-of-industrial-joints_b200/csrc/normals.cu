// normals.cu — per-point surface normals + curvature.
//
// Replaces pcl::NormalEstimationOMP<PointXYZRGBA, Normal>::compute as called at SHOT.cpp:302-308
// (k = 40), 6Dpose.cpp:275-278 (k = 10), SHOT_demo.cpp:405-411 (k = 50), SHOT_scenes.cpp:285-291
// (k = 20), CAD_desc.cpp:283-289 and, with setRadiusSearch, FPFH_demo.cpp:416-420 (r = 0.15) and
// FPFH_scenes_clustered.cpp:273-277 (r = 0.05).
//
// Arithmetic follows PCL 1.8: single-pass float32 mean/covariance with nine accumulators summed in
// neighbour order (nearest first), closed-form eigen33 for the smallest eigenvector, curvature =
// |lambda0 / trace|, flip towards the viewpoint.  The float32 sums are evaluated as the same
// mul/add sequence (the library is built with --fmad=false), so the covariance is bit-identical to
// the CPU evaluation; only the libm calls inside eigen33 (atan2f/cosf/sinf) can differ by an ulp.
//
// kNN mode: one query per thread, queries visited in cell order so a warp's threads share their
// candidate cells through L1; the k best candidates live in shared memory as a max-heap (search.cuh: 1.35 -> 1.17 ms
// against round 1's sorted list).  Radius mode: one query per
// CTA (gather → sort → accumulate).  Algorithmic HBM traffic: 16 B read + 16 B written per point.
//
// Round 2 measured three cooperative rewrites of the kNN kernel against this one on the 1 M-point scene, k = 20
// (all bit-identical to it; numbers in DESIGN.md section 4): a cell's candidate box staged once in shared memory and
// (a) a whole warp per query with a counted cut + rank sort: 1.83 ms, 909 M warp instructions; (b) 8-lane sub-groups
// with a bitonic sort: 2.45 ms, 1 331 M; (c) one query per thread with a density-estimated cut instead of the sorted
// insertion: 2.4-3.5 ms — against 1.35 ms and 893 M here.  The ballot/compaction bookkeeping of the cooperative
// forms costs about one warp instruction per candidate-query pair, more than the per-thread insertion they remove
// (62 % of this kernel's instructions, search.cuh), so the per-thread kernel stays.  A sixth variant keeps the per-thread
// form but replaces the insertion: the K smallest distances in registers by a min/max chain, then a second pass
// that collects and orders the K winners (knn_query_2pass, opt-in): 1.45 ms — with 32 queries per warp some lane
// passes the admission test at almost every candidate, so the 2 K-instruction chain runs about as often as the
// insertion loop did.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "pcl_eigen33.cuh"
#include "search.cuh"

namespace {

struct Normal4 {
  float nx, ny, nz, curv;
};

// accu: the nine sums already divided by the neighbour count.  PCL computePointNormal →
// solvePlaneParameters → eigen33 → flipNormalTowardsViewpoint.
__device__ inline Normal4 normal_from_accu(const float accu[9], float px, float py, float pz, float vpx, float vpy,
                                           float vpz) {
  float cov[9];
  cov[0] = accu[0] - accu[6] * accu[6];
  cov[1] = accu[1] - accu[6] * accu[7];
  cov[2] = accu[2] - accu[6] * accu[8];
  cov[4] = accu[3] - accu[7] * accu[7];
  cov[5] = accu[4] - accu[7] * accu[8];
  cov[8] = accu[5] - accu[8] * accu[8];
  cov[3] = cov[1];
  cov[6] = cov[2];
  cov[7] = cov[5];
  float scale = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(cov[i]));
  if (scale <= 1.17549435e-38f) scale = 1.0f;  // FLT_MIN
  float m[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) m[i] = cov[i] / scale;
  float roots[3];
  compute_roots(m, roots);
  const float eigenvalue = roots[0] * scale;
  m[0] -= roots[0];
  m[4] -= roots[0];
  m[8] -= roots[0];
  float v1[3], v2[3], v3[3];
  cross3(&m[0], &m[3], v1);
  cross3(&m[0], &m[6], v2);
  cross3(&m[3], &m[6], v3);
  const float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  const float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  const float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  float vx, vy, vz, l;
  if (l1 >= l2 && l1 >= l3) {
    vx = v1[0], vy = v1[1], vz = v1[2], l = l1;
  } else if (l2 >= l1 && l2 >= l3) {
    vx = v2[0], vy = v2[1], vz = v2[2], l = l2;
  } else {
    vx = v3[0], vy = v3[1], vz = v3[2], l = l3;
  }
  const float s = sqrtf(l);
  Normal4 r;
  r.nx = vx / s;
  r.ny = vy / s;
  r.nz = vz / s;
  const float eig_sum = cov[0] + cov[4] + cov[8];
  r.curv = (eig_sum != 0.0f) ? fabsf(eigenvalue / eig_sum) : 0.0f;
  const float dx = vpx - px, dy = vpy - py, dz = vpz - pz;
  const float cos_theta = (dx * r.nx + dy * r.ny + dz * r.nz);
  if (cos_theta < 0.0f) {
    r.nx *= -1.0f;
    r.ny *= -1.0f;
    r.nz *= -1.0f;
  }
  return r;
}

__global__ void fill_nan_kernel(float *p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = nanf32();
}

// kNN mode.  q == nullptr: the queries are the indexed surface points themselves, visited in cell
// order (thread i takes g.pts[i]); results are scattered to the original row.
// K2 > 0: the two-pass search of search.cuh for k == K2 (shared memory sized for K2 + KNN2_SLACK keys per thread).
template <int K2>
__global__ void normals_knn_kernel(GridView g, const float4 *__restrict__ q, int nq, int k, float vpx, float vpy,
                                   float vpz, float4 *__restrict__ out) {
  extern __shared__ unsigned char smem_raw[];
  const int T = blockDim.x;
  unsigned long long *sk = reinterpret_cast<unsigned long long *>(smem_raw) + threadIdx.x;
  const int i = blockIdx.x * T + threadIdx.x;
  if (i >= nq) return;
  float4 p;
  int dst;
  if (q) {
    p = q[i];
    dst = i;
  } else {
    p = g.pts[i];
    dst = orig_index(p);
  }
  float4 o = make_float4(nanf32(), nanf32(), nanf32(), nanf32());
  if (finite3(p.x, p.y, p.z)) {
    int cnt;
    if (K2 > 0) {
      cnt = knn_query_2pass<(K2 > 0 ? K2 : 1)>(g, p.x, p.y, p.z, sk, T, K2 + KNN2_SLACK);
      if (cnt < 0) cnt = knn_query(g, p.x, p.y, p.z, k, sk, T);  // more ties at the k-th distance than the list takes
    } else {
      cnt = knn_query(g, p.x, p.y, p.z, k, sk, T);
    }
    if (cnt >= 3) {
      float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < cnt; ++j) {
        const float4 n = g.raw[knn_orig(sk[j * T])];
        accu[0] += n.x * n.x;
        accu[1] += n.x * n.y;
        accu[2] += n.x * n.z;
        accu[3] += n.y * n.y;
        accu[4] += n.y * n.z;
        accu[5] += n.z * n.z;
        accu[6] += n.x;
        accu[7] += n.y;
        accu[8] += n.z;
      }
      const float fn = (float)cnt;
#pragma unroll
      for (int a = 0; a < 9; ++a) accu[a] = accu[a] / fn;
      const Normal4 r = normal_from_accu(accu, p.x, p.y, p.z, vpx, vpy, vpz);
      o = make_float4(r.nx, r.ny, r.nz, r.curv);
    }
  }
  out[dst] = o;
}

// radius mode: one query per CTA.  The sorted neighbour list is walked by nine lanes, one per
// accumulator, so each sum keeps PCL's sequential order.
__global__ void __launch_bounds__(128) normals_radius_kernel(GridView g, const float4 *__restrict__ q, int nq,
                                                             float radius, float r2, int cap,
                                                             unsigned long long *glob_key, int *glob_pos, float vpx,
                                                             float vpy, float vpz, float4 *__restrict__ out) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ int s_count;
  __shared__ float4 s_pts[128];
  unsigned long long *key;
  int *pos;
  if (glob_key) {
    key = glob_key + (size_t)blockIdx.x * cap;
    pos = glob_pos + (size_t)blockIdx.x * cap;
  } else {
    key = reinterpret_cast<unsigned long long *>(smem_raw);
    pos = reinterpret_cast<int *>(smem_raw + (size_t)cap * sizeof(unsigned long long));
  }
  for (int i = blockIdx.x; i < nq; i += gridDim.x) {
    const float4 p = q[i];
    int n = gather_radius(g, p.x, p.y, p.z, radius, r2, key, pos, cap, &s_count);
    if (n > cap) n = cap;
    bitonic_sort(key, pos, n);
    // the nine float32 sums run sequentially in PCL's (d2, index) order, one lane each; the points are staged 128
    // at a time by the whole CTA so that their loads are in flight together instead of one per loop step
    float acc = 0.0f;
    for (int c0 = 0; c0 < n && n >= 3; c0 += 128) {
      const int m = min(128, n - c0);
      if ((int)threadIdx.x < m) s_pts[threadIdx.x] = g.pts[pos[c0 + threadIdx.x]];
      __syncthreads();
      if (threadIdx.x < 9) {
        const int lane = threadIdx.x;
        for (int j = 0; j < m; ++j) {
          const float4 v = s_pts[j];
          float a, b;
          switch (lane) {
            case 0: a = v.x, b = v.x; break;
            case 1: a = v.x, b = v.y; break;
            case 2: a = v.x, b = v.z; break;
            case 3: a = v.y, b = v.y; break;
            case 4: a = v.y, b = v.z; break;
            case 5: a = v.z, b = v.z; break;
            case 6: a = v.x, b = 1.0f; break;
            case 7: a = v.y, b = 1.0f; break;
            default: a = v.z, b = 1.0f; break;
          }
          acc += (lane < 6) ? a * b : a;
        }
      }
      __syncthreads();
    }
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      if (n >= 3 && lane < 9) acc = acc / (float)n;
      float accu[9];
#pragma unroll
      for (int a = 0; a < 9; ++a) accu[a] = __shfl_sync(0xffffffffu, acc, a);
      if (lane == 0) {
        float4 o = make_float4(nanf32(), nanf32(), nanf32(), nanf32());
        if (n >= 3) {
          const Normal4 r = normal_from_accu(accu, p.x, p.y, p.z, vpx, vpy, vpz);
          o = make_float4(r.nx, r.ny, r.nz, r.curv);
        }
        out[i] = o;
      }
    }
    __syncthreads();
  }
}

}  // namespace

int dev_normals(b200_ctx *ctx, b200_cloud *c, const float4 *d_q, int nq, bool q_is_surface, int k, double radius,
                const float *vp, float *d_out) {
  // Feature::initCompute: exactly one of k / radius must be set
  if ((k != 0) == (radius != 0.0) || k < 0 || radius < 0.0)
    return ctx->fail(B200_ERR_INVALID, "normals: exactly one of k / radius must be non-zero");
  if (nq <= 0) return B200_OK;
  const float vpx = vp ? vp[0] : 0.f, vpy = vp ? vp[1] : 0.f, vpz = vp ? vp[2] : 0.f;
  float4 *out = reinterpret_cast<float4 *>(d_out);
  if (k) {
    if (k > 1024) return ctx->fail(B200_ERR_INVALID, "normals: k must be <= 1024");
    const GridView *g;
    B200_TRY(cloud_grid_for_knn(c, k, &g));
    // B200_NORMALS_KNN=2pass: the two-pass search of search.cuh for k = 10 / 20 (bit-identical; measured 1.45 ms
    // against 1.32 ms for the sorted-insertion search on the 1 M-point scene, so it is not the default; the tests
    // compare the two).
    const char *sel = getenv("B200_NORMALS_KNN");
    const bool two_pass = (k == 10 || k == 20) && sel && !strcmp(sel, "2pass");
    size_t smem;
    const int T = knn_threads_for(two_pass ? k + KNN2_SLACK : k, &smem);
    if (smem > ctx->smem_optin) return ctx->fail(B200_ERR_INVALID, "normals: k too large for shared memory");
    auto kern = two_pass ? (k == 10 ? normals_knn_kernel<10> : normals_knn_kernel<20>) : normals_knn_kernel<0>;
    B200_CUDA(ctx, ensure_dyn_smem(kern, smem));
    StageScope st_(ctx, ST_NORMALS);
    if (q_is_surface) {
      // rows with non-finite coordinates are not in the grid: they keep NaN normals (PCL: is_dense=false)
      if (c->n_valid < c->n) {
        fill_nan_kernel<<<ceil_div((long long)c->n * 4, 256), 256, 0, ctx->stream>>>(d_out, (size_t)c->n * 4);
        B200_LAUNCHED(ctx);
      }
      if (c->n_valid > 0) {
        kern<<<ceil_div(c->n_valid, T), T, smem, ctx->stream>>>(*g, nullptr, c->n_valid, k, vpx, vpy, vpz, out);
        B200_LAUNCHED(ctx);
      }
    } else {
      kern<<<ceil_div(nq, T), T, smem, ctx->stream>>>(*g, d_q, nq, k, vpx, vpy, vpz, out);
      B200_LAUNCHED(ctx);
    }
    return B200_OK;
  }
  // radius mode
  const GridView *g;
  B200_TRY(cloud_grid_for_radius(c, radius, &g));
  DevBuf<int> counts;
  DevBuf<unsigned long long> stats;
  B200_TRY(counts.alloc(ctx, (size_t)nq));
  B200_TRY(stats.alloc(ctx, 8));
  B200_TRY(dev_radius_count(ctx, *g, d_q, nq, radius, counts.p, stats.p));
  unsigned long long hstats[2];
  B200_TRY(readback_small(ctx, stats.p, hstats, sizeof(hstats)));
  const int max_count = (int)hstats[0];
  ctx->last_max_nbrs = max_count;
  ctx->last_mean_nbrs = (double)hstats[1] / nq;
  const float r2 = (float)(radius * radius);
  const int cap = next_pow2_host(std::max(max_count, 32));
  const size_t smem = (size_t)cap * 12;
  StageScope st_(ctx, ST_NORMALS);
  if (smem <= 96 * 1024) {
    B200_CUDA(ctx,
              ensure_dyn_smem(normals_radius_kernel, smem));
    const int grid = std::min(nq, ctx->sm_count * 8);
    normals_radius_kernel<<<grid, 128, smem, ctx->stream>>>(*g, d_q, nq, (float)radius, r2, cap, nullptr, nullptr, vpx,
                                                            vpy, vpz, out);
    B200_LAUNCHED(ctx);
  } else {
    const int grid = std::min(nq, ctx->sm_count * 2);
    DevBuf<unsigned long long> gk;
    DevBuf<int> gp;
    B200_TRY(gk.alloc(ctx, (size_t)grid * cap));
    B200_TRY(gp.alloc(ctx, (size_t)grid * cap));
    normals_radius_kernel<<<grid, 128, 0, ctx->stream>>>(*g, d_q, nq, (float)radius, r2, cap, gk.p, gp.p, vpx, vpy,
                                                         vpz, out);
    B200_LAUNCHED(ctx);
  }
  return B200_OK;
}
