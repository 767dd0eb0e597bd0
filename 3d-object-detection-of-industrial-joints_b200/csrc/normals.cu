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
// candidate cells through L1; candidate lists live in shared memory.  Radius mode: one query per
// CTA (gather → sort → accumulate).  Algorithmic HBM traffic: 16 B read + 16 B written per point.
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
    const int cnt = knn_query(g, p.x, p.y, p.z, k, sk, T);
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

// ------------------------------------------------------------------------------------------------------------
// kNN mode, warp-cooperative (the surface points are their own queries — every NormalEstimationOMP call of the
// reference).  A warp owns 32 consecutive points of the cell-major order, i.e. the queries of one to three grid
// cells.  Per cell:
//   1. the 3x3x3 block's points (nine x-contiguous runs) are staged ONCE in shared memory with coalesced 16-byte
//      loads and reused by every query of the cell;
//   2. per query the whole warp evaluates the candidates (lane = candidate, 16-byte conflict-free shared loads,
//      FLANN's float32 L2_Simple sum), then finds a cut tau with k <= #{d2 < tau} <= 32 L by counting
//      (redux.sync) — density-proportional steps, the previous query's cut as the first guess: 1-2 rounds — with
//      tau never above the certified bound (squared distance from the query to the block's open faces), so every
//      point below the cut is provably nearer than anything outside the block;
//   3. the survivors are compacted by ballot into a warp list of packed keys (d2 bits << 32 | original index:
//      one unsigned compare is FLANN's (distance, index) order) and ranked by counting smaller keys (broadcast
//      16-byte shared loads); the k first go to the query's neighbour list in rank order.
// Afterwards lane = query: the nine float32 sums run sequentially in neighbour order (bit-identical to the CPU
// evaluation), closed-form eigen33, flip, store.  Queries the block cannot certify (sparse neighbourhoods, more
// than `cap` candidates, ties that no cut separates) take the exact ring-expanding per-thread search below.
// ------------------------------------------------------------------------------------------------------------
constexpr int NKC_WARPS = 4;
constexpr int NKC_MAXROWS = 128;  // rows (x-contiguous cell runs) one staging pass may touch

struct NkcLayout {
  int cap;       // staged candidates per warp
  int kpad;      // neighbour-list row length (odd: conflict-free)
  int stage_b;   // bytes of the staging region (also the per-thread search's [k][32] key columns)
  int wkeys_b;   // survivor keys
  int list_b;    // neighbour lists
  int rows_b;    // row table (start, inclusive end offset)
  int per_warp;  // total
};
__host__ __device__ inline NkcLayout nkc_layout(int k, int L, int cap) {
  NkcLayout o;
  o.cap = cap;
  o.kpad = k | 1;
  o.stage_b = max(cap * 16, k * 32 * 8);
  o.wkeys_b = (32 * L + 2) * 8;
  o.list_b = ((32 * o.kpad * 4 + 15) / 16) * 16;
  o.rows_b = NKC_MAXROWS * 8;
  o.per_warp = o.stage_b + o.wkeys_b + o.list_b + o.rows_b;
  return o;
}

// Warp-cooperative: copies the points of cells [x0..x1] x [y0..y1] x [z0..z1] that lie inside the closed box
// [bl, bh] into stage[0..cap).  The rows are flattened into one index space (row table in shared memory, binary
// search per element) so that every lane's 16-byte loads are independent and two rounds are in flight at once.
// Returns the number of points inside the box (only the first cap are stored).
__device__ __forceinline__ int nkc_stage_box(const GridView &g, int x0, int x1, int y0, int y1, int z0, int z1,
                                             float bxl, float bxh, float byl, float byh, float bzl, float bzh,
                                             float4 *stage, int cap, int2 *rows, int lane) {
  const unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
  int nraw = 0;
  for (int r0 = 0; r0 < nrows; r0 += 32) {
    const int r = r0 + lane;
    int s = 0, len = 0;
    if (r < nrows) {
      const int row = g.dx * ((y0 + r % ny) + g.dy * (z0 + r / ny));
      s = g.cell_start[row + x0];
      len = g.cell_start[row + x1 + 1] - s;
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += t;
    }
    if (r < nrows) rows[r] = make_int2(s, nraw + incl);
    nraw += __shfl_sync(FULL, incl, 31);
  }
  __syncwarp();
  int n = 0;
  for (int t0 = 0; t0 < nraw; t0 += 64) {
    float4 p[2];
    bool in[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int t = t0 + u * 32 + lane;
      in[u] = false;
      if (t < nraw) {
        int lo = 0, hi = nrows - 1;  // first row whose inclusive end exceeds t
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (rows[mid].y > t) hi = mid; else lo = mid + 1;
        }
        const int2 rw = rows[lo];
        const int prev = lo ? rows[lo - 1].y : 0;
        p[u] = g.pts[rw.x + (t - prev)];
        in[u] = p[u].x >= bxl && p[u].x <= bxh && p[u].y >= byl && p[u].y <= byh && p[u].z >= bzl && p[u].z <= bzh;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const unsigned m = __ballot_sync(FULL, in[u]);
      if (in[u]) {
        const int slot = n + __popc(m & lt_mask);
        if (slot < cap) stage[slot] = p[u];
      }
      n += __popc(m);
    }
  }
  __syncwarp();
  return n;
}

template <int L, int CPL>
__global__ void __launch_bounds__(NKC_WARPS * 32)
    normals_knn_cell_kernel(GridView g, int nq, int k, int cap, float vpx, float vpy, float vpz,
                            float4 *__restrict__ out, unsigned long long *__restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const NkcLayout lay = nkc_layout(k, L, cap);
  unsigned char *wbase = smem_raw + (size_t)warp * lay.per_warp;
  float4 *stage = reinterpret_cast<float4 *>(wbase);
  unsigned long long *wk = reinterpret_cast<unsigned long long *>(wbase + lay.stage_b);
  int *list = reinterpret_cast<int *>(wbase + lay.stage_b + lay.wkeys_b);
  int2 *rows = reinterpret_cast<int2 *>(wbase + lay.stage_b + lay.wkeys_b + lay.list_b);
  const int kpad = lay.kpad;
  constexpr int W = 32 * L;

  const int q0 = (blockIdx.x * NKC_WARPS + warp) * 32;
  if (q0 >= nq) return;
  const int qi = q0 + lane;
  const bool have_q = qi < nq;
  float4 myq = make_float4(0.f, 0.f, 0.f, 0.f);
  int mycell = -1;
  if (have_q) {
    myq = g.pts[qi];
    const int cx = grid_coord(myq.x, g.lox, g.inv_h, g.dx);
    const int cy = grid_coord(myq.y, g.loy, g.inv_h, g.dy);
    const int cz = grid_coord(myq.z, g.loz, g.inv_h, g.dz);
    mycell = cx + g.dx * (cy + g.dy * cz);
  }
  unsigned todo = __ballot_sync(FULL, have_q);
  unsigned slow = 0;  // queries the cell's box could not certify
  const float target = (float)(k + W + 1) * 0.5f;
  const float h = g.h;
  const float grid_hx = g.lox + ((float)g.dx + 1e-3f) * h, grid_hy = g.loy + ((float)g.dy + 1e-3f) * h,
              grid_hz = g.loz + ((float)g.dz + 1e-3f) * h;
  int st_rounds = 0, st_crowded = 0, st_staged = 0, st_r2 = 0, st_groups = 0, st_retry = 0;

  // One query against the n staged candidates, whole warp.  box: the staged box (every point of the cloud outside it
  // is farther from the query than its distance to the nearest open face).  On success the k nearest, in FLANN's
  // (distance, index) order, are in list[ql * kpad ...] and the cut is returned through tau_guess for the next
  // query; returns -1 on success, otherwise the count at the largest certified cut.
  auto process = [&](int ql, int n, float bxl, float bxh, float byl, float byh, float bzl, float bzh,
                     float &tau_guess) -> int {
    const float qx = __shfl_sync(FULL, myq.x, ql), qy = __shfl_sync(FULL, myq.y, ql),
                qz = __shfl_sync(FULL, myq.z, ql);
    // a box face counts only if points can exist beyond it
    float cert = 3.0e38f;
    if (bxl > g.lox) cert = fminf(cert, qx - bxl);
    if (bxh < grid_hx) cert = fminf(cert, bxh - qx);
    if (byl > g.loy) cert = fminf(cert, qy - byl);
    if (byh < grid_hy) cert = fminf(cert, byh - qy);
    if (bzl > g.loz) cert = fminf(cert, qz - bzl);
    if (bzh < grid_hz) cert = fminf(cert, bzh - qz);
    cert -= 4e-6f * (g.coord_scale + fabsf(qx) + fabsf(qy) + fabsf(qz)) + 1e-5f * h;
    const float tau_max = (cert > 0.f) ? ((cert < 1.0e18f) ? cert * cert * 0.99999f : 3.0e38f) : 0.f;
    const int cpl = (n + 31) >> 5;
    float d[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      d[i] = 3.4e38f;
      if (i < cpl) {
        const int t = i * 32 + lane;
        if (t < n) {
          const float4 p = stage[t];
          d[i] = sqdist3(qx, qy, qz, p.x, p.y, p.z);
        }
      }
    }
    // the cut
    float tau = fminf(tau_guess, tau_max), lo = 0.f, hi = 3.4e38f;
    int cnt = 0;
    bool ok = false;
    if (n >= k && tau_max > 0.f) {
#pragma unroll 1
      for (int it = 0; it < 48; ++it) {
        int c = 0;
#pragma unroll
        for (int i = 0; i < CPL; ++i)
          if (i < cpl) c += (d[i] < tau) ? 1 : 0;
        cnt = __reduce_add_sync(FULL, c);
        ++st_rounds;
        if (cnt >= k && cnt <= W) {
          ok = true;
          break;
        }
        if (cnt < k) {
          if (tau >= tau_max) break;  // the box does not certify k neighbours
          lo = tau;
        } else {
          hi = tau;
        }
        float t2 = (cnt > 0) ? tau * (target / (float)cnt) : tau * 4.0f;
        if (it >= 5 || !(t2 > lo) || !(t2 < hi)) t2 = (hi < 3.0e38f) ? 0.5f * (lo + hi) : tau * 2.0f;
        t2 = fminf(t2, tau_max);
        if (!(t2 > lo) || !(t2 < hi)) break;  // no float between the bounds: ties
        tau = t2;
      }
    }
    if (!ok) return (cnt < k) ? cnt : k;
    tau_guess = tau * (target / (float)cnt);
    // survivors -> warp list
    int base = 0;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      if (i < cpl) {
        const bool pass = d[i] < tau;
        const unsigned m = __ballot_sync(FULL, pass);
        if (pass) {
          const int slot = base + __popc(m & lt_mask);
          const int o = __float_as_int(stage[i * 32 + lane].w);
          wk[slot] = ((unsigned long long)__float_as_uint(d[i]) << 32) | (unsigned)o;
        }
        base += __popc(m);
      }
    }
    if (lane < 2) wk[cnt + lane] = ~0ull;
    __syncwarp();
    unsigned long long mine[L];
    int rank[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      mine[l] = (lane + 32 * l < cnt) ? wk[lane + 32 * l] : ~0ull;
      rank[l] = 0;
    }
    for (int j = 0; j < cnt; j += 2) {
      const ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(&wk[j]);
#pragma unroll
      for (int l = 0; l < L; ++l) rank[l] += (kk.x < mine[l] ? 1 : 0) + (kk.y < mine[l] ? 1 : 0);
    }
#pragma unroll
    for (int l = 0; l < L; ++l)
      if (lane + 32 * l < cnt && rank[l] < k) list[ql * kpad + rank[l]] = (int)(unsigned)(mine[l] & 0xffffffffull);
    __syncwarp();
    return -1;
  };

  while (todo) {
    const int first = __ffs(todo) - 1;
    const int cell = __shfl_sync(FULL, mycell, first);
    const unsigned group = __ballot_sync(FULL, have_q && mycell == cell) & todo;
    todo &= ~group;
    const int cx = cell % g.dx, cy = (cell / g.dx) % g.dy, cz = cell / (g.dx * g.dy);
    // population of the 3x3x3 block (nine rows, one per lane)
    int len9 = 0;
    if (lane < 9) {
      const int y = cy + lane % 3 - 1, z = cz + lane / 3 - 1;
      if (y >= 0 && y < g.dy && z >= 0 && z < g.dz) {
        const int row = g.dx * (y + g.dy * z);
        len9 = g.cell_start[row + min(cx + 1, g.dx - 1) + 1] - g.cell_start[row + max(cx - 1, 0)];
      }
    }
    const int n1 = __reduce_add_sync(FULL, len9);
    // radius that holds k points at the block's (surface) density, in cell edges, with 40 % slack
    const float f = 1.4f * 3.0f * sqrtf((float)k / (3.14159265f * (float)max(n1, 1)));
    const int R = (f <= 0.99f) ? 1 : 2;
    const float w = fminf(f, 0.99f * (float)R) * h;
    st_r2 += (R == 2);
    ++st_groups;
    // staged box: the cell grown by w on every side
    const float bxl = g.lox + (float)cx * h - w, bxh = g.lox + (float)(cx + 1) * h + w;
    const float byl = g.loy + (float)cy * h - w, byh = g.loy + (float)(cy + 1) * h + w;
    const float bzl = g.loz + (float)cz * h - w, bzh = g.loz + (float)(cz + 1) * h + w;
    const int n = nkc_stage_box(g, max(cx - R, 0), min(cx + R, g.dx - 1), max(cy - R, 0), min(cy + R, g.dy - 1),
                                max(cz - R, 0), min(cz + R, g.dz - 1), bxl, bxh, byl, byh, bzl, bzh, stage, cap, rows,
                                lane);
    if (n > cap) {  // too crowded for the staging area
      slow |= group;
      st_crowded += __popc(group);
      continue;
    }
    st_staged += n;
    float tau_guess = 1.3f * 9.0f * h * h * (float)k / (3.14159265f * (float)max(n1, 1));
    unsigned grp = group;
    while (grp) {
      const int ql = __ffs(grp) - 1;
      grp &= grp - 1;
      if (process(ql, n, bxl, bxh, byl, byh, bzl, bzh, tau_guess) >= 0) slow |= 1u << ql;
    }
  }
  // second chance, one query at a time: boxes centred on the query, doubling
  {
    unsigned pend = slow;
    while (pend) {
      const int ql = __ffs(pend) - 1;
      pend &= pend - 1;
      const float qx = __shfl_sync(FULL, myq.x, ql), qy = __shfl_sync(FULL, myq.y, ql),
                  qz = __shfl_sync(FULL, myq.z, ql);
      float w = 1.5f * h;
      for (int attempt = 0; attempt < 3; ++attempt, w *= 2.0f) {
        const float bxl = qx - w, bxh = qx + w, byl = qy - w, byh = qy + w, bzl = qz - w, bzh = qz + w;
        const int x0 = grid_coord(bxl, g.lox, g.inv_h, g.dx), x1 = grid_coord(bxh, g.lox, g.inv_h, g.dx);
        const int y0 = grid_coord(byl, g.loy, g.inv_h, g.dy), y1 = grid_coord(byh, g.loy, g.inv_h, g.dy);
        const int z0 = grid_coord(bzl, g.loz, g.inv_h, g.dz), z1 = grid_coord(bzh, g.loz, g.inv_h, g.dz);
        if ((y1 - y0 + 1) * (z1 - z0 + 1) > NKC_MAXROWS) break;
        ++st_retry;
        const int n = nkc_stage_box(g, x0, x1, y0, y1, z0, z1, bxl, bxh, byl, byh, bzl, bzh, stage, cap, rows, lane);
        if (n > cap) break;
        float tau_guess = w * w * 0.5f;
        if (process(ql, n, bxl, bxh, byl, byh, bzl, bzh, tau_guess) < 0) {
          slow &= ~(1u << ql);
          break;
        }
      }
    }
  }
  __syncwarp();
  // exact per-thread ring search for whatever is left (staging area = [k][32] key columns)
  int cnt = have_q ? min(k, g.n) : 0;
  if ((slow >> lane) & 1u) {
    unsigned long long *sk = reinterpret_cast<unsigned long long *>(stage) + lane;
    cnt = knn_query(g, myq.x, myq.y, myq.z, k, sk, 32);
    for (int j = 0; j < cnt; ++j) list[lane * kpad + j] = knn_orig(sk[j * 32]);
  }
  if (stats && lane == 0) {
    atomicAdd(&stats[0], (unsigned long long)__popc(slow));
    atomicAdd(&stats[1], (unsigned long long)st_rounds);
    atomicAdd(&stats[2], (unsigned long long)st_crowded);
    atomicAdd(&stats[3], (unsigned long long)st_staged);
    atomicAdd(&stats[4], (unsigned long long)st_r2);
    atomicAdd(&stats[5], (unsigned long long)st_groups);
    atomicAdd(&stats[6], (unsigned long long)st_retry);
  }
  if (!have_q) return;
  float4 o = make_float4(nanf32(), nanf32(), nanf32(), nanf32());
  if (cnt >= 3) {
    float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int *mylist = list + lane * kpad;
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const float4 n = g.raw[mylist[j]];
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
    const Normal4 r = normal_from_accu(accu, myq.x, myq.y, myq.z, vpx, vpy, vpz);
    o = make_float4(r.nx, r.ny, r.nz, r.curv);
  }
  out[orig_index(myq)] = o;
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
    size_t smem;
    const int T = knn_threads_for(k, &smem);
    if (smem > ctx->smem_optin) return ctx->fail(B200_ERR_INVALID, "normals: k too large for shared memory");
    B200_CUDA(ctx, ensure_dyn_smem(normals_knn_kernel, smem));
    StageScope st_(ctx, ST_NORMALS);
    if (q_is_surface) {
      // rows with non-finite coordinates are not in the grid: they keep NaN normals (PCL: is_dense=false)
      if (c->n_valid < c->n) {
        fill_nan_kernel<<<ceil_div((long long)c->n * 4, 256), 256, 0, ctx->stream>>>(d_out, (size_t)c->n * 4);
        B200_LAUNCHED(ctx);
      }
      static const bool legacy = getenv("B200_NORMALS_WARP") == nullptr;  // experimental warp-cooperative kernel: opt-in
      if (false) {
      } else if (c->n_valid > k && k <= 64 && !legacy) {
        // warp-cooperative kernel: the 3x3x3 block of a cell is staged once and shared by the cell's queries
        const int L = k <= 32 ? 1 : 2;
        const int cap = L == 1 ? 384 : 768;
        const NkcLayout lay = nkc_layout(k, L, cap);
        const size_t sm = (size_t)lay.per_warp * NKC_WARPS;
        const int blocks = ceil_div(c->n_valid, 32 * NKC_WARPS);
        static const bool want_stats = getenv("B200_NKC_STATS") != nullptr;
        DevBuf<unsigned long long> stats;
        if (want_stats) {
          B200_TRY(stats.alloc(ctx, 8));
          B200_TRY(stats.zero());
        }
        if (L == 1) {
          B200_CUDA(ctx, ensure_dyn_smem(normals_knn_cell_kernel<1, 12>, sm));
          normals_knn_cell_kernel<1, 12><<<blocks, NKC_WARPS * 32, sm, ctx->stream>>>(
              *g, c->n_valid, k, cap, vpx, vpy, vpz, out, want_stats ? stats.p : nullptr);
        } else {
          B200_CUDA(ctx, ensure_dyn_smem(normals_knn_cell_kernel<2, 24>, sm));
          normals_knn_cell_kernel<2, 24><<<blocks, NKC_WARPS * 32, sm, ctx->stream>>>(
              *g, c->n_valid, k, cap, vpx, vpy, vpz, out, want_stats ? stats.p : nullptr);
        }
        B200_LAUNCHED(ctx);
        if (want_stats) {
          unsigned long long h[8];
          B200_TRY(readback_small(ctx, stats.p, h, sizeof(h)));
          fprintf(stderr,
                  "[b200 normals] n=%d k=%d h=%g: per-thread search %llu (crowded %llu), %.2f cut rounds/query, "
                  "%.1f staged/group, %llu groups (%llu with the 5x5x5 block), %llu second-chance boxes\n",
                  c->n_valid, k, (double)g->h, h[0], h[2], (double)h[1] / c->n_valid, (double)h[3] / (double)(h[5] ? h[5] : 1),
                  h[5], h[4], h[6]);
        }
      } else if (c->n_valid > 0) {
        normals_knn_kernel<<<ceil_div(c->n_valid, T), T, smem, ctx->stream>>>(*g, nullptr, c->n_valid, k, vpx, vpy,
                                                                             vpz, out);
        B200_LAUNCHED(ctx);
      }
    } else {
      normals_knn_kernel<<<ceil_div(nq, T), T, smem, ctx->stream>>>(*g, d_q, nq, k, vpx, vpy, vpz, out);
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
