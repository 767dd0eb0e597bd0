// keypoints.cu — keypoint extraction on the device: pcl::UniformSampling and pcl::VoxelGrid.
//
// the step between normals and descriptors in every reference program
// (UniformSampling: SHOT.cpp:314-323, SHOT_demo.cpp:246-249, CAD_desc.cpp:295-304, 6Dpose.cpp:281-284;
// VoxelGrid: SHOT_demo.cpp:413-417 / 489-491, FPFH_demo.cpp:412-415 / 494).
//
// Both filters bin the cloud on a regular lattice: ijk = floor(p * (1 / leaf)) per axis (float32, as
// PCL computes it), leaf index = (ijk - min_ijk) . (1, dx, dx*dy).
//   UniformSampling keeps, per occupied leaf, the INPUT point minimising ||p - (float)ijk||^2 + 1 (PCL
//   compares against the integer index vector, not the metric centre; the homogeneous coordinate adds 1);
//   the first point wins ties.  Here: one 64-bit atomicMin per point on (distance bits, row index).
//   VoxelGrid replaces the points of a leaf by their centroid.  PCL accumulates in float32 in the order
//   of an unstable std::sort, i.e. the last bits of its centroid are not defined by the input; here the
//   sums are float64 atomics (order independent to ~1e-16) and the centroid is rounded once.
// Output order is ascending leaf index (VoxelGrid's order in PCL; UniformSampling's is hash-map order
// there and is defined as ascending here).  Dense leaf arrays, capped at 2^26
// leaves — PCL refuses lattices whose index overflows an int in the same way ("leaf size is too small").
#include <algorithm>

#include <cstring>

#include "common.cuh"

namespace {

constexpr long long KP_MAX_LEAVES = 1ll << 26;

__device__ __forceinline__ int leaf_coord(float v, float inv) { return (int)floorf(v * inv); }

// min / max lattice coordinates of the finite points: box[0..2] = min ijk, box[3..5] = max ijk
__global__ void kp_bounds_kernel(const float *__restrict__ xyz, int n, int stride, float ix, float iy, float iz,
                                 int *__restrict__ box) {
  int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = xyz[(size_t)i * stride], y = xyz[(size_t)i * stride + 1], z = xyz[(size_t)i * stride + 2];
    if (!finite3(x, y, z)) continue;
    const int c[3] = {leaf_coord(x, ix), leaf_coord(y, iy), leaf_coord(z, iz)};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = min(mn[a], c[a]);
      mx[a] = max(mx[a], c[a]);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = __reduce_min_sync(0xffffffffu, mn[a]);
    mx[a] = __reduce_max_sync(0xffffffffu, mx[a]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&box[a], mn[a]);
      atomicMax(&box[3 + a], mx[a]);
    }
  }
}

struct Lattice {
  float ix, iy, iz;
  int mnx, mny, mnz;
  int dx, dy, dz;
};

__device__ __forceinline__ int leaf_index(const Lattice &L, float x, float y, float z, int &i, int &j, int &k) {
  i = leaf_coord(x, L.ix);
  j = leaf_coord(y, L.iy);
  k = leaf_coord(z, L.iz);
  return (i - L.mnx) + L.dx * ((j - L.mny) + L.dy * (k - L.mnz));
}

__global__ void kp_fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// UniformSampling: best[leaf] = min over points of (||p - ijk||^2 + 1 as float bits, row)
__global__ void us_select_kernel(const float *__restrict__ xyz, int n, int stride, Lattice L,
                                 unsigned long long *__restrict__ best) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float x = xyz[(size_t)r * stride], y = xyz[(size_t)r * stride + 1], z = xyz[(size_t)r * stride + 2];
  if (!finite3(x, y, z)) return;
  int i, j, k;
  const int leaf = leaf_index(L, x, y, z, i, j, k);
  const float d0 = x - (float)i, d1 = y - (float)j, d2 = z - (float)k;
  float diff = d0 * d0;
  diff += d1 * d1;
  diff += d2 * d2;
  diff += 1.0f;  // (w = 1) - (w = 0) of the homogeneous vectors
  atomicMin(&best[leaf], ((unsigned long long)__float_as_uint(diff) << 32) | (unsigned)r);
}

__global__ void us_flags_kernel(const unsigned long long *__restrict__ best, int nleaf, int *__restrict__ flags) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nleaf) flags[c] = best[c] != ~0ull;
}

__global__ void us_emit_kernel(const float *__restrict__ xyz, int stride, const unsigned long long *__restrict__ best,
                               const int *__restrict__ flags, const int *__restrict__ slots, int nleaf,
                               float *__restrict__ out_xyz, int *__restrict__ out_index) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nleaf || !flags[c]) return;
  const int r = (int)(unsigned)(best[c] & 0xffffffffull);
  const int s = slots[c];
  out_xyz[(size_t)s * 3 + 0] = xyz[(size_t)r * stride + 0];
  out_xyz[(size_t)s * 3 + 1] = xyz[(size_t)r * stride + 1];
  out_xyz[(size_t)s * 3 + 2] = xyz[(size_t)r * stride + 2];
  if (out_index) out_index[s] = r;
}

// VoxelGrid: float64 sums + counts per leaf
__global__ void vg_accumulate_kernel(const float *__restrict__ xyz, int n, int stride, Lattice L,
                                     double *__restrict__ sums, int *__restrict__ counts) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float x = xyz[(size_t)r * stride], y = xyz[(size_t)r * stride + 1], z = xyz[(size_t)r * stride + 2];
  if (!finite3(x, y, z)) return;
  int i, j, k;
  const int leaf = leaf_index(L, x, y, z, i, j, k);
  atomicAdd(&sums[(size_t)leaf * 3 + 0], (double)x);
  atomicAdd(&sums[(size_t)leaf * 3 + 1], (double)y);
  atomicAdd(&sums[(size_t)leaf * 3 + 2], (double)z);
  atomicAdd(&counts[leaf], 1);
}

__global__ void vg_flags_kernel(const int *__restrict__ counts, int nleaf, int *__restrict__ flags) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nleaf) flags[c] = counts[c] > 0;
}

__global__ void vg_emit_kernel(const double *__restrict__ sums, const int *__restrict__ counts,
                               const int *__restrict__ slots, int nleaf, float *__restrict__ out_xyz) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nleaf || counts[c] <= 0) return;
  const int s = slots[c];
  const double inv = 1.0 / (double)counts[c];
  out_xyz[(size_t)s * 3 + 0] = (float)(sums[(size_t)c * 3 + 0] * inv);
  out_xyz[(size_t)s * 3 + 1] = (float)(sums[(size_t)c * 3 + 1] * inv);
  out_xyz[(size_t)s * 3 + 2] = (float)(sums[(size_t)c * 3 + 2] * inv);
}

// lattice of the finite points (one small readback: the leaf arrays are sized on the host)
int make_lattice(b200_ctx *ctx, const float *d_xyz, int n, int stride, float lx, float ly, float lz, Lattice *L,
                 long long *nleaf) {
  if (!(lx > 0.f) || !(ly > 0.f) || !(lz > 0.f)) return ctx->fail(B200_ERR_INVALID, "keypoints: leaf size must be > 0");
  L->ix = 1.0f / lx;
  L->iy = 1.0f / ly;
  L->iz = 1.0f / lz;
  DevBuf<int> box;
  B200_TRY(box.alloc(ctx, 6));
  const int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
  B200_CUDA(ctx, cudaMemcpyAsync(box.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  kp_bounds_kernel<<<std::min(ceil_div(n, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>(d_xyz, n, stride, L->ix, L->iy,
                                                                                           L->iz, box.p);
  B200_LAUNCHED(ctx);
  int h[6];
  B200_CUDA(ctx, cudaMemcpyAsync(h, box.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, ctx->sync());
  if (h[0] > h[3]) {  // no finite point
    *nleaf = 0;
    return B200_OK;
  }
  L->mnx = h[0], L->mny = h[1], L->mnz = h[2];
  const long long dx = (long long)h[3] - h[0] + 1, dy = (long long)h[4] - h[1] + 1, dz = (long long)h[5] - h[2] + 1;
  if (dx * dy * dz > KP_MAX_LEAVES)
    return ctx->fail(B200_ERR_CAPACITY, "keypoints: leaf size is too small for the input dataset (more than 2^26 leaves)");
  L->dx = (int)dx, L->dy = (int)dy, L->dz = (int)dz;
  *nleaf = dx * dy * dz;
  return B200_OK;
}

}  // namespace

// d_count: device int (number of keypoints written)
// ---- cloud utilities around the hot path: removeNaNFromPointCloud (SHOT.cpp:298-299) and transformPointCloud
// (SHOT.cpp / SHOT_demo.cpp: the model placed by a pose before ICP) ------------------------------------------
namespace {
__global__ void finite_flags_kernel(const float *__restrict__ xyz, int n, int stride, int *__restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *p = xyz + (size_t)i * stride;
  flags[i] = finite3(p[0], p[1], p[2]) ? 1 : 0;
}
__global__ void compact_rows_kernel(const float *__restrict__ xyz, int n, int stride, const int *__restrict__ flags,
                                    const int *__restrict__ slots, float *__restrict__ out_xyz, int *__restrict__ out_index) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  const float *p = xyz + (size_t)i * stride;
  const int s = slots[i];
  out_xyz[3 * (size_t)s + 0] = p[0];
  out_xyz[3 * (size_t)s + 1] = p[1];
  out_xyz[3 * (size_t)s + 2] = p[2];
  if (out_index) out_index[s] = i;
}
struct Mat34 {
  float m[12];
};
// pcl::transformPointCloud (common/impl/transforms.hpp): x' = ((t00 x + t01 y) + t02 z) + t03, row by row; rows with
// a non-finite coordinate are copied unchanged (the non-dense branch)
__global__ void transform_rows_kernel(const float *__restrict__ xyz, int n, int stride, Mat34 T, float *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *p = xyz + (size_t)i * stride;
  const float x = p[0], y = p[1], z = p[2];
  float o[3] = {x, y, z};
  if (finite3(x, y, z)) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float v = T.m[r * 4 + 0] * x;
      v += T.m[r * 4 + 1] * y;
      v += T.m[r * 4 + 2] * z;
      v += T.m[r * 4 + 3];
      o[r] = v;
    }
  }
  out[3 * (size_t)i + 0] = o[0];
  out[3 * (size_t)i + 1] = o[1];
  out[3 * (size_t)i + 2] = o[2];
}
}  // namespace

int dev_remove_nan(b200_ctx *ctx, const float *d_xyz, int n, int stride, float *d_out_xyz, int *d_out_index, int *d_count) {
  B200_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(int), ctx->stream));
  if (n <= 0) return B200_OK;
  DevBuf<int> flags, slots;
  B200_TRY(flags.alloc(ctx, (size_t)n));
  B200_TRY(slots.alloc(ctx, (size_t)n));
  finite_flags_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, stride, flags.p);
  B200_LAUNCHED(ctx);
  B200_TRY(exclusive_scan_i32(ctx, flags.p, slots.p, n, d_count));
  compact_rows_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, stride, flags.p, slots.p, d_out_xyz, d_out_index);
  B200_LAUNCHED(ctx);
  return B200_OK;
}

int dev_transform_points(b200_ctx *ctx, const float *d_xyz, int n, int stride, const float *T16, float *d_out_xyz) {
  if (n <= 0) return B200_OK;
  Mat34 T;
  memcpy(T.m, T16, sizeof(T.m));
  transform_rows_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, stride, T, d_out_xyz);
  B200_LAUNCHED(ctx);
  return B200_OK;
}

int dev_uniform_sampling(b200_ctx *ctx, const float *d_xyz, int n, int stride, float leaf, float *d_out_xyz,
                         int *d_out_index, int *d_count) {
  B200_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(int), ctx->stream));
  if (n <= 0) return B200_OK;
  Lattice L;
  long long nleaf = 0;
  B200_TRY(make_lattice(ctx, d_xyz, n, stride, leaf, leaf, leaf, &L, &nleaf));
  if (nleaf == 0) return B200_OK;
  DevBuf<unsigned long long> best;
  DevBuf<int> flags, slots;
  B200_TRY(best.alloc(ctx, (size_t)nleaf));
  B200_TRY(flags.alloc(ctx, (size_t)nleaf));
  B200_TRY(slots.alloc(ctx, (size_t)nleaf));
  kp_fill_u64_kernel<<<ceil_div(nleaf, 256), 256, 0, ctx->stream>>>(best.p, (size_t)nleaf, ~0ull);
  B200_LAUNCHED(ctx);
  us_select_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, stride, L, best.p);
  B200_LAUNCHED(ctx);
  us_flags_kernel<<<ceil_div(nleaf, 256), 256, 0, ctx->stream>>>(best.p, (int)nleaf, flags.p);
  B200_LAUNCHED(ctx);
  B200_TRY(exclusive_scan_i32(ctx, flags.p, slots.p, (int)nleaf, d_count));
  us_emit_kernel<<<ceil_div(nleaf, 256), 256, 0, ctx->stream>>>(d_xyz, stride, best.p, flags.p, slots.p, (int)nleaf,
                                                               d_out_xyz, d_out_index);
  B200_LAUNCHED(ctx);
  return B200_OK;
}

int dev_voxel_grid(b200_ctx *ctx, const float *d_xyz, int n, int stride, float lx, float ly, float lz,
                   float *d_out_xyz, int *d_count) {
  B200_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(int), ctx->stream));
  if (n <= 0) return B200_OK;
  Lattice L;
  long long nleaf = 0;
  B200_TRY(make_lattice(ctx, d_xyz, n, stride, lx, ly, lz, &L, &nleaf));
  if (nleaf == 0) return B200_OK;
  DevBuf<double> sums;
  DevBuf<int> counts, flags, slots;
  B200_TRY(sums.alloc(ctx, (size_t)nleaf * 3));
  B200_TRY(counts.alloc(ctx, (size_t)nleaf));
  B200_TRY(flags.alloc(ctx, (size_t)nleaf));
  B200_TRY(slots.alloc(ctx, (size_t)nleaf));
  B200_TRY(sums.zero());
  B200_TRY(counts.zero());
  vg_accumulate_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, stride, L, sums.p, counts.p);
  B200_LAUNCHED(ctx);
  vg_flags_kernel<<<ceil_div(nleaf, 256), 256, 0, ctx->stream>>>(counts.p, (int)nleaf, flags.p);
  B200_LAUNCHED(ctx);
  B200_TRY(exclusive_scan_i32(ctx, flags.p, slots.p, (int)nleaf, d_count));
  vg_emit_kernel<<<ceil_div(nleaf, 256), 256, 0, ctx->stream>>>(sums.p, counts.p, slots.p, (int)nleaf, d_out_xyz);
  B200_LAUNCHED(ctx);
  return B200_OK;
}
