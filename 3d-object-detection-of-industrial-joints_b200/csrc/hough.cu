// hough.cu — pcl::Hough3DGrouping on the device (the grouping every reference
// program uses by default — SHOT.cpp:433-470, SHOT_demo.cpp:540-577, FPFH_demo.cpp:548-585 — with the
// settings they make: setUseInterpolation(false), setUseDistanceWeight(true), setHoughBinSize,
// setHoughThreshold, reference frames given through setInputRf / setSceneRf).
//
// pcl 1.8 recognition/impl/cg/hough_3d.hpp:
//   train():        centroid = float32 running sum of the model keypoints / n; model_vote[i] = the offset
//                   (centroid - keypoint_i) expressed in keypoint i's reference frame (float32 dots).
//   houghVoting():  per correspondence, scene_vote = scene_rf^T * model_vote + scene point (float32, widened
//                   to double); the Hough space spans [min, max] of the votes with cubic bins of
//                   hough_bin_size; without interpolation a vote adds its weight to one bin.  The weight is
//                   1 - distance / max_distance, and max_distance is only tracked when interpolation is on
//                   (it stays -FLT_MAX), so every weight is exactly 1.0: a bin's value is its vote count.
//   findMaxima():   every bin with value >= threshold (threshold < 0: that fraction of the maximum) is an
//                   instance, in ascending bin index; its voters (ascending correspondence index) go to
//                   RANSAC (inlier threshold = bin size), exactly as in geometric-consistency grouping.
// Votes whose bin coordinate falls outside the space (a vote exactly on the upper bound) are dropped, as
// HoughSpace3D::vote does.  Correspondences with a non-finite reference frame are skipped.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

constexpr long long HOUGH_MAX_BINS = 1ll << 27;

__device__ __forceinline__ unsigned enc_ordered(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float dec_ordered(unsigned e) {
  unsigned u = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// float32 running sum in row order (Eigen: centroid += p; centroid /= n), one thread per coordinate
__global__ void hough_centroid_kernel(const float4 *__restrict__ kp, int n, float *__restrict__ centroid) {
  const int a = threadIdx.x;
  if (a >= 3) return;
  float s = 0.f;
  for (int i = 0; i < n; ++i) {
    const float4 p = kp[i];
    s += (a == 0) ? p.x : (a == 1 ? p.y : p.z);
  }
  centroid[a] = s / (float)n;
}

__global__ void hough_model_votes_kernel(const float4 *__restrict__ kp, const float *__restrict__ rf, int n,
                                         const float *__restrict__ centroid, float *__restrict__ votes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = kp[i];
  const float dx = centroid[0] - p.x, dy = centroid[1] - p.y, dz = centroid[2] - p.z;
  const float *r = rf + (size_t)i * 9;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float v = r[a * 3 + 0] * dx;
    v += r[a * 3 + 1] * dy;
    v += r[a * 3 + 2] * dz;
    votes[(size_t)i * 3 + a] = v;
  }
}

// scene_votes (float32 values, as PCL computes them before widening) + their bounds
__global__ void hough_scene_votes_kernel(const b200_corr *__restrict__ corrs, int C, const float *__restrict__ mvotes,
                                         const float4 *__restrict__ scene_kp, const float *__restrict__ scene_rf,
                                         float *__restrict__ svotes, unsigned *__restrict__ box) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  const b200_corr c = corrs[i];
  const float *mv = mvotes + (size_t)c.index_query * 3;
  const float *r = scene_rf + (size_t)c.index_match * 9;
  const float4 sp = scene_kp[c.index_match];
  float v[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float t = r[0 * 3 + a] * mv[0];  // x axis component a
    t += r[1 * 3 + a] * mv[1];
    t += r[2 * 3 + a] * mv[2];
    t += (a == 0) ? sp.x : (a == 1 ? sp.y : sp.z);
    v[a] = t;
    svotes[(size_t)i * 3 + a] = t;
  }
  if (isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2])) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&box[a], enc_ordered(v[a]));
      atomicMax(&box[3 + a], enc_ordered(v[a]));
    }
  }
}

struct HoughSpace {
  double mn[3], bin;
  int cnt[3];
};

__device__ __forceinline__ long long hough_bin(const HoughSpace &H, const float *v) {
  if (!(isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]))) return -1;
  long long index = 0, mul = 1;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int ci = (int)floor(((double)v[a] - H.mn[a]) / H.bin);
    if (ci < 0 || ci >= H.cnt[a]) return -1;
    index += mul * ci;
    mul *= H.cnt[a];
  }
  return index;
}

__global__ void hough_vote_kernel(const float *__restrict__ svotes, int C, HoughSpace H, int *__restrict__ bins,
                                  int *__restrict__ vmax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  const long long b = hough_bin(H, svotes + (size_t)i * 3);
  if (b < 0) return;
  const int v = atomicAdd(&bins[b], 1) + 1;
  atomicMax(vmax, v);
}

// flags[b] = bin is a maximum; sizes[b] = its vote count (0 otherwise)
__global__ void hough_flags_kernel(const int *__restrict__ bins, int nbins, double threshold, const int *__restrict__ vmax,
                                   int *__restrict__ flags, int *__restrict__ sizes) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  double thr = threshold;
  if (thr < 0) {  // HoughSpace3D::findMaxima: relative to the largest bin
    const double hmax = (double)*vmax;
    thr = (thr >= -1.0) ? -thr * hmax : hmax;
  }
  const int f = (bins[b] > 0 && (double)bins[b] >= thr) ? 1 : 0;
  flags[b] = f;
  sizes[b] = f ? bins[b] : 0;
}

__global__ void hough_offsets_kernel(const int *__restrict__ flags, const int *__restrict__ slots,
                                     const int *__restrict__ starts, const int *__restrict__ sizes, int nbins,
                                     int max_inst, int *__restrict__ inst_offsets) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbins || !flags[b]) return;
  const int s = slots[b];
  if (s < max_inst) inst_offsets[s + 1] = starts[b] + sizes[b];
}

__global__ void hough_fill_kernel(const float *__restrict__ svotes, int C, HoughSpace H, const int *__restrict__ flags,
                                  const int *__restrict__ starts, int *__restrict__ cursor, int *__restrict__ members) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  const long long b = hough_bin(H, svotes + (size_t)i * 3);
  if (b < 0 || !flags[b]) return;
  members[starts[b] + atomicAdd(&cursor[b], 1)] = i;
}

// voters in ascending correspondence index (the order PCL's push_back produces): rank by counting
__global__ void hough_sort_members_kernel(const int *__restrict__ inst_offsets, const int *__restrict__ n_inst, int max_inst,
                                          const int *__restrict__ in, int *__restrict__ out) {
  const int inst = blockIdx.x;
  if (inst >= min(*n_inst, max_inst)) return;
  const int off = inst_offsets[inst], n = inst_offsets[inst + 1] - off;
  for (int a = threadIdx.x; a < n; a += blockDim.x) {
    const int v = in[off + a];
    int rank = 0;
    for (int b = 0; b < n; ++b) rank += in[off + b] < v;
    out[off + rank] = v;
  }
}

__global__ void hough_gather_points_kernel(const b200_corr *__restrict__ corrs, int C, const float4 *__restrict__ model_kp,
                                           const float4 *__restrict__ scene_kp, float4 *__restrict__ mp,
                                           float4 *__restrict__ sp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  mp[i] = model_kp[corrs[i].index_query];
  sp[i] = scene_kp[corrs[i].index_match];
}

}  // namespace

int dev_hough3d(b200_ctx *ctx, const float4 *d_model_kp, const float *d_model_rf, int Km, const float4 *d_scene_kp,
                const float *d_scene_rf, const b200_corr *d_corrs, int C, double bin_size, double threshold, float *d_T,
                int max_inst, int *d_inst_offsets, int *d_inst_counts, b200_corr *d_inst_corrs, int corr_cap,
                int *d_n_inst) {
  if (max_inst < 1) return ctx->fail(B200_ERR_INVALID, "hough3d: max_inst must be >= 1");
  if (!(bin_size > 0.0)) return ctx->fail(B200_ERR_INVALID, "hough3d: bin size must be > 0");
  B200_CUDA(ctx, cudaMemsetAsync(d_n_inst, 0, sizeof(int), ctx->stream));
  B200_CUDA(ctx, cudaMemsetAsync(d_inst_offsets, 0, sizeof(int) * ((size_t)max_inst + 1), ctx->stream));
  B200_CUDA(ctx, cudaMemsetAsync(d_inst_counts, 0, sizeof(int) * (size_t)max_inst, ctx->stream));
  if (C <= 0 || Km <= 0) return B200_OK;
  StageScope st_(ctx, ST_GC_GROUP);
  DevBuf<float> centroid, mvotes, svotes;
  DevBuf<unsigned> box;
  B200_TRY(centroid.alloc(ctx, 3));
  B200_TRY(mvotes.alloc(ctx, (size_t)Km * 3));
  B200_TRY(svotes.alloc(ctx, (size_t)C * 3));
  B200_TRY(box.alloc(ctx, 6));
  hough_centroid_kernel<<<1, 32, 0, ctx->stream>>>(d_model_kp, Km, centroid.p);
  B200_LAUNCHED(ctx);
  hough_model_votes_kernel<<<ceil_div(Km, 256), 256, 0, ctx->stream>>>(d_model_kp, d_model_rf, Km, centroid.p, mvotes.p);
  B200_LAUNCHED(ctx);
  const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  B200_CUDA(ctx, cudaMemcpyAsync(box.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  hough_scene_votes_kernel<<<ceil_div(C, 256), 256, 0, ctx->stream>>>(d_corrs, C, mvotes.p, d_scene_kp, d_scene_rf,
                                                                     svotes.p, box.p);
  B200_LAUNCHED(ctx);
  unsigned hb[6];
  B200_CUDA(ctx, cudaMemcpyAsync(hb, box.p, sizeof(hb), cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, ctx->sync());
  if (hb[0] > hb[3]) return B200_OK;  // no finite vote
  HoughSpace H;
  H.bin = bin_size;
  long long nbins = 1;
  for (int a = 0; a < 3; ++a) {
    H.mn[a] = (double)dec_ordered(hb[a]);
    const double mx = (double)dec_ordered(hb[3 + a]);
    H.cnt[a] = (int)std::ceil((mx - H.mn[a]) / bin_size);  // HoughSpace3D: bin_count = ceil((max - min) / bin)
    nbins *= std::max(H.cnt[a], 0);
    if (nbins > HOUGH_MAX_BINS) return ctx->fail(B200_ERR_CAPACITY, "hough3d: more than 2^27 bins (bin size too small)");
  }
  if (nbins == 0) return B200_OK;  // a degenerate space holds no bin (all votes on one coordinate)

  DevBuf<int> bins, vmax, flags, sizes, slots, starts, cursor, members, members_sorted;
  B200_TRY(bins.alloc(ctx, (size_t)nbins));
  B200_TRY(bins.zero());
  B200_TRY(vmax.alloc(ctx, 1));
  B200_TRY(vmax.zero());
  hough_vote_kernel<<<ceil_div(C, 256), 256, 0, ctx->stream>>>(svotes.p, C, H, bins.p, vmax.p);
  B200_LAUNCHED(ctx);
  B200_TRY(flags.alloc(ctx, (size_t)nbins));
  B200_TRY(sizes.alloc(ctx, (size_t)nbins));
  B200_TRY(slots.alloc(ctx, (size_t)nbins));
  B200_TRY(starts.alloc(ctx, (size_t)nbins));
  hough_flags_kernel<<<ceil_div(nbins, 256), 256, 0, ctx->stream>>>(bins.p, (int)nbins, threshold, vmax.p, flags.p, sizes.p);
  B200_LAUNCHED(ctx);
  B200_TRY(exclusive_scan_i32(ctx, flags.p, slots.p, (int)nbins, d_n_inst));
  B200_TRY(exclusive_scan_i32(ctx, sizes.p, starts.p, (int)nbins, nullptr));
  hough_offsets_kernel<<<ceil_div(nbins, 256), 256, 0, ctx->stream>>>(flags.p, slots.p, starts.p, sizes.p, (int)nbins,
                                                                     max_inst, d_inst_offsets);
  B200_LAUNCHED(ctx);
  B200_TRY(cursor.alloc(ctx, (size_t)nbins));
  B200_TRY(cursor.zero());
  B200_TRY(members.alloc(ctx, (size_t)C));
  B200_TRY(members_sorted.alloc(ctx, (size_t)C));
  hough_fill_kernel<<<ceil_div(C, 256), 256, 0, ctx->stream>>>(svotes.p, C, H, flags.p, starts.p, cursor.p, members.p);
  B200_LAUNCHED(ctx);
  hough_sort_members_kernel<<<max_inst, 128, 0, ctx->stream>>>(d_inst_offsets, d_n_inst, max_inst, members.p,
                                                               members_sorted.p);
  B200_LAUNCHED(ctx);
  DevBuf<float4> mp, sp;
  B200_TRY(mp.alloc(ctx, (size_t)C));
  B200_TRY(sp.alloc(ctx, (size_t)C));
  hough_gather_points_kernel<<<ceil_div(C, 256), 256, 0, ctx->stream>>>(d_corrs, C, d_model_kp, d_scene_kp, mp.p, sp.p);
  B200_LAUNCHED(ctx);
  return dev_ransac_instances(ctx, d_corrs, mp.p, sp.p, members_sorted.p, d_inst_offsets, d_n_inst, C, bin_size, d_T,
                              max_inst, d_inst_counts, d_inst_corrs, corr_cap);
}
