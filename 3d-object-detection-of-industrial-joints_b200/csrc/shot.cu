// shot.cu — SHOT local reference frame + SHOT352 descriptor, fused, one keypoint per CTA.
//
// Replaces pcl::SHOTEstimationOMP<PointXYZRGBA, Normal, SHOT352>::compute (SHOT.cpp:360-371,
// SHOT_demo.cpp:419-424 and 497-502, 6Dpose.cpp:450-461, CAD_desc.cpp:341-352), including the
// SHOTLocalReferenceFrameEstimationOMP pass that its initCompute runs first.  PCL searches the
// neighbourhood twice (once per pass, same radius); here the gathered list is reused.
//
// Per keypoint: radius gather (warp-aggregated append into shared memory) → sort by (d2, index) →
// weighted covariance in float64 (block reduction) → 3x3 eigen-solve → sign disambiguation by
// majority vote, with PCL's median rule on ties → per neighbour: 32-sector / 11-bin quadrilinear
// interpolation evaluated in float64 exactly as PCL does, accumulated into a 352-bin float32
// histogram in shared memory with atomics → L2 normalisation → one coalesced 1444-byte store.
// Discrete choices (sector, bin, sign votes) use the same float64 expressions as PCL; only the
// float32 accumulation order differs (atomics), which is far inside the 1e-4 L2 parity bound.
//
// Algorithmic HBM traffic per descriptor: 1444 B written + the neighbourhood's points and normals
// (32 B each) read once: 1444 + 32 * N/K bytes amortised.
#include <algorithm>

#include "linalg3.cuh"
#include "search.cuh"

namespace {

constexpr int SHOT_THREADS = 128;
constexpr int SHOT_LEN = 352;

__global__ void gather_normals_kernel(const float4 *__restrict__ sorted_pts, int n, const float4 *__restrict__ normals,
                                      float4 *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = normals[orig_index(sorted_pts[i])];
}

__global__ void __launch_bounds__(SHOT_THREADS)
    shot_kernel(GridView g, const float4 *__restrict__ nrm, const float4 *__restrict__ kp, int K, float radius_f,
                double radius, float r2, int cap, unsigned long long *glob_key, int *glob_pos,
                float *__restrict__ desc, float *__restrict__ rf_out, int lrf_only) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ int s_count;
  __shared__ int s_votes[3];       // skipped, plusX, plusZ
  __shared__ double s_red[7][4];   // block reduction scratch
  __shared__ double s_axes[6];     // v1 (x axis), v3 (z axis) before disambiguation
  __shared__ float s_frame[9];     // x, y, z axes (float)
  __shared__ int s_ok;
  __shared__ float s_hist[SHOT_LEN];
  __shared__ double s_norm[4];

  unsigned long long *key;
  int *pos;
  if (glob_key) {
    key = glob_key + (size_t)blockIdx.x * cap;
    pos = glob_pos + (size_t)blockIdx.x * cap;
  } else {
    key = reinterpret_cast<unsigned long long *>(smem_raw);
    pos = reinterpret_cast<int *>(smem_raw + (size_t)cap * sizeof(unsigned long long));
  }
  const int tid = threadIdx.x;
  const float4 *__restrict__ pts = g.pts;

  for (int i = blockIdx.x; i < K; i += gridDim.x) {
    const float4 c = kp[i];
    int n = gather_radius(g, c.x, c.y, c.z, radius_f, r2, key, pos, cap, &s_count);
    if (n > cap) n = cap;
    bitonic_sort(key, pos, n);
    if (tid < 3) s_votes[tid] = 0;
    for (int b = tid; b < SHOT_LEN; b += SHOT_THREADS) s_hist[b] = 0.0f;
    __syncthreads();

    // ---- local reference frame: weighted covariance (shot_lrf.hpp getLocalRF) ----
    double part[7] = {0, 0, 0, 0, 0, 0, 0};
    int skipped = 0;
    for (int j = tid; j < n; j += SHOT_THREADS) {
      const float4 p = pts[pos[j]];
      if (p.x == c.x && p.y == c.y && p.z == c.z) {
        ++skipped;
        continue;
      }
      const double vx = (double)(p.x - c.x), vy = (double)(p.y - c.y), vz = (double)(p.z - c.z);
      const double w = radius - sqrt((double)key_d2(key[j]));
      part[0] += w * (vx * vx);
      part[1] += w * (vx * vy);
      part[2] += w * (vx * vz);
      part[3] += w * (vy * vy);
      part[4] += w * (vy * vz);
      part[5] += w * (vz * vz);
      part[6] += w;
    }
#pragma unroll
    for (int a = 0; a < 7; ++a) part[a] = warp_sum(part[a]);
    skipped = warp_sum(skipped);
    if ((tid & 31) == 0) {
#pragma unroll
      for (int a = 0; a < 7; ++a) s_red[a][tid >> 5] = part[a];
      if (skipped) atomicAdd(&s_votes[0], skipped);
    }
    __syncthreads();
    const int n_skip = s_votes[0];
    const int valid = n - n_skip;
    if (tid == 0) {
      int ok = (valid >= 5) ? 1 : 0;
      if (ok) {
        double s[7];
#pragma unroll
        for (int a = 0; a < 7; ++a) s[a] = ((s_red[a][0] + s_red[a][1]) + s_red[a][2]) + s_red[a][3];
        double cov[9] = {s[0] / s[6], s[1] / s[6], s[2] / s[6], s[1] / s[6], s[3] / s[6],
                         s[4] / s[6], s[2] / s[6], s[4] / s[6], s[5] / s[6]};
        double w[3], V[9];
        eigh3_f64(cov, w, V);
        if (!isfinite(w[0]) || !isfinite(w[1]) || !isfinite(w[2])) ok = 0;
        s_axes[0] = V[0 * 3 + 2];
        s_axes[1] = V[1 * 3 + 2];
        s_axes[2] = V[2 * 3 + 2];
        s_axes[3] = V[0 * 3 + 0];
        s_axes[4] = V[1 * 3 + 0];
        s_axes[5] = V[2 * 3 + 0];
      }
      s_ok = ok;
    }
    __syncthreads();
    int ok = s_ok;
    if (ok) {
      // sign votes
      const double v1x = s_axes[0], v1y = s_axes[1], v1z = s_axes[2];
      const double v3x = s_axes[3], v3y = s_axes[4], v3z = s_axes[5];
      int px = 0, pz = 0;
      for (int j = tid; j < n; j += SHOT_THREADS) {
        const float4 p = pts[pos[j]];
        if (p.x == c.x && p.y == c.y && p.z == c.z) continue;
        const double vx = (double)(p.x - c.x), vy = (double)(p.y - c.y), vz = (double)(p.z - c.z);
        if (vx * v1x + vy * v1y + vz * v1z >= 0) ++px;
        if (vx * v3x + vy * v3y + vz * v3z >= 0) ++pz;
      }
      px = warp_sum(px);
      pz = warp_sum(pz);
      if ((tid & 31) == 0) {
        if (px) atomicAdd(&s_votes[1], px);
        if (pz) atomicAdd(&s_votes[2], pz);
      }
      __syncthreads();
      if (tid == 0) {
        double ax[2][3] = {{v1x, v1y, v1z}, {v3x, v3y, v3z}};
        for (int which = 0; which < 2; ++which) {
          int plus = 2 * s_votes[1 + which] - valid;
          double *a = ax[which];
          if (plus == 0) {
            // tie: look at the 5 valid rows around the median distance (rows are distance sorted;
            // the skipped rows are the d2 == 0 prefix of the list)
            const int med = valid / 2;
            for (int t = -2; t <= 2; ++t) {
              const float4 p = pts[pos[n_skip + med - t]];
              const double vx = (double)(p.x - c.x), vy = (double)(p.y - c.y), vz = (double)(p.z - c.z);
              if (vx * a[0] + vy * a[1] + vz * a[2] > 0) ++plus;
            }
            if (plus < 3) {
              a[0] = -a[0];
              a[1] = -a[1];
              a[2] = -a[2];
            }
          } else if (plus < 0) {
            a[0] = -a[0];
            a[1] = -a[1];
            a[2] = -a[2];
          }
        }
        const float x0 = (float)ax[0][0], x1 = (float)ax[0][1], x2 = (float)ax[0][2];
        const float z0 = (float)ax[1][0], z1 = (float)ax[1][1], z2 = (float)ax[1][2];
        s_frame[0] = x0;
        s_frame[1] = x1;
        s_frame[2] = x2;
        s_frame[3] = z1 * x2 - z2 * x1;  // y = z cross x (float)
        s_frame[4] = z2 * x0 - z0 * x2;
        s_frame[5] = z0 * x1 - z1 * x0;
        s_frame[6] = z0;
        s_frame[7] = z1;
        s_frame[8] = z2;
      }
      __syncthreads();
    }

    if (lrf_only) {
      if (tid < 9) rf_out[(size_t)i * 9 + tid] = ok ? s_frame[tid] : nanf32();
      __syncthreads();
      continue;
    }

    // SHOTEstimation::computeFeature: non-finite keypoint, NaN frame or empty search → NaN row;
    // computePointSHOT: fewer than 5 neighbours → NaN descriptor.
    const bool desc_ok = ok && n >= 5;
    if (desc_ok) {
      const float fxx = s_frame[0], fxy = s_frame[1], fxz = s_frame[2];
      const float fyx = s_frame[3], fyy = s_frame[4], fyz = s_frame[5];
      const float fzx = s_frame[6], fzy = s_frame[7], fzz = s_frame[8];
      const double radius3_4 = (radius * 3) / 4, radius1_4 = radius / 4, radius1_2 = radius / 2;
      const double RAD_45 = 0.78539816339744830961566084581988;
      const double RAD_90 = 1.5707963267948966192313216916398;
      const double RAD_135 = 2.3561944901923449288469825374596;
      const double RAD_PI_7_8 = 2.7488935718910690836548129603691;
      for (int j = tid; j < n; j += SHOT_THREADS) {
        const int pj = pos[j];
        const float4 nv = nrm[pj];
        if (!finite3(nv.x, nv.y, nv.z)) continue;
        // createBinDistanceShape
        float dotf = nv.x * fzx;
        dotf += nv.y * fzy;
        dotf += nv.z * fzz;
        double cosineDesc = (double)dotf;
        if (cosineDesc > 1.0) cosineDesc = 1.0;
        if (cosineDesc < -1.0) cosineDesc = -1.0;
        double binDistance = ((1.0 + cosineDesc) * 10) / 2;
        // interpolateSingleChannel
        const float4 p = pts[pj];
        const float dx = p.x - c.x, dy = p.y - c.y, dz = p.z - c.z;
        const double distance = sqrt((double)key_d2(key[j]));
        if (fabs(distance) < 1E-15) continue;
        float t;
        t = dx * fxx;
        t += dy * fxy;
        t += dz * fxz;
        double xInFeatRef = (double)t;
        t = dx * fyx;
        t += dy * fyy;
        t += dz * fyz;
        double yInFeatRef = (double)t;
        t = dx * fzx;
        t += dy * fzy;
        t += dz * fzz;
        double zInFeatRef = (double)t;
        if (fabs(yInFeatRef) < 1E-30) yInFeatRef = 0;
        if (fabs(xInFeatRef) < 1E-30) xInFeatRef = 0;
        if (fabs(zInFeatRef) < 1E-30) zInFeatRef = 0;

        const int bit4 = ((yInFeatRef > 0) || ((yInFeatRef == 0.0) && (xInFeatRef < 0))) ? 1 : 0;
        const int bit3 = (((xInFeatRef > 0) || ((xInFeatRef == 0.0) && (yInFeatRef > 0))) ? !bit4 : bit4) ? 1 : 0;
        int desc_index = (bit4 << 3) + (bit3 << 2);
        desc_index = desc_index << 1;
        if ((xInFeatRef * yInFeatRef > 0) || (xInFeatRef == 0.0))
          desc_index += (fabs(xInFeatRef) >= fabs(yInFeatRef)) ? 0 : 4;
        else
          desc_index += (fabs(xInFeatRef) > fabs(yInFeatRef)) ? 4 : 0;
        desc_index += zInFeatRef > 0 ? 1 : 0;
        desc_index += (distance > radius1_2) ? 2 : 0;

        const int step_index = (int)floor(binDistance + 0.5);
        const int volume_index = desc_index * 11;
        binDistance -= step_index;
        double intWeight = (1 - fabs(binDistance));
        if (binDistance > 0)
          atomicAdd(&s_hist[volume_index + ((step_index + 1) % 10)], (float)binDistance);
        else
          atomicAdd(&s_hist[volume_index + ((step_index - 1 + 10) % 10)], -(float)binDistance);

        if (distance > radius1_2) {
          const double radiusDistance = (distance - radius3_4) / radius1_2;
          if (distance > radius3_4)
            intWeight += 1 - radiusDistance;
          else {
            intWeight += 1 + radiusDistance;
            atomicAdd(&s_hist[(desc_index - 2) * 11 + step_index], -(float)radiusDistance);
          }
        } else {
          const double radiusDistance = (distance - radius1_4) / radius1_2;
          if (distance < radius1_4)
            intWeight += 1 + radiusDistance;
          else {
            intWeight += 1 - radiusDistance;
            atomicAdd(&s_hist[(desc_index + 2) * 11 + step_index], (float)radiusDistance);
          }
        }

        double inclinationCos = zInFeatRef / distance;
        if (inclinationCos < -1.0) inclinationCos = -1.0;
        if (inclinationCos > 1.0) inclinationCos = 1.0;
        const double inclination = acos(inclinationCos);
        if (inclination > RAD_90 || (fabs(inclination - RAD_90) < 1e-30 && zInFeatRef <= 0)) {
          const double inclinationDistance = (inclination - RAD_135) / RAD_90;
          if (inclination > RAD_135)
            intWeight += 1 - inclinationDistance;
          else {
            intWeight += 1 + inclinationDistance;
            atomicAdd(&s_hist[(desc_index + 1) * 11 + step_index], -(float)inclinationDistance);
          }
        } else {
          const double inclinationDistance = (inclination - RAD_45) / RAD_90;
          if (inclination < RAD_45)
            intWeight += 1 + inclinationDistance;
          else {
            intWeight += 1 - inclinationDistance;
            atomicAdd(&s_hist[(desc_index - 1) * 11 + step_index], (float)inclinationDistance);
          }
        }

        if (yInFeatRef != 0.0 || xInFeatRef != 0.0) {
          const double azimuth = atan2(yInFeatRef, xInFeatRef);
          const int sel = desc_index >> 2;
          double azimuthDistance = (azimuth - (-RAD_PI_7_8 + RAD_45 * sel)) / RAD_45;
          azimuthDistance = fmax(-0.5, fmin(azimuthDistance, 0.5));
          if (azimuthDistance > 0) {
            intWeight += 1 - azimuthDistance;
            const int interp_index = (desc_index + 4) % 32;
            atomicAdd(&s_hist[interp_index * 11 + step_index], (float)azimuthDistance);
          } else {
            const int interp_index = (desc_index - 4 + 32) % 32;
            intWeight += 1 + azimuthDistance;
            atomicAdd(&s_hist[interp_index * 11 + step_index], -(float)azimuthDistance);
          }
        }
        atomicAdd(&s_hist[volume_index + step_index], (float)intWeight);
      }
      __syncthreads();
      // normalizeHistogram: acc_norm (double) += shot[j] * shot[j] (float product)
      double acc = 0.0;
      for (int b = tid; b < SHOT_LEN; b += SHOT_THREADS) {
        const float h = s_hist[b];
        acc += (double)(h * h);
      }
      acc = warp_sum(acc);
      if ((tid & 31) == 0) s_norm[tid >> 5] = acc;
      __syncthreads();
      const double acc_norm = sqrt(((s_norm[0] + s_norm[1]) + s_norm[2]) + s_norm[3]);
      const float fnorm = (float)acc_norm;
      for (int b = tid; b < SHOT_LEN; b += SHOT_THREADS) desc[(size_t)i * SHOT_LEN + b] = s_hist[b] / fnorm;
      if (rf_out && tid < 9) rf_out[(size_t)i * 9 + tid] = s_frame[tid];
    } else {
      for (int b = tid; b < SHOT_LEN; b += SHOT_THREADS) desc[(size_t)i * SHOT_LEN + b] = nanf32();
      // rf is NaN when the row is rejected by computeFeature (frame NaN / empty search); a valid frame
      // with 1..4 neighbours cannot occur (a valid frame needs >= 5 neighbours)
      if (rf_out && tid < 9) rf_out[(size_t)i * 9 + tid] = nanf32();
    }
    __syncthreads();
  }
}

}  // namespace

int dev_shot(b200_ctx *ctx, b200_cloud *c, const float *d_normals, const float4 *d_kp, int K, double radius,
             float *d_desc, float *d_rf, bool lrf_only) {
  if (!(radius > 0.0)) return ctx->fail(B200_ERR_INVALID, "shot: radius must be > 0");
  if (K <= 0) return B200_OK;
  const GridView *g;
  B200_TRY(cloud_grid_for_radius(c, radius, &g));
  // neighbour counts size the per-CTA list (and give the bench its n-bar)
  DevBuf<int> counts;
  DevBuf<unsigned long long> stats;
  B200_TRY(counts.alloc(ctx, (size_t)K));
  B200_TRY(stats.alloc(ctx, 2));
  B200_TRY(dev_radius_count(ctx, *g, d_kp, K, radius, counts.p, stats.p));
  unsigned long long hstats[2];
  B200_CUDA(ctx, cudaMemcpyAsync(hstats, stats.p, sizeof(hstats), cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int max_count = (int)hstats[0];
  ctx->last_max_nbrs = max_count;
  ctx->last_mean_nbrs = (double)hstats[1] / K;

  StageScope st_(ctx, ST_SHOT);
  DevBuf<float4> nrm_sorted;
  if (!lrf_only) {
    B200_TRY(nrm_sorted.alloc(ctx, (size_t)std::max(c->n_valid, 1)));
    if (c->n_valid > 0) {
      gather_normals_kernel<<<ceil_div(c->n_valid, 256), 256, 0, ctx->stream>>>(
          g->pts, c->n_valid, reinterpret_cast<const float4 *>(d_normals), nrm_sorted.p);
      B200_LAUNCHED(ctx);
    }
  }
  const float r2 = (float)(radius * radius);
  const int cap = next_pow2_host(std::max(max_count, 32));
  const size_t smem = (size_t)cap * 12;
  if (smem <= 96 * 1024) {
    B200_CUDA(ctx, cudaFuncSetAttribute(shot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (smem + 4096)));
    const int grid = std::min(K, ctx->sm_count * per_sm);
    shot_kernel<<<grid, SHOT_THREADS, smem, ctx->stream>>>(*g, nrm_sorted.p, d_kp, K, (float)radius, radius, r2, cap,
                                                          nullptr, nullptr, d_desc, d_rf, lrf_only ? 1 : 0);
    B200_LAUNCHED(ctx);
  } else {
    const int grid = std::min(K, ctx->sm_count * 2);
    DevBuf<unsigned long long> gk;
    DevBuf<int> gp;
    B200_TRY(gk.alloc(ctx, (size_t)grid * cap));
    B200_TRY(gp.alloc(ctx, (size_t)grid * cap));
    shot_kernel<<<grid, SHOT_THREADS, 0, ctx->stream>>>(*g, nrm_sorted.p, d_kp, K, (float)radius, radius, r2, cap,
                                                       gk.p, gp.p, d_desc, d_rf, lrf_only ? 1 : 0);
    B200_LAUNCHED(ctx);
  }
  return B200_OK;
}
