// gc.cu — geometric-consistency grouping of correspondences + RANSAC pose, on the device.
//
// Replaces pcl::GeometricConsistencyGrouping<PointXYZRGBA, PointXYZRGBA>::recognize (SHOT.cpp:473-482,
// 6Dpose.cpp:529-538, SHOT_scenes.cpp:413-425): clusterCorrespondences (sort by distance, greedy
// seed-and-grow consensus sets under the pairwise distance-preservation test, sets larger than the
// threshold are taken) followed per set by CorrespondenceRejectorSampleConsensus (RANSAC on
// SampleConsensusModelRegistration, Umeyama on 3-samples, mt19937 seeded 12345).
//
// The greedy order is kept exactly.  A seed's consensus set depends on earlier seeds only through
// the `taken` flags, and only successful seeds change those.  A persistent cooperative grid therefore
// evaluates a window of the next G seeds (one per CTA) speculatively against the current flags; the
// successful seeds of the window mark their members (atomicMin of the window position), every seed
// then checks whether one of its own members was claimed by an EARLIER successful position — if not,
// its result is what the sequential algorithm would have computed.  All successes before the first
// conflicting position commit at once and the window restarts at the conflict.  Inside a seed, 1024
// candidates are tested per step against the set so far; survivors are admitted in index order, each
// admission re-testing the later survivors, which is the sequential rule.
// RANSAC runs afterwards, one CTA per instance: thread 0 draws the sample sequence (the RNG stream
// is inherently serial), 16 warps fit and score 16 samples at a time, thread 0 then replays the
// adaptive-termination logic in order and discards the samples past the stopping point.
#include <cooperative_groups.h>

#include <algorithm>

#include "linalg3.cuh"
#include "pcl_eigen33.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int GC_THREADS = 256;
constexpr int GC_PER_THREAD = 4;
constexpr int GC_STEP = GC_THREADS * GC_PER_THREAD;
constexpr int GC_MEMBER_CACHE = 512;  // member points kept in shared memory
constexpr int GC_INF = 0x7f7f7f7f;    // memset(0x7f) pattern

// ---- sort by (distance, original position): rank by counting -----------------------------------
__global__ void __launch_bounds__(256)
    gc_rank_kernel(const b200_corr *__restrict__ corrs, const int *__restrict__ d_C, int C_cap,
                   const float4 *__restrict__ model_kp, const float4 *__restrict__ scene_kp,
                   b200_corr *__restrict__ sorted, float4 *__restrict__ mp, float4 *__restrict__ sp) {
  __shared__ unsigned long long tile[256];
  const int C = min(*d_C, C_cap);
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.x * 256 >= C) return;
  unsigned long long mykey = 0;
  b200_corr mine;
  if (i < C) {
    mine = corrs[i];
    mykey = ((unsigned long long)__float_as_uint(mine.distance) << 32) | (unsigned)i;
  }
  int rank = 0;
  for (int base = 0; base < C; base += 256) {
    const int j = base + threadIdx.x;
    __syncthreads();
    tile[threadIdx.x] = (j < C) ? (((unsigned long long)__float_as_uint(corrs[j].distance) << 32) | (unsigned)j)
                                : ~0ull;
    __syncthreads();
#pragma unroll 8
    for (int t = 0; t < 256; ++t) rank += (tile[t] < mykey) ? 1 : 0;
  }
  if (i < C) {
    sorted[rank] = mine;
    mp[rank] = model_kp[mine.index_query];
    sp[rank] = scene_kp[mine.index_match];
  }
}

__device__ __forceinline__ float norm3f(const float4 &a, const float4 &b) {
  const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z;
  float s = d0 * d0;
  s += d1 * d1;
  s += d2 * d2;
  return sqrtf(s);
}

// |‖s_k − s_j‖ − ‖m_k − m_j‖| > gc_size → j does not fit member k
__device__ __forceinline__ bool gc_rejects(const float4 &mk, const float4 &sk, const float4 &mj, const float4 &sj,
                                           double gc_size) {
  const double distance = (double)fabsf(norm3f(sk, sj) - norm3f(mk, mj));
  return distance > gc_size;
}

struct GcState {
  const b200_corr *sorted;
  const float4 *mp;
  const float4 *sp;
  unsigned char *taken;  // [C_cap + 4]
  int *mark;             // [C_cap], GC_INF when unclaimed
  int *res_size;         // [G]
  int *conf;             // [G]
  int *scratch;          // [G][C_cap] member lists of the seeds under evaluation
  int *members;          // [C_cap] committed member lists, concatenated
  int *inst_offsets;     // [max_inst + 1]
  int *n_inst_out;
};

__global__ void __launch_bounds__(GC_THREADS)
    gc_group_kernel(GcState st, const int *__restrict__ d_C, int C_cap, double gc_size, int gc_threshold,
                    int max_inst) {
  cg::grid_group grid = cg::this_grid();
  __shared__ float4 s_mp[GC_MEMBER_CACHE];
  __shared__ float4 s_sp[GC_MEMBER_CACHE];
  __shared__ unsigned s_mask[GC_THREADS / 32];
  __shared__ int s_red[5];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, bid = blockIdx.x;
  const int C = min(*d_C, C_cap);
  int *my_members = st.scratch + (size_t)bid * C_cap;
  int cur = 0, n_inst = 0, total_members = 0;
  if (bid == 0 && tid == 0) st.inst_offsets[0] = 0;

  while (cur < C) {
    // ---- evaluate the seed of this CTA against the current `taken` flags ----
    const int seed = cur + bid;
    int size = 0;
    if (seed < C && !__ldcg(&st.taken[seed])) {
      if (tid == 0) {
        my_members[0] = seed;
        s_mp[0] = st.mp[seed];
        s_sp[0] = st.sp[seed];
      }
      size = 1;
      __syncthreads();
      for (int base = 0; base < C; base += GC_STEP) {
        const int j0 = base + tid * GC_PER_THREAD;
        unsigned alive = 0;
        float4 mj[GC_PER_THREAD], sj[GC_PER_THREAD];
        if (j0 < C) {
          const unsigned tk = __ldcg(reinterpret_cast<const unsigned *>(st.taken + j0));  // 4 flags
#pragma unroll
          for (int u = 0; u < GC_PER_THREAD; ++u) {
            const int j = j0 + u;
            if (j < C && j != seed && ((tk >> (8 * u)) & 0xffu) == 0) alive |= 1u << u;
          }
        }
#pragma unroll
        for (int u = 0; u < GC_PER_THREAD; ++u)
          if ((alive >> u) & 1u) {
            mj[u] = st.mp[j0 + u];
            sj[u] = st.sp[j0 + u];
          }
        for (int k = 0; k < size && alive; ++k) {
          float4 mk, sk;
          if (k < GC_MEMBER_CACHE) {
            mk = s_mp[k];
            sk = s_sp[k];
          } else {
            const int mi = my_members[k];
            mk = st.mp[mi];
            sk = st.sp[mi];
          }
#pragma unroll
          for (int u = 0; u < GC_PER_THREAD; ++u)
            if (((alive >> u) & 1u) && gc_rejects(mk, sk, mj[u], sj[u], gc_size)) alive &= ~(1u << u);
        }
        // admit survivors in index order; each admission re-tests the later survivors
        while (__syncthreads_or(alive != 0)) {
          const unsigned m = __ballot_sync(0xffffffffu, alive != 0);
          if (lane == 0) s_mask[warp] = m;
          __syncthreads();
          int first = -1;
#pragma unroll
          for (int w = 0; w < GC_THREADS / 32; ++w) {
            const unsigned mw = s_mask[w];
            if (first < 0 && mw) first = w * 32 + __ffs(mw) - 1;
          }
          if (tid == first) {
            const int u = __ffs(alive) - 1;
            my_members[size] = j0 + u;
            if (size < GC_MEMBER_CACHE) {
              // dynamic register-array index avoided: select by u
              float4 a = mj[0], b = sj[0];
#pragma unroll
              for (int v = 1; v < GC_PER_THREAD; ++v)
                if (u == v) {
                  a = mj[v];
                  b = sj[v];
                }
              s_mp[size] = a;
              s_sp[size] = b;
            }
            alive &= ~(1u << u);
          }
          __syncthreads();
          float4 mk, sk;
          if (size < GC_MEMBER_CACHE) {
            mk = s_mp[size];
            sk = s_sp[size];
          } else {
            const int mi = my_members[size];
            mk = st.mp[mi];
            sk = st.sp[mi];
          }
          ++size;
#pragma unroll
          for (int u = 0; u < GC_PER_THREAD; ++u)
            if (((alive >> u) & 1u) && gc_rejects(mk, sk, mj[u], sj[u], gc_size)) alive &= ~(1u << u);
        }
      }
    }
    // ---- publish: size, and successful seeds claim their members with their window position ----
    const bool succ = size > gc_threshold;
    if (tid == 0) st.res_size[bid] = size;
    if (succ)
      for (int k = tid; k < size; k += GC_THREADS) atomicMin(&st.mark[my_members[k]], bid);
    grid.sync();
    // ---- conflict: one of my members (or my seed) belongs to an earlier successful seed ----
    int conflict = 0;
    for (int k = tid; k < size; k += GC_THREADS)
      if (__ldcg(&st.mark[my_members[k]]) < bid) conflict = 1;
    conflict = __syncthreads_or(conflict);
    if (tid == 0) st.conf[bid] = conflict;
    grid.sync();
    // ---- everything before the first conflicting position is exact: commit its successes ----
    if (tid < 5) s_red[tid] = (tid == 0) ? G : 0;
    __syncthreads();
    for (int b = tid; b < G; b += GC_THREADS)
      if (__ldcg(&st.conf[b])) atomicMin(&s_red[0], b);
    __syncthreads();
    const int pstar = s_red[0];
    {
      int tot = 0, cnt = 0, off = 0, idx = 0;
      for (int b = tid; b < pstar; b += GC_THREADS) {
        const int sz = __ldcg(&st.res_size[b]);
        if (sz > gc_threshold) {
          tot += sz;
          ++cnt;
          if (b < bid) {
            off += sz;
            ++idx;
          }
        }
      }
      tot = warp_sum(tot);
      cnt = warp_sum(cnt);
      off = warp_sum(off);
      idx = warp_sum(idx);
      if (lane == 0) {
        atomicAdd(&s_red[1], tot);
        atomicAdd(&s_red[2], cnt);
        atomicAdd(&s_red[3], off);
        atomicAdd(&s_red[4], idx);
      }
    }
    __syncthreads();
    if (succ) {
      if (bid < pstar) {
        const int o = total_members + s_red[3];
        for (int k = tid; k < size; k += GC_THREADS) {
          const int mi = my_members[k];
          st.members[o + k] = mi;
          st.taken[mi] = 1;
          st.mark[mi] = GC_INF;
        }
        const int inst = n_inst + s_red[4];
        if (tid == 0 && inst < max_inst) st.inst_offsets[inst + 1] = o + size;
      } else {
        for (int k = tid; k < size; k += GC_THREADS) st.mark[my_members[k]] = GC_INF;
      }
    }
    total_members += s_red[1];
    n_inst += s_red[2];
    cur += pstar;
    grid.sync();
  }
  if (bid == 0 && tid == 0) *st.n_inst_out = n_inst;
}

// ---- RANSAC pose per instance -------------------------------------------------------------------
struct Mt19937 {
  unsigned *s;  // 624 words
  int idx;
  __device__ void seed(unsigned v) {
    s[0] = v;
    for (int i = 1; i < 624; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (unsigned)i;
    idx = 624;
  }
  __device__ unsigned next() {
    if (idx >= 624) {
      for (int i = 0; i < 624; ++i) {
        const unsigned y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
        s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    unsigned y = s[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};

struct RansacBuffers {
  const b200_corr *sorted;
  const float4 *mp;
  const float4 *sp;
  const int *members;
  const int *inst_offsets;
  const int *n_inst;
  int *shuffled;   // [C_cap]
  int *last_pos;   // [C_cap]
  int *flags;      // [C_cap]
  float *T_out;    // [max_inst][16]
  int *inst_counts;
  b200_corr *inst_corrs;
};

constexpr int RS_THREADS = 512;
constexpr int RS_BATCH = RS_THREADS / 32;

// squared residual of correspondence (s → g) under the row-major 4x4 float transform T
__device__ __forceinline__ float residual2(const float *T, const float4 &s, const float4 &g) {
  float e[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float v = T[r * 4 + 0] * s.x;
    v += T[r * 4 + 1] * s.y;
    v += T[r * 4 + 2] * s.z;
    v += T[r * 4 + 3];
    e[r] = v - (r == 0 ? g.x : (r == 1 ? g.y : g.z));
  }
  float d = e[0] * e[0];
  d += e[1] * e[1];
  d += e[2] * e[2];
  return d;
}

__global__ void __launch_bounds__(RS_THREADS)
    gc_ransac_kernel(RansacBuffers rb, int max_inst, int corr_cap, double threshold, int max_iterations) {
  __shared__ unsigned s_mt[624];
  __shared__ float s_Tb[RS_BATCH][16];
  __shared__ int s_sel[RS_BATCH][3];
  __shared__ int s_cnt[RS_BATCH];
  __shared__ float s_bestT[16];
  __shared__ int s_ctrl[4];  // 0: samples in this batch, 2: have_best
  __shared__ float s_acc[9];
  __shared__ int s_warp_cnt[RS_BATCH];
  const int b = blockIdx.x;
  const int n_inst = min(*rb.n_inst, max_inst);
  if (b >= n_inst) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int off = rb.inst_offsets[b];
  const int n = rb.inst_offsets[b + 1] - off;
  const int *mem = rb.members + off;
  int *shuffled = rb.shuffled + off;
  int *last_pos = rb.last_pos + off;
  int *flags = rb.flags + off;

  // index maps keyed by the model index: the last correspondence with a given index_query wins
  // (std::map in computeOriginalIndexMapping, unordered_map index_to_correspondence)
  for (int t = tid; t < n; t += RS_THREADS) {
    const int q = rb.sorted[mem[t]].index_query;
    int last = t;
    for (int u = t + 1; u < n; ++u)
      if (rb.sorted[mem[u]].index_query == q) last = u;
    last_pos[t] = last;
    shuffled[t] = t;
  }
  // computeSampleDistanceThreshold: float32 single-pass covariance of the source (model) points in
  // list order, one lane per accumulator
  if (tid < 9) {
    float acc = 0.0f;
    for (int t = 0; t < n; ++t) {
      const float4 v = rb.mp[mem[t]];
      float a, c;
      switch (tid) {
        case 0: a = v.x, c = v.x; break;
        case 1: a = v.x, c = v.y; break;
        case 2: a = v.x, c = v.z; break;
        case 3: a = v.y, c = v.y; break;
        case 4: a = v.y, c = v.z; break;
        case 5: a = v.z, c = v.z; break;
        case 6: a = v.x, c = 1.0f; break;
        case 7: a = v.y, c = 1.0f; break;
        default: a = v.z, c = 1.0f; break;
      }
      acc += (tid < 6) ? a * c : a;
    }
    s_acc[tid] = acc / (float)n;
  }
  __syncthreads();

  // thread-0 state
  double sample_dist_thresh = 0.0;
  Mt19937 rng;
  rng.s = s_mt;
  rng.idx = 624;
  int iterations = 0, n_best = -2147483647;
  double k = 1.0;
  const unsigned skipped = 0;  // computeModelCoefficients cannot fail for a 3-sample
  const unsigned max_skip = (unsigned)max_iterations * 10u;
  const double log_probability = log(1.0 - 0.99);
  const double one_over_indices = 1.0 / (double)n;
  const double thresh2 = threshold * threshold;
  bool have_best = false, stop = false;
  if (tid == 0) {
    const float *a = s_acc;
    float cov[9];
    cov[0] = a[0] - a[6] * a[6];
    cov[1] = a[1] - a[6] * a[7];
    cov[2] = a[2] - a[6] * a[8];
    cov[4] = a[3] - a[7] * a[7];
    cov[5] = a[4] - a[7] * a[8];
    cov[8] = a[5] - a[8] * a[8];
    cov[3] = cov[1];
    cov[6] = cov[2];
    cov[7] = cov[5];
    float ev[3];
    eigen33_values(cov, ev);
    sample_dist_thresh = ((double)(sqrtf(ev[0]) + sqrtf(ev[1]) + sqrtf(ev[2]))) / 3.0;
    sample_dist_thresh *= sample_dist_thresh;
    rng.seed(12345u);
  }

  while (true) {
    // ---- thread 0: draw the next RS_BATCH samples of the (serial) sample sequence ----
    if (tid == 0) {
      int nb = 0;
      bool draw_failed = false;
      if (!stop && (double)iterations < k && skipped < max_skip && n >= 3) {
        for (; nb < RS_BATCH; ++nb) {
          bool good = false;
          int sel[3] = {0, 0, 0};
          for (int iter = 0; iter < 1000 && !good; ++iter) {  // SampleConsensusModel::getSamples
            for (int i = 0; i < 3; ++i) {                     // drawIndexSample
              const int r = (int)(rng.next() >> 1);
              const int jx = i + (r % (n - i));
              const int tmp = shuffled[i];
              shuffled[i] = shuffled[jx];
              shuffled[jx] = tmp;
            }
            sel[0] = shuffled[0];
            sel[1] = shuffled[1];
            sel[2] = shuffled[2];
            const float4 p0 = rb.mp[mem[sel[0]]], p1 = rb.mp[mem[sel[1]]], p2 = rb.mp[mem[sel[2]]];
            auto sq = [](const float4 &u, const float4 &v) {
              const float dx = u.x - v.x, dy = u.y - v.y, dz = u.z - v.z;
              return dx * dx + dy * dy + dz * dz;
            };
            good = (double)sq(p1, p0) > sample_dist_thresh && (double)sq(p2, p0) > sample_dist_thresh &&
                   (double)sq(p2, p1) > sample_dist_thresh;  // isSampleGood
          }
          if (!good) {
            draw_failed = true;  // "No samples could be selected": the loop ends when it gets here
            break;
          }
          s_sel[nb][0] = sel[0];
          s_sel[nb][1] = sel[1];
          s_sel[nb][2] = sel[2];
        }
      }
      s_ctrl[0] = nb;
      s_ctrl[1] = draw_failed ? 1 : 0;
    }
    __syncthreads();
    const int nb = s_ctrl[0];
    if (nb == 0) break;
    // ---- warp w: model from sample w (computeModelCoefficients), then countWithinDistance ----
    if (warp < nb) {
      if (lane == 0) {
        double src[9], dst[9];
        for (int i = 0; i < 3; ++i) {
          const int t = s_sel[warp][i];
          const float4 s = rb.mp[mem[t]];
          const float4 g = rb.sp[mem[last_pos[t]]];
          src[i * 3 + 0] = s.x;
          src[i * 3 + 1] = s.y;
          src[i * 3 + 2] = s.z;
          dst[i * 3 + 0] = g.x;
          dst[i * 3 + 1] = g.y;
          dst[i * 3 + 2] = g.z;
        }
        double Td[16];
        umeyama3(src, dst, 3, Td);
        for (int i = 0; i < 16; ++i) s_Tb[warp][i] = (float)Td[i];
      }
      __syncwarp();
      int cnt = 0;
      for (int t = lane; t < n; t += 32)
        cnt += ((double)residual2(s_Tb[warp], rb.mp[mem[t]], rb.sp[mem[t]]) < thresh2) ? 1 : 0;
      cnt = warp_sum(cnt);
      if (lane == 0) s_cnt[warp] = cnt;
    }
    __syncthreads();
    // ---- thread 0: replay the sequential loop over the batch ----
    if (tid == 0) {
      for (int i = 0; i < nb; ++i) {
        if (i > 0 && !((double)iterations < k && skipped < max_skip)) {
          stop = true;
          break;
        }
        const int c = s_cnt[i];
        if (c > n_best) {
          n_best = c;
          have_best = true;
          for (int e = 0; e < 16; ++e) s_bestT[e] = s_Tb[i][e];
          const double w = (double)n_best * one_over_indices;
          double p_no_outliers = 1.0 - pow(w, 3.0);
          p_no_outliers = fmax(2.220446049250313e-16, p_no_outliers);
          p_no_outliers = fmin(1.0 - 2.220446049250313e-16, p_no_outliers);
          k = log_probability / log(p_no_outliers);
        }
        ++iterations;
        if (iterations > max_iterations) {
          stop = true;  // "RANSAC reached the maximum number of trials"
          break;
        }
      }
      if (s_ctrl[1]) stop = true;
    }
    __syncthreads();
  }
  // ---- result: inliers of the best model, filtered correspondences ----
  if (tid == 0) s_ctrl[2] = have_best ? 1 : 0;
  __syncthreads();
  const bool ok = s_ctrl[2] != 0;
  int n_inl = 0;
  if (ok) {
    // ordered compaction of the inlier positions (block-wide, chunked); flags[] receives the list
    int base_total = 0;
    for (int base = 0; base < n; base += RS_THREADS) {
      const int t = base + tid;
      int f = 0;
      if (t < n) f = ((double)residual2(s_bestT, rb.mp[mem[t]], rb.sp[mem[t]]) < thresh2) ? 1 : 0;
      const unsigned m = __ballot_sync(0xffffffffu, f);
      if (lane == 0) s_warp_cnt[warp] = __popc(m);
      __syncthreads();
      int before = base_total, chunk_total = 0;
      for (int w = 0; w < RS_BATCH; ++w) {
        if (w < warp) before += s_warp_cnt[w];
        chunk_total += s_warp_cnt[w];
      }
      if (f) flags[before + __popc(m & ((1u << lane) - 1u))] = t;
      base_total += chunk_total;
      __syncthreads();
    }
    n_inl = base_total;
  }
  __syncthreads();
  const bool use_model = ok && n_inl >= 3;
  float *T = rb.T_out + (size_t)b * 16;
  if (tid < 16) T[tid] = use_model ? s_bestT[tid] : ((tid % 5 == 0) ? 1.0f : 0.0f);
  const int out_n = use_model ? n_inl : n;
  for (int i = tid; i < out_n; i += RS_THREADS) {
    const int t = use_model ? last_pos[flags[i]] : i;
    if (off + i < corr_cap) rb.inst_corrs[off + i] = rb.sorted[mem[t]];
  }
  if (tid == 0) rb.inst_counts[b] = out_n;
}

}  // namespace

int dev_gc(b200_ctx *ctx, const float4 *d_model_kp, const float4 *d_scene_kp, const b200_corr *d_corrs,
           const int *d_C, int C_cap, double gc_size, int gc_threshold, float *d_T, int max_inst, int *d_inst_offsets,
           int *d_inst_counts, b200_corr *d_inst_corrs, int corr_cap, int *d_n_inst) {
  if (max_inst < 1) return ctx->fail(B200_ERR_INVALID, "gc: max_inst must be >= 1");
  B200_CUDA(ctx, cudaMemsetAsync(d_n_inst, 0, sizeof(int), ctx->stream));
  B200_CUDA(ctx, cudaMemsetAsync(d_inst_offsets, 0, sizeof(int) * ((size_t)max_inst + 1), ctx->stream));
  B200_CUDA(ctx, cudaMemsetAsync(d_inst_counts, 0, sizeof(int) * (size_t)max_inst, ctx->stream));
  if (C_cap <= 0) return B200_OK;
  DevBuf<b200_corr> sorted;
  DevBuf<float4> mp, sp;
  DevBuf<unsigned char> taken;
  DevBuf<int> mark, res_size, conf, scratch, members, shuffled, last_pos, flags;
  B200_TRY(sorted.alloc(ctx, (size_t)C_cap));
  B200_TRY(mp.alloc(ctx, (size_t)C_cap));
  B200_TRY(sp.alloc(ctx, (size_t)C_cap));
  B200_TRY(taken.alloc(ctx, (size_t)C_cap + 4));
  B200_TRY(taken.zero());
  B200_TRY(mark.alloc(ctx, (size_t)C_cap));
  B200_CUDA(ctx, cudaMemsetAsync(mark.p, 0x7f, sizeof(int) * (size_t)C_cap, ctx->stream));
  {
    StageScope st_(ctx, ST_GC_SORT);
    gc_rank_kernel<<<ceil_div(C_cap, 256), 256, 0, ctx->stream>>>(d_corrs, d_C, C_cap, d_model_kp, d_scene_kp,
                                                                 sorted.p, mp.p, sp.p);
    B200_LAUNCHED(ctx);
  }

  int per_sm = 0;
  B200_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gc_group_kernel, GC_THREADS, 0));
  if (per_sm < 1) return ctx->fail(B200_ERR_CUDA, "gc: cooperative kernel does not fit on an SM");
  per_sm = std::min(per_sm, 4);
  int G = ctx->sm_count * per_sm;
  G = std::max(1, std::min(G, C_cap));
  B200_TRY(res_size.alloc(ctx, (size_t)G));
  B200_TRY(conf.alloc(ctx, (size_t)G));
  B200_TRY(scratch.alloc(ctx, (size_t)G * C_cap));
  B200_TRY(members.alloc(ctx, (size_t)C_cap));
  GcState st;
  st.sorted = sorted.p;
  st.mp = mp.p;
  st.sp = sp.p;
  st.taken = taken.p;
  st.mark = mark.p;
  st.res_size = res_size.p;
  st.conf = conf.p;
  st.scratch = scratch.p;
  st.members = members.p;
  st.inst_offsets = d_inst_offsets;
  st.n_inst_out = d_n_inst;
  {
    StageScope st_(ctx, ST_GC_GROUP);
    void *args[] = {&st, (void *)&d_C, &C_cap, &gc_size, &gc_threshold, &max_inst};
    B200_CUDA(ctx,
              cudaLaunchCooperativeKernel((void *)gc_group_kernel, dim3(G), dim3(GC_THREADS), args, 0, ctx->stream));
    ctx->launches++;
  }

  B200_TRY(shuffled.alloc(ctx, (size_t)C_cap));
  B200_TRY(last_pos.alloc(ctx, (size_t)C_cap));
  B200_TRY(flags.alloc(ctx, (size_t)C_cap));
  RansacBuffers rb;
  rb.sorted = sorted.p;
  rb.mp = mp.p;
  rb.sp = sp.p;
  rb.members = members.p;
  rb.inst_offsets = d_inst_offsets;
  rb.n_inst = d_n_inst;
  rb.shuffled = shuffled.p;
  rb.last_pos = last_pos.p;
  rb.flags = flags.p;
  rb.T_out = d_T;
  rb.inst_counts = d_inst_counts;
  rb.inst_corrs = d_inst_corrs;
  StageScope st_(ctx, ST_GC_RANSAC);
  gc_ransac_kernel<<<max_inst, RS_THREADS, 0, ctx->stream>>>(rb, max_inst, corr_cap, gc_size, 10000);
  B200_LAUNCHED(ctx);
  return B200_OK;
}
