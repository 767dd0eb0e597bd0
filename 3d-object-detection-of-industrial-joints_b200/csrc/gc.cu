// gc.cu — geometric-consistency grouping of correspondences + RANSAC pose, on the device.
//
// Replaces pcl::GeometricConsistencyGrouping<PointXYZRGBA, PointXYZRGBA>::recognize (SHOT.cpp:473-482,
// 6Dpose.cpp:529-538, SHOT_scenes.cpp:413-425): clusterCorrespondences (sort by distance, greedy
// seed-and-grow consensus sets under the pairwise distance-preservation test, sets larger than the
// threshold are taken) followed per set by CorrespondenceRejectorSampleConsensus (RANSAC on
// SampleConsensusModelRegistration, Umeyama on 3-samples, mt19937 seeded 12345).
//
// The greedy order is kept exactly.  A seed's consensus set depends on earlier seeds only through
// the `taken` flags, and only successful seeds change those; moreover a set computed against an
// older (smaller) set of flags is still exact as long as none of its members has been taken since
// (rejected candidates never influence later decisions).  So:
//   1. gc_adjacency_kernel evaluates the pairwise test once for all pairs into a C x C bitmap
//      (all SMs; this is the O(C^2) part).
//   2. gc_group_kernel (one CTA, 16 warps) takes the next 16 untaken seeds, one per warp; a warp
//      intersects its seed's bitmap row with ~taken, lists the candidates in ascending order and grows
//      the set with warp ballots (candidates' points in registers).  Warp 0 then walks the 16 results
//      in order: the first seed whose set touches an element committed earlier in the walk is
//      inexact, everything before it commits, and the next window restarts there.  On the benchmark
//      scene (25 k correspondences, 3 770 instances) that is ~530 rounds of a few microseconds instead
//      of 4 276 sequential seed scans.
// RANSAC runs afterwards, one warp per instance: lane 0 draws the sample sequence (the RNG stream is
// inherently serial), the lanes fit and score one sample each, lane 0 then replays the adaptive
// termination logic in order and discards the samples past the stopping point.
#include <algorithm>

#include "linalg3.cuh"
#include "pcl_eigen33.cuh"

namespace {

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

// ---- pairwise consistency bitmap ------------------------------------------------------------------
// adj[i][j] = 1 iff correspondences i and j (sorted positions) pass the distance-preservation test,
// i != j.  Rows are row_words 32-bit words (a multiple of 8, so a row starts on a 32-byte sector).
// The decision is PCL's float expression; a MUFU-based estimate settles every pair that is not within
// a rigorous error band of the threshold and only the rest evaluate the IEEE square roots.
constexpr int ADJ_ROWS = 128;
constexpr int ADJ_WORDS = 32;  // words per CTA tile (1024 columns)
constexpr int ADJ_THREADS = 256;

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float dist2f(const float4 &a, const float4 &b) {
  const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z;
  float s = d0 * d0;
  s += d1 * d1;
  s += d2 * d2;
  return s;
}

// The grouping test with a fast path: `fits` <=> !gc_rejects.  sqrt.approx has a relative error of at
// most 2^-22, so |estimate - float expression| < 1e-6 (sa + sc); pairs outside that band around the
// threshold are decided by the estimate, the others evaluate the exact expression.
__device__ __forceinline__ bool gc_fits(const float4 &mk, const float4 &sk, const float4 &mj, const float4 &sj,
                                        float g_lo, float g_hi, double gc_size) {
  const float a = dist2f(sk, sj);
  const float c = dist2f(mk, mj);
  const float sa = sqrt_approx(a), sc = sqrt_approx(c);
  const float diff = fabsf(sa - sc);
  const float tol = 1e-6f * (sa + sc) + 1e-15f;
  if (diff > g_hi + tol) return false;
  if (diff < g_lo - tol) return true;
  return !((double)fabsf(sqrtf(a) - sqrtf(c)) > gc_size);
}

__device__ __forceinline__ int gc_row_words(int C) { return (((C + 31) >> 5) + 7) & ~7; }

__global__ void __launch_bounds__(ADJ_THREADS)
    gc_adjacency_kernel(const float4 *__restrict__ mp, const float4 *__restrict__ sp, const int *__restrict__ d_C,
                        int C_cap, double gc_size, float g_lo, float g_hi, unsigned *__restrict__ adj) {
  __shared__ float4 s_m[ADJ_WORDS * 32];
  __shared__ float4 s_s[ADJ_WORDS * 32];
  const int C = min(*d_C, C_cap);
  const int row_words = gc_row_words(C);
  const int w0 = blockIdx.x * ADJ_WORDS;
  const int i0 = blockIdx.y * ADJ_ROWS;
  if (w0 >= row_words || i0 >= C) return;
  const int tid = threadIdx.x;
  for (int t = tid; t < ADJ_WORDS * 32; t += ADJ_THREADS) {
    const int j = w0 * 32 + t;
    const bool in = j < C;
    s_m[t] = in ? mp[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    s_s[t] = in ? sp[j] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int i = i0 + (tid & (ADJ_ROWS - 1));
  const int wbase = (tid / ADJ_ROWS) * (ADJ_WORDS / 2);  // 16 words per thread
  if (i >= C) return;
  const float4 mi = mp[i], si = sp[i];
  unsigned *row = adj + (size_t)i * row_words + w0 + wbase;
#pragma unroll 1
  for (int wq = 0; wq < ADJ_WORDS / 2; wq += 4) {
    unsigned out[4];
#pragma unroll
    for (int w4 = 0; w4 < 4; ++w4) {
      const int w = wq + w4;
      unsigned bits = 0;
      const int t0 = (wbase + w) * 32;
#pragma unroll 8
      for (int b = 0; b < 32; ++b) {
        const bool ok = gc_fits(mi, si, s_m[t0 + b], s_s[t0 + b], g_lo, g_hi, gc_size);
        bits |= (ok ? 1u : 0u) << b;
      }
      const int gw = w0 + wbase + w;  // global word index
      const int j0 = gw * 32;
      if (j0 + 32 > C) bits &= (j0 >= C) ? 0u : ((1u << (C - j0)) - 1u);
      if ((i >> 5) == gw) bits &= ~(1u << (i & 31));
      out[w4] = bits;
    }
    if (w0 + wbase + wq < row_words) *reinterpret_cast<uint4 *>(row + wq) = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

// ---- greedy grouping: one CTA, GW seeds evaluated speculatively per round ------------------------
constexpr int GW = 16;                 // seeds (one per warp) per round
constexpr int GG_THREADS = GW * 32;
constexpr int G_MC = 64;               // members kept in shared memory per seed (index + points)
constexpr int G_CL = 1024;             // candidate list capacity per seed
constexpr int G_CPL = 4;               // candidates per lane per chunk

struct GroupArgs {
  const unsigned *adj;
  const float4 *mp;
  const float4 *sp;
  int *overflow;      // [GW][C_cap] member indices past G_MC
  int *members;       // [C_cap] committed member lists, concatenated
  int *inst_offsets;  // [max_inst + 1]
  int *n_inst_out;
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

__global__ void __launch_bounds__(GG_THREADS, 1)
    gc_group_kernel(GroupArgs ga, const int *__restrict__ d_C, int C_cap, double gc_size, float g_lo, float g_hi,
                    int gc_threshold, int max_inst) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ int s_seed[GW], s_size[GW], s_commit_off[GW], s_commit_inst[GW];
  __shared__ int s_nwin, s_cur, s_ninst, s_total;
  __shared__ float4 s_newm[GW], s_news[GW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = min(*d_C, C_cap);
  const int row_words = gc_row_words(C);
  // dynamic shared memory: member points | taken bitmap | candidate lists | member indices
  float4 *s_mp = reinterpret_cast<float4 *>(s_raw);                  // [GW][G_MC]
  float4 *s_sp = s_mp + GW * G_MC;                                   // [GW][G_MC]
  unsigned *s_taken = reinterpret_cast<unsigned *>(s_sp + GW * G_MC);  // [row_words]
  int *s_cand = reinterpret_cast<int *>(s_taken + row_words);        // [GW][G_CL]
  int *s_mem = s_cand + GW * G_CL;                                   // [GW][G_MC]

  for (int w = tid; w < row_words; w += GG_THREADS) {
    const int j0 = w * 32;
    s_taken[w] = (j0 + 32 <= C) ? 0u : ((j0 >= C) ? ~0u : ~((1u << (C - j0)) - 1u));  // padding counts as taken
  }
  if (tid == 0) {
    s_cur = 0;
    s_ninst = 0;
    s_total = 0;
    ga.inst_offsets[0] = 0;
  }
  __syncthreads();

  float4 *my_mp = s_mp + warp * G_MC, *my_sp = s_sp + warp * G_MC;
  int *my_cand = s_cand + warp * G_CL, *my_mem = s_mem + warp * G_MC;
  int *my_over = ga.overflow + (size_t)warp * C_cap;

  while (true) {
    // ---- window: the next GW untaken positions at or after s_cur (warp 0) ----
    if (warp == 0) {
      const int cur = s_cur;
      int n = 0;
      for (int wb = cur >> 5; wb < row_words && n < GW; wb += 32) {
        const int wi = wb + lane;
        unsigned bits = (wi < row_words) ? ~s_taken[wi] : 0u;
        if (wi == (cur >> 5)) bits &= ~((1u << (cur & 31)) - 1u);
        const int cnt = __popc(bits);
        const int incl = warp_incl_scan(cnt, lane);
        int o = n + incl - cnt;
        while (bits && o < GW) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          s_seed[o++] = wi * 32 + b;
        }
        n += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) s_nwin = min(n, GW);
    }
    __syncthreads();
    const int nwin = s_nwin;
    if (nwin == 0) break;

    // ---- evaluate: warp w grows the consensus set of seed w against the current flags ----
    if (warp < nwin) {
      const int seed = s_seed[warp];
      int size = 1;
      if (lane == 0) {
        my_mem[0] = seed;
        my_mp[0] = ga.mp[seed];
        my_sp[0] = ga.sp[seed];
      }
      __syncwarp();
      const uint4 *row4 = reinterpret_cast<const uint4 *>(ga.adj + (size_t)seed * row_words);
      // pass A: how many candidates (row & ~taken) and which is the first.  The first candidate is
      // always admitted (it only has to fit the seed); when many candidates remain, its bitmap row
      // is intersected as well, so the list below holds only candidates that fit both.
      int n_cand = 0, first = 0x7fffffff;
      for (int wb = lane * 8; wb < row_words; wb += 256) {
        const uint4 x = __ldg(row4 + wb / 4), y = __ldg(row4 + wb / 4 + 1);
        const unsigned w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
        for (int k = 7; k >= 0; --k) {
          const unsigned v = w[k] & ~s_taken[wb + k];
          n_cand += __popc(v);
          if (v && first > (wb + k) * 32) first = (wb + k) * 32 + __ffs(v) - 1;
        }
      }
      n_cand = warp_sum(n_cand);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      const bool pre = n_cand > 32 * G_CPL;
      const uint4 *rowj4 = row4;
      if (pre) {
        rowj4 = reinterpret_cast<const uint4 *>(ga.adj + (size_t)first * row_words);
        if (lane == 0) {
          my_mem[1] = first;
          my_mp[1] = ga.mp[first];
          my_sp[1] = ga.sp[first];
        }
        size = 2;
        __syncwarp();
      }
      const int k_start = size;  // members whose rows are already folded into the candidate list
      int ord_base = 0, n_total = 0;
      if (n_cand > 0) do {
        // candidates in ascending position; ordinals [ord_base, ord_base + G_CL)
        int seen = 0;
        for (int seg = 0; seg * 256 < row_words; ++seg) {
          const int wb = seg * 256 + lane * 8;
          unsigned w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          int cnt = 0;
          if (wb < row_words) {
            const uint4 x = __ldg(row4 + wb / 4), y = __ldg(row4 + wb / 4 + 1);
            w[0] = x.x, w[1] = x.y, w[2] = x.z, w[3] = x.w, w[4] = y.x, w[5] = y.y, w[6] = y.z, w[7] = y.w;
            if (pre) {
              const uint4 p = __ldg(rowj4 + wb / 4), q = __ldg(rowj4 + wb / 4 + 1);
              w[0] &= p.x, w[1] &= p.y, w[2] &= p.z, w[3] &= p.w, w[4] &= q.x, w[5] &= q.y, w[6] &= q.z, w[7] &= q.w;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              w[k] &= ~s_taken[wb + k];
              cnt += __popc(w[k]);
            }
          }
          const int incl = warp_incl_scan(cnt, lane);
          const int tot = __shfl_sync(0xffffffffu, incl, 31);
          if (tot) {
            int o = seen + incl - cnt - ord_base;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              unsigned bits = w[k];
              while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                if (o >= 0 && o < G_CL) my_cand[o] = (wb + k) * 32 + b;
                ++o;
              }
            }
          }
          seen += tot;
        }
        n_total = seen;
        const int n_list = min(G_CL, n_total - ord_base);
        __syncwarp();
        // ---- chunks of 128 candidates: test against the members so far, admit survivors in order ----
        int nj[G_CPL];
        float4 nm[G_CPL], ns[G_CPL];
        unsigned nalive = 0;
#pragma unroll
        for (int u = 0; u < G_CPL; ++u) {
          const int idx = u * 32 + lane;
          if (idx < n_list) {
            nj[u] = my_cand[idx];
            nm[u] = ga.mp[nj[u]];
            ns[u] = ga.sp[nj[u]];
            nalive |= 1u << u;
          }
        }
        for (int cb = 0; cb < n_list; cb += 32 * G_CPL) {
          int j[G_CPL];
          float4 mj[G_CPL], sj[G_CPL];
          unsigned alive = nalive;
#pragma unroll
          for (int u = 0; u < G_CPL; ++u) {
            j[u] = nj[u];
            mj[u] = nm[u];
            sj[u] = ns[u];
          }
          nalive = 0;
#pragma unroll
          for (int u = 0; u < G_CPL; ++u) {  // prefetch the next chunk
            const int idx = cb + 32 * G_CPL + u * 32 + lane;
            if (idx < n_list) {
              nj[u] = my_cand[idx];
              nm[u] = ga.mp[nj[u]];
              ns[u] = ga.sp[nj[u]];
              nalive |= 1u << u;
            }
          }
          for (int k = k_start; k < size; ++k) {
            if (!__any_sync(0xffffffffu, alive != 0)) break;
            float4 mk, sk;
            if (k < G_MC) {
              mk = my_mp[k];
              sk = my_sp[k];
            } else {
              const int mi = *reinterpret_cast<volatile int *>(my_over + k);
              mk = ga.mp[mi];
              sk = ga.sp[mi];
            }
#pragma unroll
            for (int u = 0; u < G_CPL; ++u)
              if (((alive >> u) & 1u) && !gc_fits(mk, sk, mj[u], sj[u], g_lo, g_hi, gc_size)) alive &= ~(1u << u);
          }
          while (true) {
            unsigned bal = 0;
            int usel = -1;
#pragma unroll
            for (int u = 0; u < G_CPL; ++u) {
              if (usel < 0) {
                bal = __ballot_sync(0xffffffffu, (alive >> u) & 1u);
                if (bal) usel = u;
              }
            }
            if (usel < 0) break;
            const int owner = __ffs(bal) - 1;
            if (lane == owner) {
              int jj = j[0];
              float4 a = mj[0], b = sj[0];
#pragma unroll
              for (int v = 1; v < G_CPL; ++v)
                if (usel == v) {
                  jj = j[v];
                  a = mj[v];
                  b = sj[v];
                }
              s_newm[warp] = a;
              s_news[warp] = b;
              if (size < G_MC) {
                my_mem[size] = jj;
                my_mp[size] = a;
                my_sp[size] = b;
              } else {
                *reinterpret_cast<volatile int *>(my_over + size) = jj;
              }
              alive &= ~(1u << usel);
            }
            __syncwarp();
            const float4 mk = s_newm[warp], sk = s_news[warp];
            ++size;
#pragma unroll
            for (int u = 0; u < G_CPL; ++u)
              if (((alive >> u) & 1u) && !gc_fits(mk, sk, mj[u], sj[u], g_lo, g_hi, gc_size)) alive &= ~(1u << u);
            __syncwarp();
          }
        }
        ord_base += G_CL;
      } while (ord_base < n_total);
      if (lane == 0) s_size[warp] = size;
    }
    __threadfence_block();
    __syncthreads();

    // ---- commit (warp 0): walk the window in order; a seed whose set touches an element taken
    // earlier in this walk is the first inexact one: everything before it is what the sequential
    // algorithm computes, the next window restarts there ----
    if (warp == 0) {
      int n_inst = s_ninst, total = s_total;
      int pstar = nwin;
      for (int r = 0; r < nwin; ++r) {
        const int sz = s_size[r];
        bool conflict = false;
        for (int k0 = 0; k0 < sz; k0 += 32) {
          const int k = k0 + lane;
          int m = -1;
          if (k < sz) m = (k < G_MC) ? s_mem[r * G_MC + k] : __ldcg(ga.overflow + (size_t)r * C_cap + k);
          const bool hit = m >= 0 && ((s_taken[m >> 5] >> (m & 31)) & 1u);
          if (__any_sync(0xffffffffu, hit)) conflict = true;
        }
        if (conflict) {
          pstar = r;
          break;
        }
        if (sz > gc_threshold) {
          for (int k0 = 0; k0 < sz; k0 += 32) {
            const int k = k0 + lane;
            if (k < sz) {
              const int m = (k < G_MC) ? s_mem[r * G_MC + k] : __ldcg(ga.overflow + (size_t)r * C_cap + k);
              atomicOr(&s_taken[m >> 5], 1u << (m & 31));
            }
          }
          if (lane == 0) {
            s_commit_off[r] = total;
            s_commit_inst[r] = n_inst;
          }
          total += sz;
          ++n_inst;
        } else if (lane == 0) {
          s_commit_off[r] = -1;
        }
        __syncwarp();
      }
      for (int r = pstar + lane; r < nwin; r += 32) s_commit_off[r] = -1;
      if (lane == 0) {
        s_cur = (pstar < nwin) ? s_seed[pstar] : s_seed[nwin - 1] + 1;
        s_ninst = n_inst;
        s_total = total;
      }
    }
    __syncthreads();
    // ---- committed seeds write their member lists ----
    if (warp < nwin && s_commit_off[warp] >= 0) {
      const int off = s_commit_off[warp], sz = s_size[warp], inst = s_commit_inst[warp];
      for (int k = lane; k < sz; k += 32)
        ga.members[off + k] = (k < G_MC) ? my_mem[k] : *reinterpret_cast<volatile int *>(my_over + k);
      if (lane == 0 && inst < max_inst) ga.inst_offsets[inst + 1] = off + sz;
    }
    // the next round's window selection (warp 0) only reads s_taken / s_cur, written before the barrier
  }
  if (tid == 0) *ga.n_inst_out = s_ninst;
}

// ---- RANSAC pose per instance -------------------------------------------------------------------
// boost::mt19937 seeded 12345 (RandomSampleConsensus, non-random mode).  Every instance restarts the
// same stream, so the generator state after seeding and the first twist is computed once on the host
// and copied into shared memory per warp.
struct Mt19937 {
  unsigned *s;  // 624 words
  int idx;
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

void mt19937_twisted_state(unsigned seed, unsigned out[624]) {
  out[0] = seed;
  for (int i = 1; i < 624; ++i) out[i] = 1812433253u * (out[i - 1] ^ (out[i - 1] >> 30)) + (unsigned)i;
  for (int i = 0; i < 624; ++i) {
    const unsigned y = (out[i] & 0x80000000u) | (out[(i + 1) % 624] & 0x7fffffffu);
    out[i] = out[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
}

struct RansacBuffers {
  const b200_corr *sorted;
  const float4 *mp;
  const float4 *sp;
  const int *members;
  const int *inst_offsets;
  const int *n_inst;
  const unsigned *mt_init;  // [624] twisted state for seed 12345
  int *shuffled;   // [C_cap]
  int *last_pos;   // [C_cap]
  int *flags;      // [C_cap]
  float *T_out;    // [max_inst][16]
  int *inst_counts;
  b200_corr *inst_corrs;
};

constexpr int RS_WARPS = 4;            // instances per CTA (one warp each)
constexpr int RS_THREADS = RS_WARPS * 32;
constexpr int RS_BATCH = 32;           // samples fitted per batch (one per lane)
constexpr int RS_FIRST_BATCH = 8;
constexpr int RS_SMALL = 128;          // instances up to this size keep their shuffle array in shared memory

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
  __shared__ unsigned s_mt[RS_WARPS][624];
  __shared__ float s_Tb[RS_WARPS][RS_BATCH][17];  // padded rows: one sample per lane
  __shared__ int s_sel[RS_WARPS][RS_BATCH][3];
  __shared__ int s_cnt[RS_WARPS][RS_BATCH];
  __shared__ float s_bestT[RS_WARPS][16];
  __shared__ float s_acc[RS_WARPS][9];
  __shared__ int s_shuf[RS_WARPS][RS_SMALL];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_inst = min(*rb.n_inst, max_inst);
  const int b = blockIdx.x * RS_WARPS + warp;
  if (b >= n_inst) return;
  const int off = rb.inst_offsets[b];
  const int n = rb.inst_offsets[b + 1] - off;
  const int *mem = rb.members + off;
  int *shuffled = (n <= RS_SMALL) ? s_shuf[warp] : rb.shuffled + off;  // drawn from by lane 0 only
  int *last_pos = rb.last_pos + off;
  int *flags = rb.flags + off;
  unsigned *mt = s_mt[warp];
  float(*Tb)[17] = s_Tb[warp];

  for (int i = lane; i < 624; i += 32) mt[i] = rb.mt_init[i];
  // index maps keyed by the model index: the last correspondence with a given index_query wins
  // (std::map in computeOriginalIndexMapping, unordered_map index_to_correspondence)
  for (int t = lane; t < n; t += 32) {
    const int q = rb.sorted[mem[t]].index_query;
    int last = t;
    for (int u = t + 1; u < n; ++u)
      if (rb.sorted[mem[u]].index_query == q) last = u;
    last_pos[t] = last;
    shuffled[t] = t;
  }
  // computeSampleDistanceThreshold: float32 single-pass covariance of the source (model) points in
  // list order, one lane per accumulator
  if (lane < 9) {
    float acc = 0.0f;
    for (int t = 0; t < n; ++t) {
      const float4 v = rb.mp[mem[t]];
      float a, c;
      switch (lane) {
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
      acc += (lane < 6) ? a * c : a;
    }
    s_acc[warp][lane] = acc / (float)n;
  }
  __syncwarp();

  // lane-0 state
  double sample_dist_thresh = 0.0;
  Mt19937 rng;
  rng.s = mt;
  rng.idx = 0;
  int iterations = 0, n_best = -2147483647;
  double k = 1.0;
  const unsigned skipped = 0;  // computeModelCoefficients cannot fail for a 3-sample
  const unsigned max_skip = (unsigned)max_iterations * 10u;
  const double log_probability = log(1.0 - 0.99);
  const double one_over_indices = 1.0 / (double)n;
  const double thresh2 = threshold * threshold;
  bool have_best = false, stop = false;
  if (lane == 0) {
    const float *a = s_acc[warp];
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
  }
  // isSampleGood needs three members whose pairwise squared (model) distances all exceed the
  // threshold.  When no such triple exists (typical: several scene points matched to one model
  // point) every one of getSamples' 1000 redraws fails and RANSAC ends without a model, whatever
  // the random stream is: detect that up front instead of replaying 3000 draws on one lane.
  if (n <= RS_SMALL && n >= 3) {
    sample_dist_thresh = __shfl_sync(0xffffffffu, sample_dist_thresh, 0);
    int any_good = 0;
    for (int pq = lane; pq < n * n && !any_good; pq += 32) {
      const int p = pq / n, q = pq % n;
      if (q <= p) continue;
      const float4 a = rb.mp[mem[p]], c = rb.mp[mem[q]];
      auto sq = [](const float4 &u, const float4 &v) {
        const float dx = u.x - v.x, dy = u.y - v.y, dz = u.z - v.z;
        return dx * dx + dy * dy + dz * dz;
      };
      // squared distances are symmetric bit for bit (the differences only change sign)
      if (!((double)sq(c, a) > sample_dist_thresh)) continue;
      for (int r = q + 1; r < n; ++r) {
        const float4 e = rb.mp[mem[r]];
        if ((double)sq(e, a) > sample_dist_thresh && (double)sq(e, c) > sample_dist_thresh) {
          any_good = 1;
          break;
        }
      }
    }
    if (!__any_sync(0xffffffffu, any_good)) stop = true;  // lane 0 then draws nothing: "no samples could be selected"
  }

  int batch_cap = RS_FIRST_BATCH;
  while (true) {
    // ---- lane 0: draw the next samples of the (serial) sample sequence ----
    int nb = 0, draw_failed = 0;
    if (lane == 0) {
      if (!stop && (double)iterations < k && skipped < max_skip && n >= 3) {
        for (; nb < batch_cap; ++nb) {
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
            draw_failed = 1;  // "No samples could be selected": the loop ends when it gets here
            break;
          }
          s_sel[warp][nb][0] = sel[0];
          s_sel[warp][nb][1] = sel[1];
          s_sel[warp][nb][2] = sel[2];
        }
      }
    }
    nb = __shfl_sync(0xffffffffu, nb, 0);
    draw_failed = __shfl_sync(0xffffffffu, draw_failed, 0);
    if (nb == 0) break;
    // ---- lane l: model from sample l (computeModelCoefficients) ----
    if (lane < nb) {
      double src[9], dst[9];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int t = s_sel[warp][lane][i];
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
#pragma unroll
      for (int i = 0; i < 16; ++i) Tb[lane][i] = (float)Td[i];
    }
    __syncwarp();
    // ---- countWithinDistance ----
    if (n <= 64) {
      if (lane < nb) {
        int cnt = 0;
        for (int t = 0; t < n; ++t)
          cnt += ((double)residual2(Tb[lane], rb.mp[mem[t]], rb.sp[mem[t]]) < thresh2) ? 1 : 0;
        s_cnt[warp][lane] = cnt;
      }
    } else {
      for (int sidx = 0; sidx < nb; ++sidx) {
        int cnt = 0;
        for (int t = lane; t < n; t += 32)
          cnt += ((double)residual2(Tb[sidx], rb.mp[mem[t]], rb.sp[mem[t]]) < thresh2) ? 1 : 0;
        cnt = warp_sum(cnt);
        if (lane == 0) s_cnt[warp][sidx] = cnt;
      }
    }
    __syncwarp();
    // ---- lane 0: replay the sequential loop over the batch ----
    if (lane == 0) {
      for (int i = 0; i < nb; ++i) {
        if (i > 0 && !((double)iterations < k && skipped < max_skip)) {
          stop = true;
          break;
        }
        const int c = s_cnt[warp][i];
        if (c > n_best) {
          n_best = c;
          have_best = true;
          for (int e = 0; e < 16; ++e) s_bestT[warp][e] = Tb[i][e];
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
      if (draw_failed) stop = true;
    }
    __syncwarp();
    batch_cap = RS_BATCH;
  }
  // ---- result: inliers of the best model, filtered correspondences ----
  const bool ok = __shfl_sync(0xffffffffu, have_best ? 1 : 0, 0) != 0;
  __syncwarp();
  int n_inl = 0;
  if (ok) {
    // ordered compaction of the inlier positions; flags[] receives the list
    for (int base = 0; base < n; base += 32) {
      const int t = base + lane;
      int f = 0;
      if (t < n) f = ((double)residual2(s_bestT[warp], rb.mp[mem[t]], rb.sp[mem[t]]) < thresh2) ? 1 : 0;
      const unsigned m = __ballot_sync(0xffffffffu, f);
      if (f) flags[n_inl + __popc(m & ((1u << lane) - 1u))] = t;
      n_inl += __popc(m);
    }
  }
  __syncwarp();
  const bool use_model = ok && n_inl >= 3;
  float *T = rb.T_out + (size_t)b * 16;
  if (lane < 16) T[lane] = use_model ? s_bestT[warp][lane] : ((lane % 5 == 0) ? 1.0f : 0.0f);
  const int out_n = use_model ? n_inl : n;
  for (int i = lane; i < out_n; i += 32) {
    const int t = use_model ? last_pos[flags[i]] : i;
    if (off + i < corr_cap) rb.inst_corrs[off + i] = rb.sorted[mem[t]];
  }
  if (lane == 0) rb.inst_counts[b] = out_n;
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
  // The consistency bitmap is C x C bits.  Up to GC_ASYNC_CAP correspondences it is sized for the
  // capacity and the whole stage stays asynchronous; above that the actual count is read back first.
  constexpr int GC_ASYNC_CAP = 131072;          // 2 GiB bitmap
  constexpr long long GC_MAX_C = 524288;        // 32 GiB bitmap
  int C_eff = C_cap;
  if (C_cap > GC_ASYNC_CAP) {
    int C_now = 0;
    B200_CUDA(ctx, cudaMemcpyAsync(&C_now, d_C, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    C_eff = std::max(1, std::min(C_now, C_cap));
    if (C_eff > GC_MAX_C) return ctx->fail(B200_ERR_CAPACITY, "gc: more than 524288 correspondences");
  }
  const int row_words_cap = ((((C_eff + 31) >> 5) + 7) & ~7);
  const float g_lo = nextafterf((float)gc_size, -INFINITY), g_hi = nextafterf((float)gc_size, INFINITY);
  DevBuf<b200_corr> sorted;
  DevBuf<float4> mp, sp;
  DevBuf<unsigned> adj;
  DevBuf<int> overflow, members, shuffled, last_pos, flags;
  B200_TRY(sorted.alloc(ctx, (size_t)C_eff));
  B200_TRY(mp.alloc(ctx, (size_t)C_eff));
  B200_TRY(sp.alloc(ctx, (size_t)C_eff));
  B200_TRY(adj.alloc(ctx, (size_t)C_eff * row_words_cap));
  {
    StageScope st_(ctx, ST_GC_SORT);
    gc_rank_kernel<<<ceil_div(C_eff, 256), 256, 0, ctx->stream>>>(d_corrs, d_C, C_eff, d_model_kp, d_scene_kp,
                                                                 sorted.p, mp.p, sp.p);
    B200_LAUNCHED(ctx);
  }
  {
    StageScope st_(ctx, ST_GC_ADJ);
    dim3 grid(ceil_div(row_words_cap, ADJ_WORDS), ceil_div(C_eff, ADJ_ROWS));
    gc_adjacency_kernel<<<grid, ADJ_THREADS, 0, ctx->stream>>>(mp.p, sp.p, d_C, C_eff, gc_size, g_lo, g_hi, adj.p);
    B200_LAUNCHED(ctx);
  }
  B200_TRY(overflow.alloc(ctx, (size_t)GW * C_eff));
  B200_TRY(members.alloc(ctx, (size_t)C_eff));
  GroupArgs ga;
  ga.adj = adj.p;
  ga.mp = mp.p;
  ga.sp = sp.p;
  ga.overflow = overflow.p;
  ga.members = members.p;
  ga.inst_offsets = d_inst_offsets;
  ga.n_inst_out = d_n_inst;
  {
    StageScope st_(ctx, ST_GC_GROUP);
    const size_t smem = (size_t)GW * G_MC * 2 * sizeof(float4) + (size_t)row_words_cap * sizeof(unsigned) +
                        (size_t)GW * G_CL * sizeof(int) + (size_t)GW * G_MC * sizeof(int);
    static_assert(GW * G_MC * 2 * sizeof(float4) + (GC_MAX_C / 32 + 8) * 4 + GW * G_CL * 4 + GW * G_MC * 4 <= 200 * 1024,
                  "grouping kernel shared memory");
    B200_CUDA(ctx, cudaFuncSetAttribute(gc_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gc_group_kernel<<<1, GG_THREADS, smem, ctx->stream>>>(ga, d_C, C_eff, gc_size, g_lo, g_hi, gc_threshold,
                                                          max_inst);
    B200_LAUNCHED(ctx);
  }

  if (!ctx->mt_state) {
    unsigned host_state[624];
    mt19937_twisted_state(12345u, host_state);
    B200_CUDA(ctx, cudaMalloc(&ctx->mt_state, sizeof(host_state)));
    B200_CUDA(ctx, cudaMemcpyAsync(ctx->mt_state, host_state, sizeof(host_state), cudaMemcpyHostToDevice, ctx->stream));
    B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  B200_TRY(shuffled.alloc(ctx, (size_t)C_eff));
  B200_TRY(last_pos.alloc(ctx, (size_t)C_eff));
  B200_TRY(flags.alloc(ctx, (size_t)C_eff));
  RansacBuffers rb;
  rb.sorted = sorted.p;
  rb.mp = mp.p;
  rb.sp = sp.p;
  rb.members = members.p;
  rb.inst_offsets = d_inst_offsets;
  rb.n_inst = d_n_inst;
  rb.mt_init = ctx->mt_state;
  rb.shuffled = shuffled.p;
  rb.last_pos = last_pos.p;
  rb.flags = flags.p;
  rb.T_out = d_T;
  rb.inst_counts = d_inst_counts;
  rb.inst_corrs = d_inst_corrs;
  StageScope st_(ctx, ST_GC_RANSAC);
  gc_ransac_kernel<<<ceil_div(max_inst, RS_WARPS), RS_THREADS, 0, ctx->stream>>>(rb, max_inst, corr_cap, gc_size, 10000);
  B200_LAUNCHED(ctx);
  return B200_OK;
}
