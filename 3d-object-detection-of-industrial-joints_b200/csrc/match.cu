// match.cu — model↔scene descriptor correspondence search (exact float32 path).
//
// Replaces pcl::KdTreeFLANN<SHOT352 / FPFHSignature33>::setInputCloud + nearestKSearch and the user
// threshold loop around them: SHOT.cpp:405-423 (k = 1, d2 < 0.20), SHOT_scenes.cpp:359-365 (0.25),
// 6Dpose.cpp:464-482, SHOT_demo.cpp:508-530 and FPFH_demo.cpp:516-538 (k = 2, d0/d1 <= 1).
//
// A kd-tree over 352 dimensions prunes almost nothing, so the reference effectively evaluates all
// Ks x Km distances; this kernel does exactly that, tiled through shared memory.  Every distance is
// the FLANN L2_Simple float32 sum over d = 0..D-1 in order (separate multiply and add, no FMA), so
// the winning distance and the arg-min (ties → lower model index) are bit-identical to the CPU.
// The per-scene-row minimum across CTAs is taken with a 64-bit atomicMin on (d2 bits << 32 | index).
//
// This is the exact path; match_tc.cu adds a tcgen05 tensor-core pre-filter for large libraries and
// falls back to this kernel's arithmetic for the final (exact) rescoring.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int MT = 64;      // tile rows (scene and model)
constexpr int MKC = 32;     // descriptor chunk
constexpr int MTHREADS = 256;

__global__ void row_valid_kernel(const float *__restrict__ desc, int rows, int D, int all_dims,
                                 unsigned char *__restrict__ valid, int *__restrict__ n_valid) {
  // one warp per row
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows) return;
  bool ok = true;
  if (all_dims) {
    for (int d = lane; d < D; d += 32) ok = ok && isfinite(desc[(size_t)w * D + d]);
    ok = __all_sync(0xffffffffu, ok);
  } else {
    ok = isfinite(desc[(size_t)w * D]);
  }
  if (lane == 0) {
    valid[w] = ok ? 1 : 0;
    if (ok && n_valid) atomicAdd(n_valid, 1);
  }
}

__global__ void init_best_kernel(unsigned long long *best, int *zero_cnt, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    best[i] = ~0ull;
    zero_cnt[i] = 0;
  }
}

// Exact float32 distances of one 64-row scene block against the model tiles t0, t0 + tstep, ...; the running
// minimum per scene row goes to best[] / zero_cnt[] with atomics.  s_row: the block's scene rows (-1 = none).
__device__ __forceinline__ void match_block_tiles(const float *__restrict__ model, int Km,
                                                  const unsigned char *__restrict__ model_valid,
                                                  const float *__restrict__ scene, int D, float (*As)[MKC + 1],
                                                  float (*Bs)[MKC + 1], const int *s_row, int t0, int tstep, int tend,
                                                  unsigned long long *__restrict__ best, int *__restrict__ zero_cnt,
                                                  int n_block_rows = MT) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // a thread's scene rows are ty + 16 i: with fewer than 64 rows in the block (the handful of rows the tensor-core
  // filter leaves over) the upper i do no arithmetic
  const int imax = (n_block_rows + 15) >> 4;
  unsigned long long rbest[4] = {~0ull, ~0ull, ~0ull, ~0ull};
  int rzero[4] = {0, 0, 0, 0};
  for (int tile = t0; tile < tend; tile += tstep) {
    const int m0 = tile * MT;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int d0 = 0; d0 < D; d0 += MKC) {
      __syncthreads();
      for (int idx = threadIdx.x; idx < MT * MKC; idx += MTHREADS) {
        const int r = idx / MKC, c = idx % MKC;
        const int d = d0 + c;
        const int sr = s_row[r], mr = m0 + r;
        As[r][c] = (sr >= 0 && d < D) ? scene[(size_t)sr * D + d] : 0.0f;
        Bs[r][c] = (mr < Km && d < D) ? model[(size_t)mr * D + d] : 0.0f;
      }
      __syncthreads();
#pragma unroll 8
      for (int c = 0; c < MKC; ++c) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < imax) a[i] = As[ty + 16 * i][c];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[tx + 16 * j][c];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < imax) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float diff = a[i] - b[j];
              acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(diff, diff));
            }
          }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + tx + 16 * j;
        if (m < Km && model_valid[m]) {
          const unsigned long long k = ((unsigned long long)__float_as_uint(acc[i][j]) << 32) | (unsigned)m;
          rbest[i] = (k < rbest[i]) ? k : rbest[i];
          rzero[i] += (acc[i][j] == 0.0f) ? 1 : 0;
        }
      }
  }
  // reduce over the 16 threads (tx) that share a scene row: they are 16 consecutive lanes
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unsigned long long k = rbest[i];
    int z = rzero[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k, o);
      k = (ok < k) ? ok : k;
      z += __shfl_xor_sync(0xffffffffu, z, o);
    }
    const int sr = s_row[ty + 16 * i];
    if (tx == 0 && sr >= 0) {
      if (k != ~0ull) atomicMin(&best[sr], k);
      if (z) atomicAdd(&zero_cnt[sr], z);
    }
  }
}

// all scene rows: grid (scene blocks, model-tile slices)
__global__ void __launch_bounds__(MTHREADS)
    match_tile_kernel(const float *__restrict__ model, int Km, const unsigned char *__restrict__ model_valid,
                      const float *__restrict__ scene, int Ks, int D, unsigned long long *__restrict__ best,
                      int *__restrict__ zero_cnt) {
  __shared__ float As[MT][MKC + 1];
  __shared__ float Bs[MT][MKC + 1];
  __shared__ int s_row[MT];
  const int s0 = blockIdx.x * MT;
  if (s0 >= Ks) return;
  if (threadIdx.x < MT) {
    const int r = s0 + threadIdx.x;
    s_row[threadIdx.x] = (r < Ks) ? r : -1;
  }
  __syncthreads();
  match_block_tiles(model, Km, model_valid, scene, D, As, Bs, s_row, blockIdx.y, gridDim.y, (Km + MT - 1) / MT, best,
                    zero_cnt);
}

// the rows listed in row_map (the tensor-core filter's uncertified rows; their number is only known on the device):
// persistent CTAs take (row block, model tile) items in turn, so that a handful of rows still spreads over the GPU
__global__ void __launch_bounds__(MTHREADS)
    match_rows_kernel(const float *__restrict__ model, int Km, const unsigned char *__restrict__ model_valid,
                      const float *__restrict__ scene, int D, const int *__restrict__ row_map,
                      const int *__restrict__ n_rows_dev, unsigned long long *__restrict__ best,
                      int *__restrict__ zero_cnt) {
  __shared__ float As[MT][MKC + 1];
  __shared__ float Bs[MT][MKC + 1];
  __shared__ int s_row[MT];
  const int n_rows = *n_rows_dev;
  const int ntiles = (Km + MT - 1) / MT;
  const long long items = (long long)((n_rows + MT - 1) / MT) * ntiles;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int s0 = (int)(item / ntiles) * MT, tile = (int)(item % ntiles);
    __syncthreads();  // the previous item's readers of s_row are done
    if (threadIdx.x < MT) {
      const int r = s0 + threadIdx.x;
      s_row[threadIdx.x] = (r < n_rows) ? row_map[r] : -1;
    }
    __syncthreads();
    match_block_tiles(model, Km, model_valid, scene, D, As, Bs, s_row, tile, ntiles, tile + 1, best, zero_cnt,
                      min(MT, n_rows - s0));
  }
}

__global__ void match_flags_kernel(const unsigned long long *__restrict__ best, const int *__restrict__ zero_cnt,
                                   const unsigned char *__restrict__ scene_valid, const int *__restrict__ n_model_valid,
                                   int Ks, int mode, float thr, int *__restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ks) return;
  int f = 0;
  const unsigned long long k = best[i];
  if (scene_valid[i] && k != ~0ull) {
    const float d2 = __uint_as_float((unsigned)(k >> 32));
    if (mode == 1) {
      f = (d2 < thr) ? 1 : 0;
    } else {
      // tau = d0 / d1 <= 1: fails only when it is NaN (d0 == d1 == 0) or there is no second neighbour
      f = (*n_model_valid >= 2 && zero_cnt[i] < 2) ? 1 : 0;
    }
  }
  flags[i] = f;
}

__global__ void match_emit_kernel(const unsigned long long *__restrict__ best, const int *__restrict__ flags,
                                  const int *__restrict__ slots, int Ks, b200_corr *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ks || !flags[i]) return;
  const unsigned long long k = best[i];
  b200_corr c;
  c.index_query = (int)(unsigned)(k & 0xffffffffull);
  c.index_match = i;
  c.distance = __uint_as_float((unsigned)(k >> 32));
  out[slots[i]] = c;
}

}  // namespace

int match_tc_filter(b200_ctx *ctx, const float *d_model, int Km, const unsigned char *mvalid, const TcModelPrep *prep,
                    const float *d_scene, int Ks, const unsigned char *svalid, int D, int plan,
                    unsigned long long *best, int *zero_cnt, int *fb_rows, int *fb_count);
int match_tc_prepare_model(b200_ctx *ctx, const float *d_model, int Km, int D, const unsigned char *mvalid,
                           TcModelPrep *out);

// Smallest library / all-pairs volume for which the tensor-core filter pays off.
static bool tc_worthwhile(int Km, double Ks, int D) {
  return D >= 16 && D <= 1024 && Km >= 1024 && (double)Km * Ks * D >= 2.0e10;
}

// 0: exact float32 kernel only; 1 / 3: tcgen05 pre-filter with 1 / 3 fp16 terms + exact rescoring; 13: one term for
// every row, three terms for the rows the first pass cannot certify (a few per cent), then the exact kernel for
// what is left (a few rows).  B200_MATCH=exact|tc1|tc3|tc13 overrides the size heuristic (testing aid).
static int match_mode_for(int Km, int Ks, int D) {
  const char *e = getenv("B200_MATCH");
  if (e) {
    if (!strcmp(e, "exact")) return 0;
    if (D >= 16 && D <= 1024) {
      if (!strcmp(e, "tc1")) return 1;
      if (!strcmp(e, "tc3")) return 3;
      if (!strcmp(e, "tc13")) return 13;
    }
  }
  return tc_worthwhile(Km, (double)Ks, D) ? 13 : 0;
}

// Keeps the model side of the filter resident with the model (b200_model_create_shot, b200_library_*): validity
// flags, fp16 split operands, norms and scale are computed once instead of once per scene.
int match_prepare_rows(b200_ctx *ctx, const float *d_desc, int K, int D, TcModelPrep *t) {
  if (K < 1024 || D < 16 || D > 1024 || getenv("B200_MATCH_NOCACHE")) return B200_OK;
  B200_TRY(t->mvalid.alloc(ctx, (size_t)K));
  B200_TRY(t->nmv.alloc(ctx, 1));
  B200_TRY(t->nmv.zero());
  row_valid_kernel<<<ceil_div((long long)K * 32, 256), 256, 0, ctx->stream>>>(d_desc, K, D, 1, t->mvalid.p, t->nmv.p);
  B200_LAUNCHED(ctx);
  return match_tc_prepare_model(ctx, d_desc, K, D, t->mvalid.p, t);
}

int match_prepare_model(b200_ctx *ctx, b200_model *m) {
  if (!m) return B200_OK;
  return match_prepare_rows(ctx, m->desc.p, m->K, m->D, &m->tc);
}

// best[i] = (d2 bits << 32 | model row) of scene row i's exact float32 nearest valid model row, zero_cnt[i] = number
// of model rows at distance exactly 0; both must be initialised (init_best_kernel).  The tensor-core filter takes
// the bulk of the work for large libraries; the result is the exact kernel's, bit for bit.
static int match_best_rows(b200_ctx *ctx, const float *d_model, int Km, const unsigned char *mvalid_p,
                           const TcModelPrep *prep, const float *d_scene, int Ks, const unsigned char *svalid_p, int D,
                           unsigned long long *best_p, int *zero_cnt_p) {
  if (Km > 0) {
    const int sx = ceil_div(Ks, MT);
    const int mtiles = ceil_div(Km, MT);
    // enough CTAs for >= 4 waves when the scene is small
    int sy = std::max(1, std::min(mtiles, (ctx->sm_count * 8) / std::max(sx, 1)));
    dim3 grid(sx, sy);
    const int tc_terms = match_mode_for(Km, Ks, D);
    ctx->last_match_fallback = -1;
    if (tc_terms) {
      DevBuf<int> fb_rows, fb_count;
      B200_TRY(fb_rows.alloc(ctx, (size_t)Ks));
      B200_TRY(fb_count.alloc(ctx, 1));
      B200_TRY(match_tc_filter(ctx, d_model, Km, mvalid_p, prep, d_scene, Ks, svalid_p, D, tc_terms, best_p,
                               zero_cnt_p, fb_rows.p, fb_count.p));
      if (ctx->profiling) {  // bench statistic: how many rows needed the exact kernel
        B200_CUDA(ctx, cudaMemcpyAsync(&ctx->last_match_fallback, fb_count.p, sizeof(int), cudaMemcpyDeviceToHost,
                                       ctx->stream));
      }
      // uncertified rows: exact evaluation by persistent CTAs over (row block, model tile) items
      match_rows_kernel<<<ctx->sm_count * 4, MTHREADS, 0, ctx->stream>>>(d_model, Km, mvalid_p, d_scene, D, fb_rows.p,
                                                                        fb_count.p, best_p, zero_cnt_p);
      B200_LAUNCHED(ctx);
    } else {
      match_tile_kernel<<<grid, MTHREADS, 0, ctx->stream>>>(d_model, Km, mvalid_p, d_scene, Ks, D, best_p, zero_cnt_p);
      B200_LAUNCHED(ctx);
    }
  }
  return B200_OK;
}

__global__ void nearest1_emit_kernel(const unsigned long long *__restrict__ best, const unsigned char *__restrict__ qvalid,
                                     int nq, int *__restrict__ idx, float *__restrict__ d2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const unsigned long long k = best[i];
  const bool have = qvalid[i] && k != ~0ull;
  idx[i] = have ? (int)(unsigned)(k & 0xffffffffull) : -1;
  d2[i] = have ? __uint_as_float((unsigned)(k >> 32)) : __int_as_float(0x7f800000);
}

// Nearest valid model row of every query row (k = 1 of KdTreeFLANN::nearestKSearch for a batch of queries): the
// correspondence search without the acceptance test.  Queries with a non-finite value get -1 / +inf.
int dev_nearest1(b200_ctx *ctx, const float *d_model, int Km, const unsigned char *mvalid_p, const TcModelPrep *prep,
                 const float *d_q, int nq, int D, int *d_idx, float *d_d2) {
  if (nq <= 0) return B200_OK;
  if (prep && (!prep->ready || prep->Km != Km || prep->D != D)) prep = nullptr;
  DevBuf<unsigned char> qvalid;
  DevBuf<int> zero_cnt;
  DevBuf<unsigned long long> best;
  B200_TRY(qvalid.alloc(ctx, (size_t)nq));
  B200_TRY(zero_cnt.alloc(ctx, (size_t)nq));
  B200_TRY(best.alloc(ctx, (size_t)nq));
  row_valid_kernel<<<ceil_div((long long)nq * 32, 256), 256, 0, ctx->stream>>>(d_q, nq, D, 1, qvalid.p, nullptr);
  B200_LAUNCHED(ctx);
  init_best_kernel<<<ceil_div(nq, 256), 256, 0, ctx->stream>>>(best.p, zero_cnt.p, nq);
  B200_LAUNCHED(ctx);
  B200_TRY(match_best_rows(ctx, d_model, Km, mvalid_p, prep, d_q, nq, qvalid.p, D, best.p, zero_cnt.p));
  nearest1_emit_kernel<<<ceil_div(nq, 256), 256, 0, ctx->stream>>>(best.p, qvalid.p, nq, d_idx, d_d2);
  B200_LAUNCHED(ctx);
  return B200_OK;
}

int dev_match(b200_ctx *ctx, const float *d_model, int Km, const float *d_scene, int Ks, int D, int mode, float thr,
              b200_corr *d_out, int *d_count, const TcModelPrep *prep) {
  if (mode != 1 && mode != 2) return ctx->fail(B200_ERR_INVALID, "match: mode must be 1 or 2");
  if (D <= 0 || Km < 0 || Ks < 0) return ctx->fail(B200_ERR_INVALID, "match: bad sizes");
  B200_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(int), ctx->stream));
  if (Ks == 0) return B200_OK;
  StageScope st_(ctx, ST_MATCH);
  DevBuf<unsigned char> mvalid, svalid;
  DevBuf<int> nmv, zero_cnt, flags, slots;
  DevBuf<unsigned long long> best;
  B200_TRY(mvalid.alloc(ctx, (size_t)std::max(Km, 1)));
  B200_TRY(svalid.alloc(ctx, (size_t)Ks));
  B200_TRY(nmv.alloc(ctx, 1));
  B200_TRY(nmv.zero());
  B200_TRY(zero_cnt.alloc(ctx, (size_t)Ks));
  B200_TRY(flags.alloc(ctx, (size_t)Ks));
  B200_TRY(slots.alloc(ctx, (size_t)Ks));
  B200_TRY(best.alloc(ctx, (size_t)Ks));
  if (prep && (!prep->ready || prep->Km != Km || prep->D != D)) prep = nullptr;
  const unsigned char *mvalid_p = prep ? prep->mvalid.p : mvalid.p;
  const int *nmv_p = prep ? prep->nmv.p : nmv.p;
  if (Km > 0 && !prep) {
    row_valid_kernel<<<ceil_div((long long)Km * 32, 256), 256, 0, ctx->stream>>>(d_model, Km, D, 1, mvalid.p, nmv.p);
    B200_LAUNCHED(ctx);
  }
  row_valid_kernel<<<ceil_div((long long)Ks * 32, 256), 256, 0, ctx->stream>>>(d_scene, Ks, D, 0, svalid.p, nullptr);
  B200_LAUNCHED(ctx);
  init_best_kernel<<<ceil_div(Ks, 256), 256, 0, ctx->stream>>>(best.p, zero_cnt.p, Ks);
  B200_LAUNCHED(ctx);
  B200_TRY(match_best_rows(ctx, d_model, Km, mvalid_p, prep, d_scene, Ks, svalid.p, D, best.p, zero_cnt.p));
  match_flags_kernel<<<ceil_div(Ks, 256), 256, 0, ctx->stream>>>(best.p, zero_cnt.p, svalid.p, nmv_p, Ks, mode, thr,
                                                                 flags.p);
  B200_LAUNCHED(ctx);
  B200_TRY(exclusive_scan_i32(ctx, flags.p, slots.p, Ks, d_count));
  match_emit_kernel<<<ceil_div(Ks, 256), 256, 0, ctx->stream>>>(best.p, flags.p, slots.p, Ks, d_out);
  B200_LAUNCHED(ctx);
  return B200_OK;
}
