// desc_index.cu — resident descriptor index for per-query nearest-neighbour calls.
//
// Drop-in for the way the reference drives pcl::KdTreeFLANN<SHOT352 / FPFHSignature33>:
// setInputCloud(model_descriptors) once (SHOT.cpp:405-406, SHOT_demo.cpp:508-509, FPFH_demo.cpp:516-517),
// then nearestKSearch(scene_descriptor, k, indices, sqr_dists) per scene descriptor inside a user loop
// (SHOT.cpp:417, k = 1; SHOT_demo.cpp:521, k = 2).  The model rows stay in HBM; a call evaluates the
// exact float32 L2_Simple distance of each query to every finite model row (one CTA per query, one
// sequential accumulation chain per pair, so distances are bit-identical to FLANN's) and returns the k
// smallest in (distance, index) order.  The batched b200_match is the fast path for whole clouds; this
// entry point exists so the reference's loop runs unchanged.
#include <algorithm>

#include "common.cuh"

struct b200_desc_index {
  b200_ctx *ctx = nullptr;
  int K = 0, D = 0, n_valid = 0;
  DevBuf<float> desc;
  DevBuf<unsigned char> valid;
  TcModelPrep tc;  // tensor-core filter operands (large indices): batched k = 1 queries take the matching path
};

namespace {

constexpr int DK_THREADS = 128;
constexpr int DK_MAXK = 16;

__global__ void desc_valid_kernel(const float *__restrict__ desc, int rows, int D, unsigned char *__restrict__ valid,
                                  int *__restrict__ n_valid) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows) return;
  bool ok = true;
  for (int d = lane; d < D; d += 32) ok = ok && isfinite(desc[(size_t)w * D + d]);
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    valid[w] = ok ? 1 : 0;
    if (ok) atomicAdd(n_valid, 1);
  }
}

__global__ void __launch_bounds__(DK_THREADS)
    desc_knn_kernel(const float *__restrict__ model, const unsigned char *__restrict__ valid, int Km, int D,
                    const float *__restrict__ queries, int nq, int k, int *__restrict__ idx_out,
                    float *__restrict__ d2_out) {
  extern __shared__ float s_q[];  // D floats
  __shared__ unsigned long long s_red[DK_THREADS / 32];
  __shared__ unsigned long long s_win;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int qi = blockIdx.x; qi < nq; qi += gridDim.x) {
    for (int d = tid; d < D; d += DK_THREADS) s_q[d] = queries[(size_t)qi * D + d];
    __syncthreads();
    // per-thread sorted list of the k best (d2 bits << 32 | index)
    unsigned long long best[DK_MAXK];
#pragma unroll
    for (int t = 0; t < DK_MAXK; ++t) best[t] = ~0ull;
    for (int j = tid; j < Km; j += DK_THREADS) {
      if (!valid[j]) continue;
      const float *b = model + (size_t)j * D;
      float acc = 0.0f;
      for (int d = 0; d < D; ++d) {
        const float diff = s_q[d] - b[d];
        acc = __fadd_rn(acc, __fmul_rn(diff, diff));
      }
      unsigned long long key = ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned)j;
      if (key < best[DK_MAXK - 1]) {
        // insertion into the (fully unrolled, register resident) sorted list
#pragma unroll
        for (int t = 0; t < DK_MAXK; ++t) {
          if (key < best[t]) {
            const unsigned long long tmp = best[t];
            best[t] = key;
            key = tmp;
          }
        }
      }
    }
    // k rounds of block-wide arg-min over the list heads
    int head = 0;
    for (int r = 0; r < k; ++r) {
      unsigned long long mine = ~0ull;
#pragma unroll
      for (int t = 0; t < DK_MAXK; ++t)
        if (t == head) mine = best[t];
      unsigned long long m = mine;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ov = __shfl_xor_sync(0xffffffffu, m, o);
        m = (ov < m) ? ov : m;
      }
      if (lane == 0) s_red[warp] = m;
      __syncthreads();
      if (tid == 0) {
        unsigned long long w = s_red[0];
        for (int x = 1; x < DK_THREADS / 32; ++x) w = (s_red[x] < w) ? s_red[x] : w;
        s_win = w;
        const bool have = (w != ~0ull);
        idx_out[(size_t)qi * k + r] = have ? (int)(unsigned)(w & 0xffffffffull) : -1;
        d2_out[(size_t)qi * k + r] = have ? __uint_as_float((unsigned)(w >> 32)) : __int_as_float(0x7f800000);
      }
      __syncthreads();
      if (mine == s_win && mine != ~0ull) ++head;  // keys are unique (they carry the row index)
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" {

int b200_desc_index_create(b200_ctx *ctx, const float *desc, int K, int D, b200_desc_index **out) {
  if (!ctx) return B200_ERR_INVALID;
  if (!out || K < 0 || D <= 0 || (K > 0 && !desc)) return ctx->fail(B200_ERR_INVALID, "desc_index_create: bad arguments");
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return ctx->fail_cuda(e, "cudaSetDevice", __FILE__, __LINE__);
  b200_desc_index *ix = new b200_desc_index();
  ix->ctx = ctx;
  ix->K = K;
  ix->D = D;
  int rc = B200_OK;
  DevBuf<int> nv;
  do {
    if ((rc = ix->desc.alloc(ctx, (size_t)std::max(K, 1) * D)) != B200_OK) break;
    if ((rc = ix->valid.alloc(ctx, (size_t)std::max(K, 1))) != B200_OK) break;
    if ((rc = nv.alloc(ctx, 1)) != B200_OK) break;
    if ((rc = nv.zero()) != B200_OK) break;
    if (K > 0) {
      e = cudaMemcpyAsync(ix->desc.p, desc, (size_t)K * D * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) {
        rc = ctx->fail_cuda(e, "H2D descriptors", __FILE__, __LINE__);
        break;
      }
      desc_valid_kernel<<<ceil_div((long long)K * 32, 256), 256, 0, ctx->stream>>>(ix->desc.p, K, D, ix->valid.p, nv.p);
      ctx->launches++;
      if ((rc = match_prepare_rows(ctx, ix->desc.p, K, D, &ix->tc)) != B200_OK) break;
    }
    e = cudaMemcpyAsync(&ix->n_valid, nv.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = ctx->sync();
    if (e != cudaSuccess) rc = ctx->fail_cuda(e, "desc_index_create sync", __FILE__, __LINE__);
  } while (0);
  if (rc != B200_OK) {
    delete ix;
    return rc;
  }
  *out = ix;
  return B200_OK;
}

int b200_desc_index_destroy(b200_desc_index *ix) {
  if (!ix) return B200_OK;
  cudaSetDevice(ix->ctx->device);
  delete ix;
  return B200_OK;
}

int b200_desc_index_size(const b200_desc_index *ix) { return ix ? ix->n_valid : 0; }

int b200_desc_index_knn(b200_ctx *ctx, const b200_desc_index *ix, const float *queries, int nq, int k, int *idx,
                        float *d2, int *k_found) {
  if (!ctx) return B200_ERR_INVALID;
  if (!ix || nq < 0 || k < 1 || k > DK_MAXK || (nq > 0 && (!queries || !idx || !d2)))
    return ctx->fail(B200_ERR_INVALID, "desc_index_knn: bad arguments (1 <= k <= 16)");
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return ctx->fail_cuda(e, "cudaSetDevice", __FILE__, __LINE__);
  if (k_found) *k_found = std::min(k, ix->n_valid);
  if (nq == 0) return B200_OK;
  const int D = ix->D;
  DevBuf<float> dq, dd2;
  DevBuf<int> didx;
  B200_TRY(dq.alloc(ctx, (size_t)nq * D));
  B200_TRY(dd2.alloc(ctx, (size_t)nq * k));
  B200_TRY(didx.alloc(ctx, (size_t)nq * k));
  B200_CUDA(ctx, cudaMemcpyAsync(dq.p, queries, (size_t)nq * D * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  if (k == 1 && nq >= 32) {
    // a batch of k = 1 queries (the look-ahead of the PCL-style adapter, or a caller's own batch) is the correspondence
    // search itself: tensor-core filter + exact rescoring, same answers as the per-query kernel below
    B200_TRY(dev_nearest1(ctx, ix->desc.p, ix->K, ix->valid.p, &ix->tc, dq.p, nq, D, didx.p, dd2.p));
    B200_CUDA(ctx, cudaMemcpyAsync(idx, didx.p, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(ctx, cudaMemcpyAsync(d2, dd2.p, (size_t)nq * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(ctx, ctx->sync());
    return B200_OK;
  }
  const size_t smem = (size_t)D * sizeof(float);
  if (smem > 48 * 1024)
    B200_CUDA(ctx, ensure_dyn_smem(desc_knn_kernel, smem));
  desc_knn_kernel<<<std::min(nq, ctx->sm_count * 8), DK_THREADS, smem, ctx->stream>>>(ix->desc.p, ix->valid.p, ix->K, D,
                                                                                     dq.p, nq, k, didx.p, dd2.p);
  B200_LAUNCHED(ctx);
  B200_CUDA(ctx, cudaMemcpyAsync(idx, didx.p, (size_t)nq * k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, cudaMemcpyAsync(d2, dd2.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(ctx, ctx->sync());
  return B200_OK;
}

} /* extern "C" */
