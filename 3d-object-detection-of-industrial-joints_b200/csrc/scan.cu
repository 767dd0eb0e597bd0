// scan.cu — exclusive prefix sum of int32 (cell counts → cell starts, flags → compaction slots).
// Three launches: per-tile sums, scan of the tile sums (one CTA), per-tile scan + offset.
// HBM-bound: reads n ints twice, writes n ints once.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int &block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) s_warp[32] = wi;
  }
  __syncthreads();
  block_total = s_warp[32];
  int res = incl - v + s_warp[warp];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) tile_sums_kernel(const int *__restrict__ in, int n,
                                                                 int *__restrict__ tile_sums) {
  __shared__ int s_warp[33];
  const long long base = (long long)blockIdx.x * SCAN_TILE;
  int v = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    long long j = base + (long long)i * SCAN_THREADS + threadIdx.x;
    if (j < n) v += in[j];
  }
  int total;
  block_exclusive_scan(v, s_warp, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(int *tile_sums, int ntiles, int *total_out) {
  __shared__ int s_warp[33];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < ntiles; base += 1024) {
    int j = base + threadIdx.x;
    int v = (j < ntiles) ? tile_sums[j] : 0;
    int total;
    int ex = block_exclusive_scan(v, s_warp, total);
    int carry = s_carry;
    if (j < ntiles) tile_sums[j] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = s_carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) tile_scan_kernel(const int *__restrict__ in, int n,
                                                                 const int *__restrict__ tile_offsets,
                                                                 int *__restrict__ out) {
  __shared__ int s_warp[33];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    long long j = base + i;
    v[i] = (j < n) ? in[j] : 0;
    sum += v[i];
  }
  int total;
  int ex = block_exclusive_scan(sum, s_warp, total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    long long j = base + i;
    if (j < n) out[j] = ex;
    ex += v[i];
  }
}

}  // namespace

// d_out may alias d_in.  If d_total is given it receives the sum of all n inputs.
int exclusive_scan_i32(b200_ctx *ctx, const int *d_in, int *d_out, int n, int *d_total) {
  if (n <= 0) {
    if (d_total) B200_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(int), ctx->stream));
    return B200_OK;
  }
  const int ntiles = ceil_div(n, SCAN_TILE);
  DevBuf<int> sums;
  B200_TRY(sums.alloc(ctx, ntiles));
  tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, sums.p);
  B200_LAUNCHED(ctx);
  scan_tile_sums_kernel<<<1, 1024, 0, ctx->stream>>>(sums.p, ntiles, d_total);
  B200_LAUNCHED(ctx);
  tile_scan_kernel<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, sums.p, d_out);
  B200_LAUNCHED(ctx);
  return B200_OK;
}
