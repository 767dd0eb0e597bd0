// match_tc.cu — tensor-core pre-filter for the descriptor correspondence search.
//
// Same contract as match.cu (KdTreeFLANN<SHOT352>::nearestKSearch + the user threshold loop,
// SHOT.cpp:405-423, SHOT_demo.cpp:508-530): the answer is the exact float32 L2_Simple arg-min.  For a
// large model library the all-pairs float32 evaluation is FP32-pipe bound, so the bulk of it moves to
// the 5th-generation tensor cores:
//
//   d2(i,j) = |a_i|^2 + |b_j|^2 - 2 a_i.b_j ; for a fixed scene row i only s_ij = |b_j|^2 - 2 a_i.b_j
//   matters.  a.b is evaluated as a dense contraction on tcgen05 (fp16 operands, fp32 accumulation in
//   TMEM).  Each fp32 value is split into fp16 terms x = x1 + x2 (x1 = rn16(x), x2 = rn16(x - x1)) and
//   the contraction runs over the concatenated K' = [a1|a1|a2] . [b1|b2|b1] (3 terms, ~2^-22 relative
//   operand error) or just a1.b1 (1 term, 2^-11), after an exact power-of-two rescale into fp16 range.
//   The epilogue (16 warps: one TMEM lane = one scene row per thread, four column parts per row) keeps the
//   TC_CAND smallest s_ij of the row in registers across all model tiles; the parts' lists are merged per row.
//   A second kernel rescoring those candidates with the exact sequential float32 distance then
//   CERTIFIES the arg-min: every non-candidate has approximate s >= the TC_CAND-th smallest s_c, so if the best
//   exact distance is below |a_i|^2 + s_c - eps (eps = proven error bound of the approximation) no
//   other row can win or tie.  Rows that cannot be certified are re-evaluated exactly by
//   match_rows_kernel.  The result is therefore bit-identical to the exact path.
//   TC_CAND = 4: with 32 independent rows per warp some lane inserts at almost every 32-column chunk, and
//   the insertion cost made the epilogue co-limiting at 8 (1.78 ms, tensor pipe 73 % active; 1.45 ms and
//   88 % at 4; a handful of rows per 100 k then miss the certificate and take the exact path).
//
// Kernel structure (one CTA per SM, persistent over (scene tile, model split) work items):
//   warp 0   TMA producer: 4-stage ring of {A 128x64, B 256x64} fp16 tiles, SWIZZLE_128B
//   warp 1   TMEM allocator (512 columns = two 128x256 fp32 accumulators) + single-thread MMA issuer:
//            tcgen05.mma.cta_group::1.kind::f16, M=128 N=256 K=16, smem descriptors, commit → mbarrier
//   warps 2-17 epilogue (four per TMEM lane quarter, a quarter of the columns each): tcgen05.ld 32x32b.x32 →
//            s = nb_j - 2*acc → sorted insert into the row's candidate list; overlaps the next tile's MMAs
//            through the second accumulator
// Roofline: tensor pipe; 2*Ks*Km*K' flop per call.
#include <stdint.h>
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 64;  // fp16 elements per k-block = 128 bytes = one swizzle row
constexpr int TC_STAGES = 4;
constexpr int TC_CAND = 4;
constexpr int TC_MAX_SPLIT = 4;
constexpr int TC_EPI_WARPS = 16;  // TC_PARTS warps per TMEM lane quarter, each handles 1/TC_PARTS of the tile's columns
constexpr int TC_PARTS = TC_EPI_WARPS / 4;
constexpr int TC_EPI_THREADS = 32 * TC_EPI_WARPS;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr uint32_t TC_B_BYTES = TC_BN * TC_BK * 2;
constexpr uint32_t TC_SMEM = TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + 1024 /*align*/ + 256 /*barriers*/ +
                             2 * TC_BN * 4 /*|b|^2 of the two accumulators' model tiles*/ +
                             (TC_PARTS - 1) * TC_BM * TC_CAND * 8 /*candidate hand-over between the column parts*/ +
                             (TC_PARTS - 1) * TC_BM * 4 /*... and their non-candidate bounds*/;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni LAB_DONE;\n"
      "bra.uni LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, rows of 128 bytes, SWIZZLE_128B: 8-row groups are 1024 bytes apart (SBO).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset
  d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}

// 32 lanes x 32 columns of fp32 accumulators → 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------- operand preparation
__global__ void absmax_kernel(const float *__restrict__ x, size_t n, unsigned *__restrict__ out_bits) {
  float m = 0.f;
  auto take = [&](float v) {
    v = fabsf(v);
    if (isfinite(v)) m = fmaxf(m, v);
  };
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // 16-byte loads, four in flight per thread (scalar loads reached 2.2 TB/s: 57 us for the 128 MB of a scene)
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const size_t n4 = n >> 2;
    for (size_t i = tid; i < n4; i += 4 * nthr) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * nthr < n4) v[u] = __ldg(x4 + i + u * nthr);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * nthr < n4) take(v[u].x), take(v[u].y), take(v[u].z), take(v[u].w);
    }
    for (size_t i = (n4 << 2) + tid; i < n; i += nthr) take(x[i]);
  } else {
    for (size_t i = tid; i < n; i += nthr) take(x[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));  // non-negative floats order as uints
}

// scale[0] = 2^(TC_SCALE_EXP - e): absmax * scale lies in [2^(TC_SCALE_EXP-1), 2^TC_SCALE_EXP); scale[1] = absmax.
// The operands sit high in the fp16 range so that the second split term x2 = rn16(x - x1) (about 2^-11 |x|) is a
// normal fp16 number for every element within 2^-13 of the largest one; products and the K'-term fp32 sums stay
// far below the fp32 range (2^22 * K').
constexpr int TC_SCALE_EXP = 11;
__global__ void make_scale_kernel(const unsigned *__restrict__ absmax_bits, float *__restrict__ scale) {
  const float m = __uint_as_float(*absmax_bits);
  int e = 0;
  if (m > 0.f) frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)
  scale[0] = ldexpf(1.0f, TC_SCALE_EXP - e);
  scale[1] = m;
}

// One warp per row.  is_b = 0: row = [x1 | x1 | x2], is_b = 1: row = [x1 | x2 | x1] (terms == 3);
// terms == 1: row = [x1].  Rows are zero padded to Kp halves.  norm2[row] = float32 sum of squares of
// the ORIGINAL values (3e38 for rows that are not valid).
// row_map / n_rows_dev (both or neither): output row w is input row row_map[w], for w < *n_rows_dev only (the rows a
// first pass could not certify; their number is known on the device only — rows beyond it are left untouched).
__global__ void tc_prep_kernel(const float *__restrict__ X, int rows, int rows_padded, int D, int Kp, int terms,
                               int is_b, const float *__restrict__ scale, const unsigned char *__restrict__ valid,
                               __half *__restrict__ out, float *__restrict__ norm2, unsigned *__restrict__ normmax_bits,
                               const int *__restrict__ row_map, const int *__restrict__ n_rows_dev, int perm_tiles) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows_padded) return;
  int src = w;
  if (row_map) {
    if (w >= *n_rows_dev) return;
    src = row_map[w];
  }
  // model side: operand row w holds model row tc_unpermute(w) — consecutive operand rows (the 32-column chunks of the
  // filter's epilogue) are 256 model rows apart.  Keypoints that are neighbours in space are neighbours in index and
  // have similar descriptors: without this a row's best and runner-up often share a chunk, where only one of them
  // can become a candidate.
  if (perm_tiles) src = (w % perm_tiles) * TC_BN + w / perm_tiles;
  __half *o = out + (size_t)w * Kp;
  const bool ok = (src < rows) && (valid == nullptr || valid[src]);
  if (!ok) {
    for (int d = lane; d < Kp; d += 32) o[d] = __float2half_rn(0.f);
    if (lane == 0) norm2[w] = 3.0e38f;  // "never the minimum", and finite: the filter's epilogue packs the column index
                                        // into the value's low mantissa bits, which would turn +inf into a NaN
    return;
  }
  const float sc = scale[0];
  float n2 = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float x = X[(size_t)src * D + d];
    n2 += x * x;
    const float xs = x * sc;
    const __half h1 = __float2half_rn(xs);
    if (terms == 3) {
      const __half h2 = __float2half_rn(xs - __half2float(h1));
      o[d] = h1;
      o[D + d] = is_b ? h2 : h1;
      o[2 * D + d] = is_b ? h1 : h2;
    } else {
      o[d] = h1;
    }
  }
  for (int d = terms * D + lane; d < Kp; d += 32) o[d] = __float2half_rn(0.f);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, off);
  if (lane == 0) {
    norm2[w] = n2;
    if (normmax_bits && isfinite(n2)) atomicMax(normmax_bits, __float_as_uint(n2));
  }
}

// ---------------------------------------------------------------- the tensor-core filter
struct TcParams {
  int Ks, Km;
  int m_tiles, n_tiles, n_split;
  int k_blocks;
  const float *nb;     // |b_j|^2, padded to n_tiles * TC_BN (3e38 for padding and invalid rows)
  const float *scaleA;
  const float *scaleB;
  float *cand_s;       // [Ks][n_split * TC_CAND]
  int *cand_j;
  float *cand_b;       // [Ks][n_split]: lower bound of every approximate value that is not a chunk minimum
  const int *rows_dev;  // nullable: the number of scene rows lives on the device (second pass), Ks is the capacity
};

__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_filter_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char *sA = smem;
  unsigned char *sB = smem + TC_STAGES * TC_A_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TC_STAGES * (TC_A_BYTES + TC_B_BYTES));
  uint64_t *full = bars;                 // [TC_STAGES]
  uint64_t *empty = bars + TC_STAGES;    // [TC_STAGES]
  uint64_t *tfull = bars + 2 * TC_STAGES;       // [2]
  uint64_t *tempty = bars + 2 * TC_STAGES + 2;  // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * TC_STAGES + 4);
  float *s_nb = reinterpret_cast<float *>(smem + TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + 256);  // [2][TC_BN]
  float *s_hs = s_nb + 2 * TC_BN;  // [TC_PARTS - 1][TC_BM][TC_CAND] candidates of the column parts 1..
  int *s_hj = reinterpret_cast<int *>(s_hs + (TC_PARTS - 1) * TC_BM * TC_CAND);
  float *s_hb = reinterpret_cast<float *>(s_hj + (TC_PARTS - 1) * TC_BM * TC_CAND);  // [TC_PARTS - 1][TC_BM]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 32 * TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int Ks_eff = p.rows_dev ? min(*p.rows_dev, p.Ks) : p.Ks;
  const int n_items = ((Ks_eff + TC_BM - 1) / TC_BM) * p.n_split;
  const int tiles_per_split = (p.n_tiles + p.n_split - 1) / p.n_split;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int mt = item / p.n_split, sp = item % p.n_split;
        const int nt0 = sp * tiles_per_split, nt1 = min(p.n_tiles, nt0 + tiles_per_split);
        for (int nt = nt0; nt < nt1; ++nt)
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1u);
            mbar_expect_tx(&full[stage], TC_A_BYTES + TC_B_BYTES);
            tma_load_2d(sA + stage * TC_A_BYTES, &mapA, &full[stage], kb * TC_BK, mt * TC_BM);
            tma_load_2d(sB + stage * TC_B_BYTES, &mapB, &full[stage], kb * TC_BK, nt * TC_BN);
            if (++stage == TC_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D = f32, A = B = f16, both K-major, N = 256, M = 128
      const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int sp = item % p.n_split;
        const int nt0 = sp * tiles_per_split, nt1 = min(p.n_tiles, nt0 + tiles_per_split);
        for (int nt = nt0; nt < nt1; ++nt) {
          mbar_wait(&tempty[acc], acc_phase ^ 1u);  // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_BN);
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t adesc = make_sdesc(smem_u32(sA + stage * TC_A_BYTES));
            const uint64_t bdesc = make_sdesc(smem_u32(sB + stage * TC_B_BYTES));
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            umma_commit(&empty[stage]);  // smem slot free when these MMAs retire
            if (++stage == TC_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit(&tfull[acc]);  // accumulator complete
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..) =====================
    // warps 2-5 take the first TC_BN / TC_PARTS columns of every accumulator tile, warps 6-9 the next, ...;
    // a row's candidate lists are merged once per work item
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // column part
    const int col_lo = half * (TC_BN / TC_PARTS);
    const int row_in_tile = quarter * 32 + lane;
    const float inv = 1.0f / (p.scaleA[0] * p.scaleB[0]);
    const float m2inv = -2.0f * inv;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int mt = item / p.n_split, sp = item % p.n_split;
      const int nt0 = sp * tiles_per_split, nt1 = min(p.n_tiles, nt0 + tiles_per_split);
      float cs[TC_CAND];
      int cj[TC_CAND];
      float bnd = __int_as_float(0x7f800000);  // smallest second-in-chunk value seen
#pragma unroll
      for (int t = 0; t < TC_CAND; ++t) {
        cs[t] = __int_as_float(0x7f800000);
        cj[t] = -1;
      }
      for (int nt = nt0; nt < nt1; ++nt) {
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TC_BN);
        const int n0 = nt * TC_BN;
        // |b_j|^2 of this model tile → shared memory (every epilogue thread needs all 256 values)
        float *snb = s_nb + acc * TC_BN;
        {
          const int t = threadIdx.x - 64;
          if (t < TC_BN) snb[t] = __ldg(&p.nb[n0 + t]);
          asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
        }
#pragma unroll 1
        for (int c0 = col_lo; c0 < col_lo + TC_BN / TC_PARTS; c0 += 32) {
          float v[32];
          tmem_ld32(t_addr + (uint32_t)c0, v);
          // s = |b|^2 - 2 a.b for the chunk's 32 model rows.  Only the chunk's MINIMUM competes for the row's
          // candidate list (one insertion attempt per chunk instead of one per element: with 32 independent rows per
          // warp some lane inserted at almost every element and the whole warp walked the insertion code); the chunk's
          // SECOND smallest value bounds everything else in the chunk from below and goes into the row's
          // non-candidate bound.  The column (5 bits) rides in the low mantissa bits of the value, so minimum and
          // arg-min are one FMNMX chain; the perturbation (<= 31 ulp) is part of the certificate's error budget.
          float m1 = __int_as_float(0x7f800000), m2 = m1;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 nb4 = *reinterpret_cast<const float4 *>(snb + c0 + e);
            const float nbv[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float x = fmaf(m2inv, v[e + q], nbv[q]);  // padding / invalid model rows carry |b|^2 = 3e38
              const float xb = __uint_as_float((__float_as_uint(x) & ~31u) | (unsigned)(e + q));
              m2 = fminf(m2, fmaxf(m1, xb));
              m1 = fminf(m1, xb);
            }
          }
          bnd = fminf(bnd, m2);
          if (m1 < cs[TC_CAND - 1]) {
            cs[TC_CAND - 1] = m1;
            cj[TC_CAND - 1] = n0 + c0 + (int)(__float_as_uint(m1) & 31u);
#pragma unroll
            for (int t = TC_CAND - 1; t > 0; --t)
              if (cs[t] < cs[t - 1]) {
                const float ts = cs[t];
                cs[t] = cs[t - 1];
                cs[t - 1] = ts;
                const int tj = cj[t];
                cj[t] = cj[t - 1];
                cj[t - 1] = tj;
              }
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      // hand the other parts' candidates to part 0 and merge (ascending lists)
      if (half > 0) {
#pragma unroll
        for (int t = 0; t < TC_CAND; ++t) {
          s_hs[((half - 1) * TC_BM + row_in_tile) * TC_CAND + t] = cs[t];
          s_hj[((half - 1) * TC_BM + row_in_tile) * TC_CAND + t] = cj[t];
        }
        s_hb[(half - 1) * TC_BM + row_in_tile] = bnd;
      }
      asm volatile("bar.sync 2, %0;" ::"n"(TC_EPI_THREADS) : "memory");
      if (half == 0) {
#pragma unroll
        for (int u = 0; u < TC_PARTS - 1; ++u) bnd = fminf(bnd, s_hb[u * TC_BM + row_in_tile]);
#pragma unroll 1
        for (int u = 0; u < (TC_PARTS - 1) * TC_CAND; ++u) {
          const int o = ((u / TC_CAND) * TC_BM + row_in_tile) * TC_CAND + (u % TC_CAND);
          const float s = s_hs[o];
          if (s < cs[TC_CAND - 1]) {
            cs[TC_CAND - 1] = s;
            cj[TC_CAND - 1] = s_hj[o];
#pragma unroll
            for (int t = TC_CAND - 1; t > 0; --t)
              if (cs[t] < cs[t - 1]) {
                const float ts = cs[t];
                cs[t] = cs[t - 1];
                cs[t - 1] = ts;
                const int tj = cj[t];
                cj[t] = cj[t - 1];
                cj[t - 1] = tj;
              }
          }
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(TC_EPI_THREADS) : "memory");  // the hand-over buffer is free again
      const int row = mt * TC_BM + row_in_tile;
      if (half == 0 && row < Ks_eff) {
        float *os = p.cand_s + ((size_t)row * p.n_split + sp) * TC_CAND;
        int *oj = p.cand_j + ((size_t)row * p.n_split + sp) * TC_CAND;
#pragma unroll
        for (int t = 0; t < TC_CAND; ++t) {
          os[t] = cs[t];
          oj[t] = cj[t];
        }
        p.cand_b[(size_t)row * p.n_split + sp] = bnd;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------- exact rescoring + certificate
// A warp handles 32 / nc scene rows at a time (nc = n_split * TC_CAND candidates per row: 8, 16 or 32),
// one (row, candidate) pair per lane.  The descriptor rows are staged through shared memory in
// 64-column chunks with coalesced loads (the candidate rows are scattered over the library); each lane
// then adds its 64 squared differences in column order, so the distance is the exact sequential
// float32 L2_Simple value.
constexpr int RS_WARPS_PER_CTA = 8;

__global__ void __launch_bounds__(RS_WARPS_PER_CTA * 32)
    tc_rescore_kernel(const float *__restrict__ model, int Km, const float *__restrict__ scene, int Ks, int D,
                      const unsigned char *__restrict__ svalid, int n_split, const float *__restrict__ cand_s,
                      const int *__restrict__ cand_j, const float *__restrict__ cand_b, const float *__restrict__ na,
                      const float *__restrict__ normmaxB, const float *__restrict__ scaleA,
                      const float *__restrict__ scaleB, float eta, unsigned long long *__restrict__ best,
                      int *__restrict__ zero_cnt, int *__restrict__ fb_rows, int *__restrict__ fb_count,
                      unsigned *__restrict__ err_ratio_bits, const int *__restrict__ row_map,
                      const int *__restrict__ n_rows_dev, int n_tiles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nc = n_split * TC_CAND;  // 4, 8 or 16
  const int rpw = 32 / nc;           // rows per warp
  const int gw = blockIdx.x * RS_WARPS_PER_CTA + warp;
  const int r = lane / nc, c = lane % nc;  // this lane's (row, candidate)
  // ci: row of the filter's operand / candidate arrays; i: the scene row it stands for
  const int n_rows = row_map ? min(*n_rows_dev, Ks) : Ks;
  if (gw * rpw >= n_rows) return;
  const int ci = gw * rpw + r;
  const int i = (ci < n_rows) ? (row_map ? row_map[ci] : ci) : -1;
  const bool row_ok = i >= 0 && svalid[i];  // rows skipped by the caller's flags (pcl_isfinite(descriptor[0]))
  // float32 error of a D-term sequential sum / a warp-tree norm, relative to the sum of magnitudes
  const float c_sum = (float)(D + 8) * 1.1920929e-7f;
  // fp16 flush of tiny elements: absolute 2^-25 per element of the scaled operand (2^-(25+10) of the largest
  // element), against a vector of at most sqrt(D) |b| in the 1-norm; both operands, times two for -2 a.b
  const float c_sub = 1.2e-10f * sqrtf((float)D);
  // the filter keeps the column index in the 5 low mantissa bits of s: <= 31 ulp of |s| <= |b|^2 + 2 |a||b|
  const float c_pay = 4e-6f;
  int j = -1;
  float s = __int_as_float(0x7f800000);
  if (row_ok) {
    j = cand_j[(size_t)ci * nc + c];
    if (j >= 0) j = (j % n_tiles) * TC_BN + j / n_tiles;  // operand row -> model row (see tc_prep_kernel)
    s = cand_s[(size_t)ci * nc + c];
  }
  bool pair_ok = row_ok && j >= 0 && j < Km;
  // s_cut: every model row that is not a candidate has approximate s >= the worst kept value of its split
  float s_cut = __int_as_float(0x7f800000);
  // ... a value that was its 32-column chunk's minimum but not kept is >= the split's worst kept one; any other value
  // is >= its chunk's second smallest, whose minimum over the split the filter recorded (cand_b)
  if (row_ok && (c % TC_CAND) == TC_CAND - 1) s_cut = fminf(s, cand_b[(size_t)ci * n_split + c / TC_CAND]);
  for (int o = nc >> 1; o > 0; o >>= 1) s_cut = fminf(s_cut, __shfl_xor_sync(0xffffffffu, s_cut, o));

  // Prune: a candidate whose approximate value exceeds the row's smallest one by more than twice the error bound
  // cannot have the smallest float32 distance (nor tie with it), so its exact distance is never needed.  Typically
  // a handful of the nc candidates survive; their rows are the only ones staged below.
  {
    float s_min = pair_ok ? s : __int_as_float(0x7f800000);
    for (int o = nc >> 1; o > 0; o >>= 1) s_min = fminf(s_min, __shfl_xor_sync(0xffffffffu, s_min, o));
    if (pair_ok) {
      const float nai = na[ci], nbm = normmaxB[0];
      const float d_lo = fmaxf(nai + s_min, 0.0f);
      const float e = 2.0f * eta * sqrtf(nai * nbm) + c_sum * (nai + nbm + d_lo + 4.0f * eta * sqrtf(nai * nbm)) +
                      c_sub * (scaleA[1] * sqrtf(nbm) + scaleB[1] * sqrtf(nai)) +
                      c_pay * (nbm + 2.0f * sqrtf(nai * nbm)) + 1e-30f;
      pair_ok = s <= s_min + 2.0f * e;
    }
  }

  // Exact distance of this lane's surviving (row, candidate) pair: the two descriptor rows are read straight from
  // global memory with independent 16-byte loads (several in flight per lane; the lanes of a warp touch a handful of
  // rows, which stay in L1) and summed in column order — FLANN's sequential float32 L2_Simple value.
  float acc = 0.0f;
  if (pair_ok) {
    const float *a = scene + (size_t)i * D, *b = model + (size_t)j * D;
    if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0) {
      const float4 *a4 = reinterpret_cast<const float4 *>(a), *b4 = reinterpret_cast<const float4 *>(b);
      const int n4 = D >> 2;
      int d = 0;
      for (; d + 4 <= n4; d += 4) {
        float4 x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          x[u] = __ldg(a4 + d + u);
          y[u] = __ldg(b4 + d + u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float diff = x[u].x - y[u].x;
          acc = __fadd_rn(acc, __fmul_rn(diff, diff));
          diff = x[u].y - y[u].y;
          acc = __fadd_rn(acc, __fmul_rn(diff, diff));
          diff = x[u].z - y[u].z;
          acc = __fadd_rn(acc, __fmul_rn(diff, diff));
          diff = x[u].w - y[u].w;
          acc = __fadd_rn(acc, __fmul_rn(diff, diff));
        }
      }
      for (; d < n4; ++d) {
        const float4 x = __ldg(a4 + d), y = __ldg(b4 + d);
        float diff = x.x - y.x;
        acc = __fadd_rn(acc, __fmul_rn(diff, diff));
        diff = x.y - y.y;
        acc = __fadd_rn(acc, __fmul_rn(diff, diff));
        diff = x.z - y.z;
        acc = __fadd_rn(acc, __fmul_rn(diff, diff));
        diff = x.w - y.w;
        acc = __fadd_rn(acc, __fmul_rn(diff, diff));
      }
    } else {
      for (int d = 0; d < D; ++d) {
        const float diff = __ldg(a + d) - __ldg(b + d);
        acc = __fadd_rn(acc, __fmul_rn(diff, diff));
      }
    }
  }
  unsigned long long key = ~0ull;
  int zeros = 0;
  if (pair_ok) {
    key = ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned)j;
    zeros = (acc == 0.0f) ? 1 : 0;
    if (err_ratio_bits) {
      // observed |approximate - exact| distance of this candidate, relative to the bound the certificate
      // assumes (statistic for the tests: must stay well below 1)
      const float nai = na[ci], nbm = normmaxB[0];
      const float eps = 2.0f * eta * sqrtf(nai * nbm) + c_sum * (nai + nbm + acc) +
                        c_sub * (scaleA[1] * sqrtf(nbm) + scaleB[1] * sqrtf(nai)) +
                        c_pay * (nbm + 2.0f * sqrtf(nai * nbm)) + 1e-30f;
      const float ratio = fabsf((nai + s) - acc) / eps;
      if (ratio == ratio) atomicMax(err_ratio_bits, __float_as_uint(ratio));
    }
  }
  for (int o = nc >> 1; o > 0; o >>= 1) {
    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
    key = (ok < key) ? ok : key;
    zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
  }
  if (c == 0 && row_ok) {
    bool certified = false;
    if (key != ~0ull) {
      const float best_d2 = __uint_as_float((unsigned)(key >> 32));
      if (!(s_cut < __int_as_float(0x7f800000))) {
        certified = true;  // every valid model row was a candidate
      } else {
        const float nai = na[ci];
        const float nbm = normmaxB[0];
        // |s_approx - s_real| <= 2*eta*|a||b| (tc_eta); float32 sequential sums and norms: (D + 8) 2^-23 relative;
        // fp16 flush of tiny elements: c_sub
        const float eps = 2.0f * eta * sqrtf(nai * nbm) + c_sum * (nai + nbm + best_d2) +
                          c_sub * (scaleA[1] * sqrtf(nbm) + scaleB[1] * sqrtf(nai)) +
                          c_pay * (nbm + 2.0f * sqrtf(nai * nbm)) + 1e-30f;
        certified = best_d2 < (nai + s_cut) - eps;
      }
    }
    if (certified) {
      best[i] = key;
      zero_cnt[i] = zeros;
    } else {
      fb_rows[atomicAdd(fb_count, 1)] = i;
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_map(b200_ctx *ctx, CUtensorMap *map, const __half *base, int rows, int Kp, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return ctx->fail(B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ctx->fail(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return B200_OK;
}

}  // namespace

// Error bound of the approximate inner product, relative to |a||b| (derivation):
//   operands: fp16 round-to-nearest keeps 11 significant bits, |x^ - x| <= 2^-11 |x| for elements that stay normal
//   (tiny ones are covered by c_sub in the rescoring kernel).  One term: |a^.b^ - a.b| <= (2^-10 + 2^-22) sum|a_k b_k|.
//   Three terms: x = x1 + x2 + r with |x2| <= 2^-11 |x|, |r| <= 2^-22 |x|; the contraction a1.b1 + a1.b2 + a2.b1
//   drops a2.b2 and the r terms: <= (3 * 2^-22 + 2^-32) sum|a_k b_k|.  Cauchy-Schwarz: sum|a_k b_k| <= |a||b|.
//   accumulation: the products of two fp16 numbers are exact in fp32; the tensor core adds the K' products of a
//   row in fp32 with an unspecified order and possibly truncation.  Model: every one of the K' additions loses at
//   most 2^-23 of the running magnitude, itself <= sum|a^_k b^_k| <= 1.001 |a||b|; a factor 2 of safety on top:
//   K' * 2^-22.  (Observed errors are about sqrt(K') * 2^-24: tests/test_gpu_match_tc.py checks the observed /
//   assumed ratio and cross-checks every row against the exact kernel.)
static float tc_eta(int terms, int Kp) {
  const float operand = (terms == 3) ? (3.0f * 2.3841858e-7f + 2.4e-10f) : (9.765625e-4f + 2.3841858e-7f);
  return 1.001f * (operand + (float)Kp * 2.3841858e-7f);
}

static int tc_kp(int terms, int D) { return ((terms * D + TC_BK - 1) / TC_BK) * TC_BK; }

// Model-side operands, both splits.  mvalid: per-row validity flags (device).
int match_tc_prepare_model(b200_ctx *ctx, const float *d_model, int Km, int D, const unsigned char *mvalid,
                           TcModelPrep *out) {
  const int n_tiles = ceil_div(Km, TC_BN), rowsB = n_tiles * TC_BN;
  out->ready = false;
  out->Km = Km;
  out->D = D;
  out->rowsB = rowsB;
  B200_TRY(out->B1.alloc(ctx, (size_t)rowsB * tc_kp(1, D)));
  B200_TRY(out->B3.alloc(ctx, (size_t)rowsB * tc_kp(3, D)));
  B200_TRY(out->nb.alloc(ctx, (size_t)rowsB));
  B200_TRY(out->scaleB.alloc(ctx, 2));
  B200_TRY(out->bits.alloc(ctx, 2));
  B200_TRY(out->bits.zero());
  const int rb = std::min(ctx->sm_count * 8, 4096);
  absmax_kernel<<<rb, 256, 0, ctx->stream>>>(d_model, (size_t)Km * D, out->bits.p + 0);
  B200_LAUNCHED(ctx);
  make_scale_kernel<<<1, 1, 0, ctx->stream>>>(out->bits.p + 0, out->scaleB.p);
  B200_LAUNCHED(ctx);
  for (int terms = 1; terms <= 3; terms += 2) {
    tc_prep_kernel<<<ceil_div((long long)rowsB * 32, 256), 256, 0, ctx->stream>>>(
        d_model, Km, rowsB, D, tc_kp(terms, D), terms, 1, out->scaleB.p, mvalid,
        reinterpret_cast<__half *>(terms == 1 ? out->B1.p : out->B3.p), out->nb.p, out->bits.p + 1, nullptr, nullptr,
        n_tiles);
    B200_LAUNCHED(ctx);
  }
  out->ready = true;
  return B200_OK;
}

namespace {

// One filter + rescoring pass.  row_map / rows_dev: nullable (first pass: all Ks rows in order).  Rows it cannot
// certify are appended to fb_rows / fb_count.
int tc_pass(b200_ctx *ctx, const float *d_model, int Km, const TcModelPrep &B, const float *d_scene, int Ks,
            const unsigned char *svalid, int D, int terms, const float *scA, const int *row_map, const int *rows_dev,
            unsigned long long *best, int *zero_cnt, int *fb_rows, int *fb_count, unsigned *err_bits) {
  const int Kp = tc_kp(terms, D);
  const int m_tiles = ceil_div(Ks, TC_BM), n_tiles = B.rowsB / TC_BN;
  const int rowsA = m_tiles * TC_BM;
  int n_split = 1;
  if (rows_dev) {
    n_split = std::min(TC_MAX_SPLIT, n_tiles);  // few rows expected: spread them over the model tiles
    while (n_split & (n_split - 1)) --n_split;
  } else {
    while (n_split < TC_MAX_SPLIT && m_tiles * n_split < ctx->sm_count && n_split * 2 <= n_tiles) n_split *= 2;
  }
  DevBuf<__half> A16;
  DevBuf<float> na, cand_s, cand_b;
  DevBuf<int> cand_j;
  B200_TRY(A16.alloc(ctx, (size_t)rowsA * Kp));
  B200_TRY(na.alloc(ctx, (size_t)rowsA));
  B200_TRY(cand_s.alloc(ctx, (size_t)Ks * n_split * TC_CAND));
  B200_TRY(cand_j.alloc(ctx, (size_t)Ks * n_split * TC_CAND));
  B200_TRY(cand_b.alloc(ctx, (size_t)Ks * n_split));
  tc_prep_kernel<<<ceil_div((long long)(rows_dev ? Ks : rowsA) * 32, 256), 256, 0, ctx->stream>>>(
      d_scene, Ks, rows_dev ? Ks : rowsA, D, Kp, terms, 0, scA, nullptr, A16.p, na.p, nullptr, row_map, rows_dev, 0);
  B200_LAUNCHED(ctx);
  CUtensorMap mapA, mapB;
  B200_TRY(make_map(ctx, &mapA, A16.p, rowsA, Kp, TC_BM));
  B200_TRY(make_map(ctx, &mapB, reinterpret_cast<const __half *>(terms == 1 ? B.B1.p : B.B3.p), B.rowsB, Kp, TC_BN));
  TcParams p;
  p.Ks = Ks;
  p.Km = Km;
  p.m_tiles = m_tiles;
  p.n_tiles = n_tiles;
  p.n_split = n_split;
  p.k_blocks = Kp / TC_BK;
  p.nb = B.nb.p;
  p.scaleA = scA;
  p.scaleB = B.scaleB.p;
  p.cand_s = cand_s.p;
  p.cand_j = cand_j.p;
  p.cand_b = cand_b.p;
  p.rows_dev = rows_dev;
  B200_CUDA(ctx, ensure_dyn_smem(tc_filter_kernel, TC_SMEM));
  const int grid = std::min(ctx->sm_count, m_tiles * n_split);
  {
    StageScope st_(ctx, ST_MATCH_FILTER);
    tc_filter_kernel<<<grid, TC_THREADS, TC_SMEM, ctx->stream>>>(mapA, mapB, p);
    B200_LAUNCHED(ctx);
  }
  const float eta = tc_eta(terms, Kp);
  const int rows_per_cta = RS_WARPS_PER_CTA * (32 / (n_split * TC_CAND));
  tc_rescore_kernel<<<ceil_div(Ks, rows_per_cta), RS_WARPS_PER_CTA * 32, 0, ctx->stream>>>(
      d_model, Km, d_scene, Ks, D, svalid, n_split, cand_s.p, cand_j.p, cand_b.p, na.p,
      reinterpret_cast<const float *>(B.bits.p + 1), scA, B.scaleB.p, eta, best, zero_cnt, fb_rows, fb_count,
      err_bits, row_map, rows_dev, n_tiles);
  B200_LAUNCHED(ctx);
  return B200_OK;
}

}  // namespace

// plan: 1 = one-term pass only, 3 = three-term pass only, 13 = one-term pass over every row, three-term pass over
// the rows it could not certify.  On return best[] / zero_cnt[] hold the certified rows; fb_rows[0..*fb_count)
// lists the rows that need the exact kernel.  prep: the model side if it is resident (nullable).
int match_tc_filter(b200_ctx *ctx, const float *d_model, int Km, const unsigned char *mvalid, const TcModelPrep *prep,
                    const float *d_scene, int Ks, const unsigned char *svalid, int D, int plan,
                    unsigned long long *best, int *zero_cnt, int *fb_rows, int *fb_count) {
  if (tc_kp(3, D) > 4096) return ctx->fail(B200_ERR_INVALID, "match: descriptor too long for the tensor-core filter");
  TcModelPrep local;
  if (!prep || !prep->ready || prep->Km != Km || prep->D != D) {
    B200_TRY(match_tc_prepare_model(ctx, d_model, Km, D, mvalid, &local));
    prep = &local;
  }
  DevBuf<float> scA;
  DevBuf<unsigned> bitsA;
  DevBuf<int> fb1_rows, fb1_count;
  B200_TRY(scA.alloc(ctx, 2));
  B200_TRY(bitsA.alloc(ctx, 2));
  B200_TRY(bitsA.zero());
  B200_CUDA(ctx, cudaMemsetAsync(fb_count, 0, sizeof(int), ctx->stream));
  const int rb = std::min(ctx->sm_count * 8, 4096);
  absmax_kernel<<<rb, 256, 0, ctx->stream>>>(d_scene, (size_t)Ks * D, bitsA.p + 0);
  B200_LAUNCHED(ctx);
  make_scale_kernel<<<1, 1, 0, ctx->stream>>>(bitsA.p + 0, scA.p);
  B200_LAUNCHED(ctx);
  DevBuf<unsigned> errb;
  B200_TRY(errb.alloc(ctx, 1));
  B200_TRY(errb.zero());
  unsigned *err_bits = ctx->profiling ? errb.p : nullptr;
  if (plan == 13) {
    B200_TRY(fb1_rows.alloc(ctx, (size_t)Ks));
    B200_TRY(fb1_count.alloc(ctx, 1));
    B200_TRY(fb1_count.zero());
    B200_TRY(tc_pass(ctx, d_model, Km, *prep, d_scene, Ks, svalid, D, 1, scA.p, nullptr, nullptr, best, zero_cnt,
                     fb1_rows.p, fb1_count.p, err_bits));
    if (ctx->profiling)
      B200_CUDA(ctx, cudaMemcpyAsync(&ctx->last_match_pass1_fail, fb1_count.p, sizeof(int), cudaMemcpyDeviceToHost,
                                     ctx->stream));
    B200_TRY(tc_pass(ctx, d_model, Km, *prep, d_scene, Ks, svalid, D, 3, scA.p, fb1_rows.p, fb1_count.p, best,
                     zero_cnt, fb_rows, fb_count, nullptr));
  } else {
    B200_TRY(tc_pass(ctx, d_model, Km, *prep, d_scene, Ks, svalid, D, plan, scA.p, nullptr, nullptr, best, zero_cnt,
                     fb_rows, fb_count, err_bits));
  }
  if (ctx->profiling)
    B200_CUDA(ctx, cudaMemcpyAsync(&ctx->last_match_err_ratio, errb.p, sizeof(float),
                                   cudaMemcpyDeviceToHost, ctx->stream));
  return B200_OK;
}
