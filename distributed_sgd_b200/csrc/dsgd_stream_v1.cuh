// dsgd_stream.cuh -- streaming pass over many row windows: Master.localLoss/localAccuracy (core/Master.scala:
// 100-107), SlaveImpl.forward (core/Slave.scala:129-140) and large-batch SlaveImpl.gradient (142-157).
//
// This is the HBM-bound form of the hot path (roofline: 8*nnz + 16 bytes per sample).  What the
// kernel does to stay on the HBM roof instead of the L2 one:
//   * the weight vector is staged ONCE per CTA into shared memory as fp32 (47 236 x 4 B = 189 KB of the 227 KB),
//     so the ~94 gathers per row hit shared-memory banks, not L2 sectors (a 4-byte gather costs a 32-byte
//     sector at L2: 4x the row stream itself);
//   * one persistent CTA per SM, 32 warps; a warp owns blocks of 32 consecutive rows (~24 KB contiguous), loads
//     their bounds and labels with one coalesced access, and two 16-lane groups walk rows with 128-bit loads
//     (2 pairs per lane); the next block's bounds are fetched while the current block is processed;
//   * products are exact in fp64 ((double)x * (double)w32) and accumulated in fp64.
// Exactness against the fp64 weights the reference uses: rounding w to fp32 perturbs x.w by at most
// 2^-24 * max|w| * sum|x_j|.  Rows whose |x.w| is inside that band (about one in a million) are recomputed with
// the fp64 weights from L2, so predictions and gate decisions are those of the fp64 arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dsgd_kernels.cuh"

namespace dsgd {

constexpr int kStreamThreadsV1 = 1024;

struct StreamParamsV1 {
  const uint32_t *rp16;
  const uint4 *units;      // the pair array viewed as 16-byte units (2 pairs)
  const int8_t *label;
  const int32_t *samples;  // nullptr: rows [row_begin, row_begin + n)
  int64_t row_begin, n;
  const double *w;         // fp64 weights (exact fallback, L2)
  const float *w32;        // fp32 shadow of the same weights
  int dim;
  double *g;               // scatter target (fp64, L2) or nullptr
  double *preds;           // per-sample predictions or nullptr
  unsigned long long *cnt; // kCntHinge / kCntCorrect
  unsigned long long *n_exact;  // how many rows took the exact fallback (diagnostic), may be nullptr
  unsigned long long *next_block;  // work counter (zero on entry): blocks beyond the first wave are claimed dynamically
};

template <bool kScatter, bool kPreds>
__global__ void __launch_bounds__(kStreamThreadsV1, 1) k_stream_rows_v1(const StreamParamsV1 p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *ws = reinterpret_cast<float *>(smem_raw);
  __shared__ float s_wmax[kStreamThreadsV1 / 32];
  __shared__ unsigned long long s_cnt[2];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- stage the fp32 weights, find max|w| ----
  float wmax = 0.f;
  {
    const float4 *src = reinterpret_cast<const float4 *>(p.w32);
    float4 *dst = reinterpret_cast<float4 *>(ws);
    const int n4 = p.dim >> 2;
    for (int i = threadIdx.x; i < n4; i += kStreamThreadsV1) {
      const float4 v = __ldg(&src[i]);
      dst[i] = v;
      wmax = fmaxf(wmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (int i = (n4 << 2) + threadIdx.x; i < p.dim; i += kStreamThreadsV1) {
      const float v = __ldg(&p.w32[i]);
      ws[i] = v;
      wmax = fmaxf(wmax, fabsf(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_wmax[warp] = wmax;
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0ull;
    __syncthreads();
    wmax = 0.f;
#pragma unroll
    for (int i = 0; i < kStreamThreadsV1 / 32; ++i) wmax = fmaxf(wmax, s_wmax[i]);
  }
  // |x.w - x.w32| <= 2^-24 * max|w| * sum|x| (+ fp32 underflow slack); 1.5x covers fp32 rounding of the bound itself
  const float band_scale = 1.5f * 5.9604645e-8f * wmax;

  const int half = lane >> 4, hl = lane & 15;
  const int64_t n_blocks = (p.n + 31) >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * (kStreamThreadsV1 / 32) + warp;
  const int64_t n_warps = (int64_t)gridDim.x * (kStreamThreadsV1 / 32);
  unsigned hinge = 0, correct = 0, n_exact = 0;

  // bounds of the block being processed / the next one: lane l holds row l of the block
  auto load_block = [&](int64_t blk, uint32_t &b, uint32_t &e, int &y, int64_t &rid) {
    const int64_t i = (blk << 5) + lane;
    b = 0u; e = 0u; y = 0; rid = -1;
    if (blk < n_blocks && i < p.n) {
      rid = p.samples ? (int64_t)__ldg(&p.samples[i]) : p.row_begin + i;
      b = __ldg(&p.rp16[rid]);
      e = __ldg(&p.rp16[rid + 1]);
      y = (int)__ldg(&p.label[rid]);
    }
  };
  // Work distribution: the first wave is static (block = warp id), later blocks are claimed from a global counter
  // one step ahead (so the next block's bounds are prefetched while the current one is processed).  Rows are
  // heavy-tailed (1..2000 non-zeros): dynamic claiming keeps the tail of the pass short.
  auto claim = [&]() -> int64_t {
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(p.next_block, 1ull);
    return (int64_t)__shfl_sync(0xffffffffu, v, 0) + n_warps;
  };
  uint32_t nb, ne; int ny; int64_t nrid;
  int64_t blk = warp_global;
  int64_t blk_next = (blk < n_blocks) ? claim() : n_blocks;
  load_block(blk, nb, ne, ny, nrid);
  for (; blk < n_blocks;) {
    const uint32_t cb = nb, ce = ne; const int cy = ny; const int64_t crid = nrid;
    load_block(blk_next, nb, ne, ny, nrid);
    // 16 iterations: in iteration j the two halves take rows 2j and 2j+1 of the block.  A row is walked in
    // super-steps of kUnroll 128-bit loads per lane (kUnroll * 32 pairs per 16-lane group), all issued before the
    // first use: the bytes in flight per SM, not the instruction count, decide how close to the HBM roof this
    // runs, and long rows (a third of the non-zeros sit beyond a row's first 96 pairs) must not serialise.
    constexpr int kUnroll = 4;
#pragma unroll 1
    for (int j = 0; j < 16; ++j) {
      const int row_l = 2 * j + half;
      const uint32_t b = __shfl_sync(0xffffffffu, cb, row_l), e = __shfl_sync(0xffffffffu, ce, row_l);
      const int yi = __shfl_sync(0xffffffffu, cy, row_l);
      const int64_t rid = __shfl_sync(0xffffffffu, crid, row_l);
      double acc = 0.0;
      float asum = 0.f;
      for (uint32_t u0 = b + hl; u0 < e; u0 += 16u * kUnroll) {
        uint4 q[kUnroll];
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) {
          q[i] = make_uint4(0u, 0u, 0u, 0u);  // col 0, val +0.0f: inert
          if (u0 + 16u * i < e) q[i] = __ldg(&p.units[u0 + 16u * i]);
        }
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) {
          const float x0 = __uint_as_float(q[i].y), x1 = __uint_as_float(q[i].w);
          const float w0 = ws[q[i].x], w1 = ws[q[i].z];
          // fp32 x fp32 products are exact in fp64 (24 + 24 significant bits), so a fused multiply-add rounds
          // exactly like multiply-then-add: same bits as the unfused form, one instruction less
          acc = fma((double)x0, (double)w0, acc);
          acc = fma((double)x1, (double)w1, acc);
          asum += fabsf(x0) + fabsf(x1);
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        asum += __shfl_xor_sync(0xffffffffu, asum, o);
      }
      double dot = acc;
      const bool valid = rid >= 0;
      // exact fallback: the fp32-rounded weights cannot decide the sign (includes dot == 0 with non-empty rows)
      const bool ambiguous = valid && (e > b) && (fabs(dot) <= (double)(band_scale * asum) + 1e-300);
      if (ambiguous) {
        double ex = 0.0;
        for (uint32_t u = b + hl; u < e; u += 16) {
          const uint4 q = __ldg(&p.units[u]);
          ex += filt(filt((double)__uint_as_float(q.y)) * __ldcg(&p.w[q.x]));
          ex += filt(filt((double)__uint_as_float(q.w)) * __ldcg(&p.w[q.z]));
        }
        // only this 16-lane group is here (the other group's row may not be ambiguous): group-local mask
        const unsigned gmask = half ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) ex += __shfl_xor_sync(gmask, ex, o);
        dot = ex;
        if (hl == 0) ++n_exact;
      }
      if (valid) {
        const int pred = (dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0);
        if (hl == 0) {
          hinge += (unsigned)(1 - yi * pred);
          correct += (unsigned)(pred == yi);
          if (kPreds) p.preds[((blk << 5) + row_l)] = (double)pred;
        }
        if (kScatter) {
          const double y = (double)yi;
          if (!(y * dot < 0.0)) {  // SparseSVM.scala:28
            for (uint32_t u = b + hl; u < e; u += 16) {
              const uint4 q = __ldg(&p.units[u]);
              const double g0 = filt(filt((double)__uint_as_float(q.y)) * y);
              const double g1 = filt(filt((double)__uint_as_float(q.w)) * y);
              if (g0 != 0.0) atomicAdd(&p.g[q.x], g0);
              if (g1 != 0.0) atomicAdd(&p.g[q.z], g1);
            }
          }
        }
      }
    }
    blk = blk_next;
    blk_next = (blk < n_blocks) ? claim() : n_blocks;
  }
  // ---- counters: lane -> warp -> CTA -> one atomic per CTA ----
  hinge = __reduce_add_sync(0xffffffffu, hinge);
  correct = __reduce_add_sync(0xffffffffu, correct);
  n_exact = __reduce_add_sync(0xffffffffu, n_exact);
  if (lane == 0) {
    atomicAdd(&s_cnt[0], (unsigned long long)hinge);
    atomicAdd(&s_cnt[1], (unsigned long long)correct);
    if (p.n_exact && n_exact) atomicAdd(p.n_exact, (unsigned long long)n_exact);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_cnt[0]) atomicAdd(&p.cnt[kCntHinge], s_cnt[0]);
    if (s_cnt[1]) atomicAdd(&p.cnt[kCntCorrect], s_cnt[1]);
  }
}

}  // namespace dsgd
