// dsgd_stream_x.cuh -- EXPERIMENTAL variants of the streaming pass of dsgd_stream.cuh (selected with DSGD_STREAM_OPT;
// not the default path: written after the round's GPU budget ran out, to be measured first thing in round 2).  The body
// is the shipped kernel's; the differences are the two blocks described below.  When a variant wins it replaces
// k_stream_rows and this file goes away.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dsgd_kernels.cuh"
#include "dsgd_stream.cuh"

namespace dsgd {

struct StreamParamsX {
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
  // ---- experimental variants (template parameter kOpt; not the default path) ----
  const uint32_t *hot_bits;    // kOpt & 2: bit c set = column c has a shared-memory accumulator slot
  const uint16_t *hot_prefix;  //           slots before word c >> 5 (slot = prefix + popc of the lower bits of the word)
  const int32_t *hot_cols;     //           slot -> column
  int n_hot;                   //           slots in use (<= kHotSlots)
};

// ---- variant kOpt & 1: fp32 fast path.  The dot is needed for its SIGN only (prediction, gate), so it is
// accumulated with fp32 FMAs -- no fp32->fp64 conversions (ncu: the XU pipe they run on was 46 % busy) -- and the
// exact fp64 recomputation takes every row whose |dot| is inside the rounding band
//   (m + 1) * 2^-24 * max|w| * sum|x|,  m = 2 * ceil(pairs / 32) + 4 roundings on the longest add chain
// (m for the accumulation, + 1 for rounding w to fp32; Higham's gamma_m bound, 1.5x slack as before).
// ---- variant kOpt & 2 (scatter): hot-column accumulators in shared memory.  The large-batch scatter is bound by the
// fp64 RED rate at L2 (0.48 per SM-cycle) and column frequencies are Zipfian: ~70 % of the non-zeros fall on a few
// thousand columns.  Each CTA keeps an exact fixed-point accumulator (g * 2^40 as three 32-bit limbs: two of 14 bits, a
// signed rest; shared memory has native 32-bit atomics only, 64-bit and fp64 ones compile to CAS loops) for the
// kHotSlots most frequent columns and flushes it with one RED per touched slot at the end.  Integer sums are exact and
// order-free; values that are not multiples of 2^-40 or exceed 1 in magnitude take the RED path.  A launch covers at
// most kHotMaxRows rows so a limb cannot overflow (2^18 adds of < 2^14, resp. <= 2^12 in magnitude).
constexpr int kHotSlots = 2688;
constexpr int64_t kHotMaxRows = 1 << 18;
__host__ __device__ constexpr size_t stream_smem_bytes(int dim, int opt) {
  const size_t ws = (((size_t)dim + 3) & ~(size_t)3) * sizeof(float);
  const size_t words = ((size_t)dim + 31) / 32;
  return (opt & 2) ? ws + (size_t)kHotSlots * 12 + words * 4 + ((words * 2 + 15) & ~(size_t)15) : (size_t)dim * sizeof(float);
}

template <bool kScatter, bool kPreds, int kOpt>
__global__ void __launch_bounds__(kStreamThreads, 1) k_stream_rows_x(const StreamParamsX p) {
  constexpr bool kFast32 = (kOpt & 1) != 0;
  constexpr bool kHot = kScatter && (kOpt & 2) != 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *ws = reinterpret_cast<float *>(smem_raw);
  __shared__ float s_wmax[kStreamThreads / 32];
  __shared__ unsigned long long s_cnt[2];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- stage the fp32 weights, find max|w| ----
  float wmax = 0.f;
  {
    const float4 *src = reinterpret_cast<const float4 *>(p.w32);
    float4 *dst = reinterpret_cast<float4 *>(ws);
    const int n4 = p.dim >> 2;
    for (int i = threadIdx.x; i < n4; i += kStreamThreads) {
      const float4 v = __ldg(&src[i]);
      dst[i] = v;
      wmax = fmaxf(wmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (int i = (n4 << 2) + threadIdx.x; i < p.dim; i += kStreamThreads) {
      const float v = __ldg(&p.w32[i]);
      ws[i] = v;
      wmax = fmaxf(wmax, fabsf(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_wmax[warp] = wmax;
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0ull;
    if constexpr (kHot) {
      const int words = (p.dim + 31) >> 5;
      uint32_t *acc = reinterpret_cast<uint32_t *>(ws + ((p.dim + 3) & ~3));
      uint32_t *hb = acc + 3 * kHotSlots;
      uint16_t *hp = reinterpret_cast<uint16_t *>(hb + words);
      for (int i = threadIdx.x; i < 3 * kHotSlots; i += kStreamThreads) acc[i] = 0u;
      for (int i = threadIdx.x; i < words; i += kStreamThreads) { hb[i] = __ldg(&p.hot_bits[i]); hp[i] = __ldg(&p.hot_prefix[i]); }
    }
    __syncthreads();
    wmax = 0.f;
#pragma unroll
    for (int i = 0; i < kStreamThreads / 32; ++i) wmax = fmaxf(wmax, s_wmax[i]);
  }
  // |x.w - x.w32| <= 2^-24 * max|w| * sum|x| (+ fp32 underflow slack); 1.5x covers fp32 rounding of the bound itself
  const float band_scale = 1.5f * 5.9604645e-8f * wmax;
  uint32_t *hot_acc = nullptr;
  const uint32_t *hot_bits = nullptr;
  const uint16_t *hot_prefix = nullptr;
  if constexpr (kHot) {
    hot_acc = reinterpret_cast<uint32_t *>(ws + ((p.dim + 3) & ~3));
    hot_bits = hot_acc + 3 * kHotSlots;
    hot_prefix = reinterpret_cast<const uint16_t *>(hot_bits + ((p.dim + 31) >> 5));
  }
  // one gradient entry: into the CTA's fixed-point slot if the column has one and the value is exactly representable
  auto scatter_one = [&](uint32_t col, double gv) {
    if (gv == 0.0) return;
    if constexpr (kHot) {
      const uint32_t bits = hot_bits[col >> 5], bit = 1u << (col & 31);
      if (bits & bit) {
        const double sv = gv * 1099511627776.0;  // 2^40: exact scaling
        if (fabs(sv) <= 1099511627776.0) {
          const long long iv = __double2ll_rn(sv);
          if ((double)iv == sv) {
            const uint32_t slot = (uint32_t)hot_prefix[col >> 5] + (uint32_t)__popc(bits & (bit - 1u));
            const int l2 = (int)(iv >> 28);                                  // signed rest, |l2| <= 2^12
            const uint32_t rem = (uint32_t)(iv - ((long long)l2 << 28));     // in [0, 2^28)
            atomicAdd(&hot_acc[3 * slot], rem & 0x3fffu);
            atomicAdd(&hot_acc[3 * slot + 1], rem >> 14);
            atomicAdd(&hot_acc[3 * slot + 2], (uint32_t)l2);
            return;
          }
        }
      }
    }
    atomicAdd(&p.g[col], gv);
  };

  const int half = lane >> 4, hl = lane & 15;
  const int64_t n_blocks = (p.n + 31) >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * (kStreamThreads / 32) + warp;
  const int64_t n_warps = (int64_t)gridDim.x * (kStreamThreads / 32);
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
      float acc32 = 0.f;
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
          if constexpr (kFast32) {
            acc32 = __fmaf_rn(x0, w0, acc32);
            acc32 = __fmaf_rn(x1, w1, acc32);
          } else {
            acc = fma((double)x0, (double)w0, acc);
            acc = fma((double)x1, (double)w1, acc);
          }
          asum += fabsf(x0) + fabsf(x1);
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        if constexpr (kFast32) acc32 += __shfl_xor_sync(0xffffffffu, acc32, o);
        else acc += __shfl_xor_sync(0xffffffffu, acc, o);
        asum += __shfl_xor_sync(0xffffffffu, asum, o);
      }
      double dot = kFast32 ? (double)acc32 : acc;
      const bool valid = rid >= 0;
      // exact fallback: the fp32-rounded weights cannot decide the sign (includes dot == 0 with non-empty rows)
      bool ambiguous;
      if constexpr (kFast32) {
        const float m1 = (float)(2u * (((e - b) * 2u + 31u) >> 5) + 5u);   // m + 1 (pairs = 2 per 16-byte unit)
        ambiguous = valid && (e > b) && (fabs(dot) <= (double)(band_scale * asum) * (double)m1 + 1e-30);
      } else {
        ambiguous = valid && (e > b) && (fabs(dot) <= (double)(band_scale * asum) + 1e-300);
      }
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
              if constexpr (kHot) {
                scatter_one(q.x, g0);
                scatter_one(q.z, g1);
              } else {
                if (g0 != 0.0) atomicAdd(&p.g[q.x], g0);
                if (g1 != 0.0) atomicAdd(&p.g[q.z], g1);
              }
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
  if constexpr (kHot) {
    // every warp is past its last scatter (the barrier above): flush the touched slots, one RED each
    for (int slot = threadIdx.x; slot < p.n_hot; slot += kStreamThreads) {
      const long long tot = (long long)hot_acc[3 * slot] + ((long long)hot_acc[3 * slot + 1] << 14) +
                            ((long long)(int)hot_acc[3 * slot + 2] << 28);
      if (tot != 0) {
        double *dst = &p.g[__ldg(&p.hot_cols[slot])];
        if (tot > -(1ll << 53) && tot < (1ll << 53)) {
          atomicAdd(dst, (double)tot * 0x1p-40);             // the conversion is exact
        } else {
          const long long hi = tot >> 30, lo = tot - (hi << 30);             // both convert exactly
          atomicAdd(dst, (double)hi * 0x1p-10);
          atomicAdd(dst, (double)lo * 0x1p-40);
        }
      }
    }
  }
}

}  // namespace dsgd
