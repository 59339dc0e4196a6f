// dsgd_persistent.cuh -- the synchronous SGD loop as ONE persistent cooperative kernel (sm_100a).
//
// Replaces, for a whole run of consecutive steps, the body of Master.fit's batch loop
// (core/Master.scala:179-198) together with the slave's gradient request (core/Slave.scala:142-157):
// no launch, no host round trip and exactly ONE grid-wide barrier per SGD step.
//
// Why it looks the way it does (measured in profiles/r1a: a two-kernel step costs 25 us while its 256
// row windows are 197 KB -- the step is a chain of dependent latencies, not bandwidth):
//   * Row windows do not depend on the weights, so they are fetched AHEAD of the step that needs them:
//     each row warp owns a ring of shared-memory slots filled by TMA bulk copies (cp.async.bulk +
//     mbarrier complete_tx), issued kSlots rows in advance; the sample id -> row pointer -> window
//     address chain is software-pipelined two further rows ahead in registers.
//   * Weights are double-buffered and gradients triple-buffered in L2 so that the update of step t and
//     the gradient of step t+1 run in the SAME barrier interval: a row warp reads W_{t-1}[col] and
//     g_{t-1}[col] and applies the update arithmetic itself ("on the fly") while the update warps write
//     the same values to the W_t buffer for the interval after.  One barrier per step instead of two.
//   * c = 2*lambda*(w . d) and ||w||^2 of every new weight vector are produced by the update warps as
//     per-CTA partials and summed by every warp in a fixed order: deterministic, no extra barrier.
//
// Interval I_t (between barrier t-1 and barrier t), with W_t the weights step t differentiates at:
//   row warps   : x.W_t with W_t[col] computed on the fly from (W_{t-1}, g_{t-1}, c_{t-1}); gate; RED y*x into g_t
//   update warps: W_t buffer <- update(W_{t-1}, g_{t-1}, c_{t-1}); zero g_{t+1}'s buffer; partials of c_t, ||W_t||^2
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dsgd_kernels.cuh"

namespace dsgd {

struct PersistParams {
  const uint32_t *rp16;
  const uint2 *pairs;
  const int8_t *label;
  const int32_t *samples;  // n_steps * batch ids, step-major
  int64_t n_steps;
  int32_t batch;
  int32_t dim;
  double *wbuf[2];  // on entry wbuf[1] holds the initial weights ("W_{-1}" == W_0)
  double *gbuf[3];  // all zero on entry and on exit
  const double *d;
  double *partial;         // [2][gridDim.x][2]
  unsigned *hinge;         // [n_steps], zero on entry
  double *losses;          // [n_steps] or nullptr
  double *w_out;           // resident weights after the last step
  float *w32_out;
  double *scal;            // kScalC / kScalNrm2 of the resident weights
  unsigned *bar;           // grid barrier counter, zero on entry
  int *abort_flag;         // set to 1 if a wait hit the watchdog
  double lambda, lr, k_den;
  long long timeout_cycles;
  long long *tl;           // debug timeline: [256 steps][16 stamps] of clock64 (CTA 0), or nullptr
};

// ---- PTX helpers: mbarrier + TMA bulk copy -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *b, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// w_j after one SGD update given the raw gradient-sum entry (same arithmetic as k_update<true>):
// regularize on the surviving key (SparseSVM.scala:31), mean over workers, times lr, subtract, each with the
// Sparse constructor's 1e-20 filter (core/Master.scala:194,197; math/Sparse.scala:108-118).
__device__ __forceinline__ double apply_update(double wv, double graw, double c, bool add_c, double k_den, double lr) {
  double v = filt(graw);
  if (v != 0.0) {
    if (add_c) v = filt(v + c);
    if (v != 0.0) {
      const double mean = filt(v / k_den);
      const double step = filt(mean * lr);
      wv = filt(wv - step);
    }
  }
  return wv;
}

// Fixed-order sum of the per-CTA partials (stride 2 doubles per CTA); identical in every warp.
__device__ __forceinline__ double sum_partials(const double *p, int n_cta, int lane) {
  double s = 0.0;
  for (int b = lane; b < n_cta; b += 32) s += __ldcg(&p[2 * b]);
  return warp_sum(s);
}

// One grid-wide barrier: every CTA arrives once; `target` = number of arrivals that completes this phase.
// Returns false if the watchdog fired (or another CTA raised the abort flag).
__device__ __forceinline__ bool grid_barrier(unsigned *bar, unsigned target, int *abort_flag, long long timeout,
                                             int *smem_ok, long long *tl = nullptr) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (tl) tl[0] = clock64();
    red_release_gpu_add(bar, 1u);
    int ok = 1;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (ld_acquire_gpu(bar) < target) {
      if ((++spins & 1023u) == 0u) {
        if (clock64() - t0 > timeout || *(volatile int *)abort_flag) {
          *(volatile int *)abort_flag = 1;
          ok = 0;
          break;
        }
      }
    }
    *smem_ok = ok;
    if (tl) tl[1] = clock64();
  }
  __syncthreads();
  return *smem_ok != 0;
}

#define DSGD_TL(slot_)                                                                        \
  do {                                                                                        \
    if (p.tl && blockIdx.x == 0 && lane == 0 && t < 256) p.tl[t * 16 + (slot_)] = clock64(); \
  } while (0)

template <int kRowWarps, int kUpdWarps, int kSlots, int kCapPairs>
struct PersistSmem {
  uint2 ring[kRowWarps][kSlots][kCapPairs];
  uint64_t mbar[kRowWarps][kSlots];
  double red[kUpdWarps][2];
  int ok;
};

template <int kRowWarps, int kUpdWarps, int kSlots, int kCapPairs>
__global__ void __launch_bounds__((kRowWarps + kUpdWarps) * 32, 1) k_sync_persistent(const PersistParams p) {
  using Smem = PersistSmem<kRowWarps, kUpdWarps, kSlots, kCapPairs>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const bool is_row = warp < kRowWarps;
  const int G = gridDim.x;
  const int B = p.batch;
  const int64_t S = p.n_steps;

  if (is_row && lane == 0) {
#pragma unroll
    for (int s = 0; s < kSlots; ++s) mbar_init(&sm.mbar[warp][s], 1u);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  // ---- row-warp prefetch pipeline state (all lanes hold the same values) -------------------------------
  const int NW = G * kRowWarps;
  const int wg = warp * G + blockIdx.x;  // row i of a step goes to CTA i % G: spreads a small batch over all CTAs
  const int RPW = (B + NW - 1) / NW;     // rows per warp per step
  const int64_t Q = S * RPW;             // length of this warp's row sequence (some entries may be holes)
  // per-slot metadata of the rows currently in the ring
  uint32_t m_start[kSlots];  // window start, 16-byte units
  int m_len[kSlots];         // pairs, padding included; -1: hole (no row)
  int m_y[kSlots];
  unsigned m_par[kSlots];    // mbarrier phase parity the copy into this slot completes
  unsigned par_mask = 0u;    // bit s: parity of the NEXT copy into slot s
  // two-deep register pipeline ahead of the copy: ids, then row pointers
  int64_t q_next = 0;           // next sequence number to enter stage A
  int32_t stA_row = -1;         // stage A result: row id of q_next-1 (or -1 hole)
  int32_t stB_len = -1;         // stage B result: window of q_next-2
  uint32_t stB_start = 0;
  int stB_y = 0;

  auto seq_row = [&](int64_t q) -> int32_t {  // row id of sequence entry q, or -1
    if (q >= Q) return -1;
    const int64_t t = q / RPW;
    const int i = wg + (int)(q % RPW) * NW;
    return i < B ? __ldg(&p.samples[t * B + i]) : -1;
  };
  auto stage_a = [&]() { stA_row = seq_row(q_next); ++q_next; };
  auto stage_b = [&]() {  // consumes stA_row
    if (stA_row >= 0) {
      const uint32_t b = __ldg(&p.rp16[stA_row]), e = __ldg(&p.rp16[stA_row + 1]);
      stB_start = b;
      stB_len = (int)(e - b) * 2;
      stB_y = (int)__ldg(&p.label[stA_row]);
    } else {
      stB_len = -1;
    }
  };
  auto stage_c = [&](int slot) {  // consumes stB_*: TMA the window (its first kCapPairs pairs) into `slot`
    m_start[slot] = stB_start;
    m_len[slot] = stB_len;
    m_y[slot] = stB_y;
    m_par[slot] = (par_mask >> slot) & 1u;
    if (stB_len > 0) par_mask ^= (1u << slot);
    if (stB_len > 0 && lane == 0) {
      const unsigned bytes = (unsigned)(stB_len < kCapPairs ? stB_len : kCapPairs) * 8u;
      mbar_expect_tx(&sm.mbar[warp][slot], bytes);
      bulk_g2s(&sm.ring[warp][slot][0], p.pairs + (size_t)stB_start * 2, bytes, &sm.mbar[warp][slot]);
    }
  };

  if (is_row) {
    // prologue: fill the ring (blocking loads, once)
    stage_a();
    stage_b();
    stage_a();
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      stage_c(s);
      stage_b();
      stage_a();
    }
  }

  const int n_upd = G * kUpdWarps * 32;
  const int u0 = blockIdx.x * kUpdWarps * 32 + (threadIdx.x - kRowWarps * 32);
  const double k_den = p.k_den, lr = p.lr;
  unsigned phase = 0;
  int64_t q = 0;  // row warps: next sequence entry to consume

  for (int64_t t = 0; t <= S; ++t) {
    const double *Wprev = p.wbuf[(t + 1) & 1];
    double *Wcur = p.wbuf[t & 1];
    const double *Gprev = p.gbuf[(t + 2) % 3];
    double *Gcur = p.gbuf[t % 3];
    double *Gzero = p.gbuf[(t + 1) % 3];
    const double *part_prev = p.partial + (size_t)((t + 1) & 1) * G * 2;
    double *part_cur = p.partial + (size_t)(t & 1) * G * 2;

    // c_{t-1} = 2*lambda*(W_{t-1} . d): at t == 0 g_{-1} is all zero, so its value is irrelevant
    if (warp == 0) DSGD_TL(0);
    if (warp == kRowWarps) DSGD_TL(8);
    double c_prev = 0.0;
    if (t > 0) c_prev = p.lambda * 2.0 * sum_partials(part_prev, G, lane);
    if (warp == 0) DSGD_TL(1);
    const bool add_c = (c_prev != 0.0) && (fabs(c_prev) > kEps);

    if (is_row) {
      if (t < S) {
        unsigned hinge = 0;
        for (int m = 0; m < RPW; ++m, ++q) {
          const int slot = (int)(q % kSlots);
          // static indexing of the register arrays
          uint32_t start = 0;
          unsigned parity = 0u;
          int len = -1, yi = 0;
#pragma unroll
          for (int s = 0; s < kSlots; ++s)
            if (s == slot) { start = m_start[s]; len = m_len[s]; yi = m_y[s]; parity = m_par[s]; }
          if (len >= 0) {
            const int n_smem = len < kCapPairs ? len : kCapPairs;
            const uint2 *srow = &sm.ring[warp][slot][0];
            const uint2 *grow = p.pairs + (size_t)start * 2;
            if (len > 0) {
              while (!mbar_try_wait(&sm.mbar[warp][slot], parity)) {}
            }
            if (warp == 0) DSGD_TL(2);
            double acc = 0.0;
            for (int k0 = 0; k0 < len; k0 += 128) {
              uint2 pr[4];
              double wv[4], gv[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int k = k0 + u * 32 + lane;
                pr[u] = make_uint2(0u, 0u);
                if (k < len) pr[u] = (k < n_smem) ? srow[k] : __ldg(&grow[k]);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int k = k0 + u * 32 + lane;
                wv[u] = 0.0;
                gv[u] = 0.0;
                if (k < len) {
                  wv[u] = __ldcg(&Wprev[pr[u].x]);
                  gv[u] = __ldcg(&Gprev[pr[u].x]);
                }
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const double xv = filt((double)__uint_as_float(pr[u].y));
                const double wt = apply_update(wv[u], gv[u], c_prev, add_c, k_den, lr);
                acc += filt(xv * wt);  // (x * w).sum  (math/Vec.scala:58)
              }
            }
            const double dot = warp_sum(acc);
            if (warp == 0) DSGD_TL(3);
            const double y = (double)yi;
            const int pred = (dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0);
            hinge += (unsigned)(1 - yi * pred);
            if (!(y * dot < 0.0)) {  // SparseSVM.scala:28
              for (int k = lane; k < len; k += 32) {
                const uint2 pr = (k < n_smem) ? srow[k] : __ldg(&grow[k]);
                const double gvv = filt(filt((double)__uint_as_float(pr.y)) * y);
                if (gvv != 0.0) atomicAdd(&Gcur[pr.x], gvv);
              }
            }
          }
          if (warp == 0) DSGD_TL(4);
          __syncwarp();
          // the slot is free: refill it kSlots rows ahead, advance the register pipeline
          {
#pragma unroll
            for (int s = 0; s < kSlots; ++s)
              if (s == slot) stage_c(s);
            stage_b();
            stage_a();
          }
        }
        if (lane == 0 && hinge) atomicAdd(&p.hinge[t], hinge);
        if (warp == 0) DSGD_TL(5);
      }
    } else {
      // ---- update warps: W_t buffer, zero the buffer g_{t+1} will use, partials of c_t and ||W_t||^2 ----
      double pd = 0.0, pn = 0.0;
      for (int j = u0; j < p.dim; j += n_upd) {
        const double wn = apply_update(__ldcg(&Wprev[j]), __ldcg(&Gprev[j]), c_prev, add_c, k_den, lr);
        Wcur[j] = wn;
        Gzero[j] = 0.0;
        pd += filt(wn * __ldg(&p.d[j]));
        pn += wn * wn;
      }
      pd = warp_sum(pd);
      pn = warp_sum(pn);
      if (warp == kRowWarps) DSGD_TL(9);
      const int uw = warp - kRowWarps;
      if (lane == 0) { sm.red[uw][0] = pd; sm.red[uw][1] = pn; }
      asm volatile("bar.sync 1, %0;" ::"r"(kUpdWarps * 32) : "memory");
      if (uw == 0 && lane == 0) {
        double sd = 0.0, sn = 0.0;
#pragma unroll
        for (int i = 0; i < kUpdWarps; ++i) { sd += sm.red[i][0]; sn += sm.red[i][1]; }
        part_cur[2 * blockIdx.x] = sd;
        part_cur[2 * blockIdx.x + 1] = sn;
      }
      // loss of step t-1 = lambda*||W_{t-1}||^2 + hinge_{t-1}/batch  (SparseSVM.scala:20-23)
      if (t > 0 && p.losses && blockIdx.x == 0 && uw == 1) {
        const double nrm = sum_partials(part_prev + 1, G, lane);
        if (lane == 0) p.losses[t - 1] = p.lambda * nrm + (double)__ldcg(&p.hinge[t - 1]) / (double)B;
      }
    }
    if (warp == kRowWarps) DSGD_TL(10);
    ++phase;
    if (!grid_barrier(p.bar, phase * (unsigned)G, p.abort_flag, p.timeout_cycles, &sm.ok, (p.tl && blockIdx.x == 0 && t < 256) ? p.tl + t * 16 + 6 : nullptr)) return;
  }

  // ---- epilogue: W_S is complete in wbuf[S & 1]; publish it as the resident weights, clear g_{S-1} ----------
  if (!is_row) {
    const double *Wfin = p.wbuf[S & 1];
    double *Glast = p.gbuf[(S + 2) % 3];
    for (int j = u0; j < p.dim; j += n_upd) {
      const double wv = __ldcg(&Wfin[j]);
      p.w_out[j] = wv;
      p.w32_out[j] = (float)wv;
      Glast[j] = 0.0;
    }
    if (blockIdx.x == 0 && warp == kRowWarps) {
      const double *part = p.partial + (size_t)(S & 1) * G * 2;
      const double sd = sum_partials(part, G, lane);
      const double sn = sum_partials(part + 1, G, lane);
      if (lane == 0) {
        p.scal[kScalC] = p.lambda * 2.0 * sd;
        p.scal[kScalNrm2] = sn;
      }
    }
  }
}

}  // namespace dsgd
