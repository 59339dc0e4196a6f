// dsgd_persistent.cuh -- the synchronous SGD loop as ONE persistent, warp-specialised cooperative kernel.
//
// Replaces, for a whole run of consecutive steps, the body of Master.fit's batch loop
// (core/Master.scala:179-198) together with the slave's gradient request (core/Slave.scala:142-157):
// no launch, no host round trip and exactly ONE grid-wide barrier per SGD step.
//
// Measured facts this design answers (profiles/r1a, r1b; tools/microbench.cu on a B200):
//   L2 hit 307 cycles; a gpu-scope release/acquire grid barrier ~2300 cycles; a two-kernel step 25 us for
//   197 KB of row windows.  The step is a chain of dependent latencies, so the kernel removes links:
//   * PRODUCER warp (one per CTA): row windows do not depend on the weights, so it walks the sample ids
//     kStages steps ahead -- ids -> row pointers -> one TMA bulk copy (cp.async.bulk, mbarrier
//     complete_tx) per row into the stage's shared-memory partition, plus a chunk list.  Full/empty
//     mbarriers per stage, the classic TMA pipeline.
//   * CONSUMER warps: the CTA's rows of a step are cut into 128-pair chunks dealt round-robin to the
//     warps, so one 2000-nnz row does not serialise a warp (the step time is the MAX over rows).  Pass 1:
//     partial dots per chunk (fixed order -> deterministic); pass 2: gate per row, RED y*x into g.
//   * UPDATE warps: weights are double-buffered and gradients triple-buffered in L2 so the update of step
//     t-1 and the gradient of step t share one barrier interval: consumers read W_{t-1}[col], g_{t-1}[col]
//     and apply the update arithmetic themselves ("on the fly") while the update warps write the same
//     values into the W_t buffer for the interval after, zero the buffer g_{t+1} will use and reduce
//     c_t = 2*lambda*(W_t . d) and ||W_t||^2 to per-CTA partials (summed in a fixed order by one warp per
//     CTA and handed to the others through shared memory).
//
// Interval I_t (between grid barrier t-1 and t), with W_t the weights step t differentiates at:
//   consumers: x.W_t with W_t[col] = update(W_{t-1}[col], g_{t-1}[col], c_{t-1}); gate; RED into g_t
//   updaters : W_t buffer <- update(W_{t-1}, g_{t-1}, c_{t-1}); zero g_{t+1}'s buffer; partials of c_t, ||W_t||^2
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dsgd_kernels.cuh"

namespace dsgd {

constexpr int kMaxWorld = 16;

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
  double *partial;  // [2][gridDim.x][2]
  unsigned *hinge;  // [n_steps], zero on entry
  double *losses;   // [n_steps] or nullptr
  double *w_out;    // resident weights after the last step
  float *w32_out;
  double *scal;     // kScalC / kScalNrm2 of the resident weights
  unsigned *bar;    // grid barrier counter, zero on entry
  int *abort_flag;  // set to 1 if a wait hit the watchdog
  double lambda, lr, k_den;
  long long timeout_cycles;
  long long *tl;    // debug timeline: [256 steps][16 stamps] of clock64 (CTA 0), or nullptr
  // ---- multi-GPU exchange over peer memory (world > 1): every rank's gradient buffers and flag words are mapped
  //      into every other rank's address space (cudaIpc / peer access over NVLink) ----
  int world, rank;
  int64_t step_base;                           // global step number of this launch's first step (same on all ranks)
  double *xg[3];                               // this rank's gradient buffers (local, dim + 8 doubles each)
  double *xrecv[kMaxWorld];                    // xrecv[k]: receive area of rank k: [sender][parity][dim + 8]; [rank] is local
  unsigned long long *xflag[kMaxWorld];        // xflag[k]: flag words of rank k: [sender][cta]; [rank] is local
  int xstride;                                 // dim + 8
  unsigned long long *llw[2];                  // mode 3: weights as LL words (16 B per column), double-buffered by step parity
  // ---- experimental variants (template parameter kOpt; not the default path) ----
  unsigned *bar_flags;                         // kOpt & 1: release flags of the grid barrier, one 128-byte line per kBarGroup CTAs, zero on entry
  double *push;                                // kOpt & 2: pushed partials [2 parities][dest CTA][src CTA][2]
  unsigned long long *xllw[kMaxWorld][2];      // mode 4: every rank's LL weight buffers (inside its exported block); [rank] is local
};
static_assert(sizeof(PersistParams) <= 4000, "kernel parameter space is 4 KB");

// ---- PTX helpers: mbarrier + TMA bulk copy -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
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
// Bounded wait: if the phase does not complete within `timeout` cycles (or somebody already raised the abort
// flag) the flag is raised and the caller carries on -- results are then garbage, but nothing deadlocks and the
// host turns the flag into DSGD_ERR_TIMEOUT.
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity, int *abort_flag, long long timeout) {
  if (mbar_try_wait(b, parity)) return;
  const long long t0 = clock64();
  unsigned spins = 0;
  while (!mbar_try_wait(b, parity)) {
    if ((++spins & 255u) == 0u && (clock64() - t0 > timeout || *(volatile int *)abort_flag)) {
      *(volatile int *)abort_flag = 1;
      return;
    }
  }
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// "LL" element of the cross-GPU exchange: a double travels as two 8-byte words {low 32 bits, tag} and
// {high 32 bits, tag}.  An aligned 8-byte store is single-copy atomic, so a word whose tag matches carries valid
// data: no fence, no separate flag, one one-way NVLink store per word (the scheme of NCCL's LL protocol).
__device__ __forceinline__ void ll_store(unsigned long long *dst, double v, unsigned tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (bits & 0xffffffffull) | ((unsigned long long)tag << 32);
  const unsigned long long w1 = (bits >> 32) | ((unsigned long long)tag << 32);
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ll_try_load(const unsigned long long *src, unsigned tag, double &v) {
  unsigned long long w0, w1;
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
  if ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag) return false;
  v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
  return true;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// w_j after one SGD update given the raw gradient-sum entry (same arithmetic as k_update<true>):
// regularize on the surviving key (SparseSVM.scala:31), mean over workers, times lr, subtract, each with the
// Sparse constructor's 1e-20 filter (core/Master.scala:194,197; math/Sparse.scala:108-118).
__device__ __forceinline__ double apply_update(double wv, double graw, double c, bool add_c, double k_den, double lr) {
  double v = filt(graw);
  if (v != 0.0) {
    if (add_c) v = filt(v + c);
    if (v != 0.0) {
      const double mean = (k_den == 1.0) ? v : filt(v / k_den);  // x / 1.0 == x exactly
      const double step = filt(mean * lr);
      wv = filt(wv - step);
    }
  }
  return wv;
}

// Fixed-order sums of the per-CTA partials {c-dot, ||w||^2} (2 doubles per CTA); the same in every CTA.
// All loads are issued before the first add (up to kPartLoads per lane: covers 160 CTAs), so the cost is one
// L2 round trip plus the shuffle tree rather than one round trip per 32 CTAs.
constexpr int kPartLoads = 5;
__device__ __forceinline__ void sum_partials2(const double *p, int n_cta, int lane, double &s0, double &s1) {
  double2 v[kPartLoads];
#pragma unroll
  for (int i = 0; i < kPartLoads; ++i) {
    const int b = lane + 32 * i;
    v[i] = make_double2(0.0, 0.0);
    if (b < n_cta) v[i] = __ldcg(reinterpret_cast<const double2 *>(p) + b);
  }
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int i = 0; i < kPartLoads; ++i) { a0 += v[i].x; a1 += v[i].y; }
  for (int b = lane + 32 * kPartLoads; b < n_cta; b += 32) {  // more than 160 CTAs: not on a B200
    const double2 w2 = __ldcg(reinterpret_cast<const double2 *>(p) + b);
    a0 += w2.x; a1 += w2.y;
  }
  s0 = warp_sum(a0);
  s1 = warp_sum(a1);
}

// One grid-wide barrier among the barrier-synchronised warps of every CTA (the producer warp stays out):
// CTA-level named barrier, one release arrival, relaxed polling, one acquire fence.
// `target` = number of arrivals that completes this phase.  Returns false if the watchdog fired.
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool grid_barrier(unsigned *bar, unsigned target, int *abort_flag, long long timeout,
                                             int *smem_ok, int n_sync_threads, long long *tl = nullptr,
                                             bool tl_ns = false) {
  named_bar_sync(3, n_sync_threads);
  if (threadIdx.x == 0) {
    if (tl) tl[0] = tl_ns ? global_ns() : clock64();
    red_release_gpu_add(bar, 1u);
    int ok = 1;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (ld_relaxed_gpu(bar) < target) {
      if ((++spins & 1023u) == 0u) {
        if (clock64() - t0 > timeout || *(volatile int *)abort_flag) {
          *(volatile int *)abort_flag = 1;
          ok = 0;
          break;
        }
      }
    }
    fence_acq_rel_gpu();
    *smem_ok = ok;
    if (tl) tl[1] = tl_ns ? global_ns() : clock64();
  }
  named_bar_sync(3, n_sync_threads);
  return *(volatile int *)smem_ok != 0;
}

// ---- variant kOpt & 1: grid barrier with separate arrival counter and release flags -------------------------------
// In grid_barrier() every CTA polls the word the arrivals are added to, so 148 pollers and 148 arrivals queue on one
// L2 line (measured: 1.2-1.6 us from the last arrival to the release, profiles/r1c_summary.md).  Here nobody polls the
// counter: the arrival is an atom that returns the count, the LAST arriver raises one flag per group of kBarGroup
// CTAs (each on its own 128-byte line) and everybody else polls only its group's flag.
constexpr int kBarGroup = 8;
constexpr int kBarFlagStride = 32;  // unsigned words per flag line
__device__ __forceinline__ unsigned atom_acq_rel_gpu_add(unsigned *p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned *p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool grid_barrier_flags(unsigned *bar, unsigned *flags, unsigned phase, unsigned n_cta,
                                                   int *abort_flag, long long timeout, int *smem_ok, int n_sync_threads,
                                                   long long *tl = nullptr, bool tl_ns = false) {
  named_bar_sync(3, n_sync_threads);
  if (threadIdx.x == 0) {
    if (tl) tl[0] = tl_ns ? global_ns() : clock64();
    int ok = 1;
    const unsigned old = atom_acq_rel_gpu_add(bar, 1u);  // release: this CTA's writes; acquire: every earlier arrival's
    if (old + 1u == phase * n_cta) {
      fence_acq_rel_gpu();                               // fence + relaxed stores: a release pattern per flag
      const unsigned n_groups = (n_cta + kBarGroup - 1) / kBarGroup;
      for (unsigned g = 0; g < n_groups; ++g) st_relaxed_gpu(flags + g * kBarFlagStride, phase);
    } else {
      const unsigned *f = flags + (blockIdx.x / kBarGroup) * kBarFlagStride;
      const long long t0 = clock64();
      unsigned spins = 0;
      while ((int)(ld_relaxed_gpu(f) - phase) < 0) {     // the flag only ever steps from phase - 1 to phase
        if ((++spins & 1023u) == 0u) {
          if (clock64() - t0 > timeout || *(volatile int *)abort_flag) {
            *(volatile int *)abort_flag = 1;
            ok = 0;
            break;
          }
        }
      }
      fence_acq_rel_gpu();
    }
    *smem_ok = ok;
    if (tl) tl[1] = tl_ns ? global_ns() : clock64();
  }
  named_bar_sync(3, n_sync_threads);
  return *(volatile int *)smem_ok != 0;
}

constexpr int kChunkPairs = 128;             // 4 pairs per lane per chunk
constexpr uint32_t kChunkGlobal = 1u << 31;  // chunk offset flag: read from global, the row did not fit the stage
constexpr int kMaxRowsPerCta = 32;           // rows of one step per CTA (one producer lane each)

template <int kMaxChunks>
struct StageMeta {
  int n_rows;
  int n_chunks;
  int row_y[kMaxRowsPerCta];
  uint32_t row_b[kMaxRowsPerCta];    // window start (16-byte units) -- for rows that missed the chunk list
  int row_len[kMaxRowsPerCta];       // pairs, padding included
  short row_first[kMaxRowsPerCta];   // first chunk of the row
  short row_nch[kMaxRowsPerCta];     // chunks of the row; -1: not in the chunk list (whole-row slow path)
  uint32_t ch_off[kMaxChunks];       // pair offset inside the stage partition, or kChunkGlobal | global pair index
  short ch_n[kMaxChunks];            // pairs in the chunk (<= kChunkPairs)
  short ch_row[kMaxChunks];          // local row
  double part[kMaxChunks];           // pass-1 partial dot of the chunk
};

template <int kCons, int kUpd, int kStages, int kStagePairs, int kMaxChunks>
struct PersistSmem {
  uint2 ring[kStages][kStagePairs];
  StageMeta<kMaxChunks> meta[kStages];
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t c_bar[2];
  double c_val[2];
  double nrm_val[2];
  double red[kUpd][2];
  double red_all[kCons + kUpd][2];
  unsigned hinge_acc;
  int ok;
};

#define DSGD_TL(slot_)                                                                        \
  do {                                                                                        \
    if (p.tl && blockIdx.x == 0 && lane == 0 && t < 256) p.tl[t * 16 + (slot_)] = clock64(); \
  } while (0)

// kOpt (mode 0 only): bit 2 = single-chunk rows are gated and scattered inside pass 1 (no partial, no second pass);
// bit 0 = grid_barrier_flags instead of grid_barrier; bit 1 = per-CTA partials of c / ||w||^2 are
// PUSHED to a private area of every CTA instead of 148 CTAs reading the same 2.4 KB (measured: c is handed over 1 880
// cycles after the barrier).  Same values summed in the same order: results are bit-identical to kOpt == 0.
template <int kCons, int kUpd, int kStages, int kStagePairs, int kMaxChunks, int kMode, int kOpt = 0>
__global__ void __launch_bounds__((kCons + kUpd + 1) * 32, 1) k_sync_persistent(const PersistParams p) {
  using Smem = PersistSmem<kCons, kUpd, kStages, kStagePairs, kMaxChunks>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const bool is_cons = warp < kCons;
  const bool is_upd = warp >= kCons && warp < kCons + kUpd;
  const int G = gridDim.x;
  const int B = p.batch;
  const int64_t S = p.n_steps;
  constexpr int kSyncThreads = (kCons + kUpd) * 32;
  // rows of a step owned by this CTA: i = blockIdx.x + m * G  (a small batch is spread over all CTAs); <= 32
  const int n_r = (B > (int)blockIdx.x) ? (B - 1 - (int)blockIdx.x) / G + 1 : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1u);
      mbar_init(&sm.empty[s], (unsigned)kCons);
    }
    mbar_init(&sm.c_bar[0], 1u);
    mbar_init(&sm.c_bar[1], 1u);
    sm.hinge_acc = 0u;
    sm.ok = 1;
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  // =========================================================================================================
  // PRODUCER warp: runs ahead of everybody else, bounded only by the empty[] barriers.  Lane m owns row m.
  // =========================================================================================================
  if (!is_cons && !is_upd) {
    auto load_id = [&](int64_t t) -> int32_t {
      return (t < S && lane < n_r) ? __ldg(&p.samples[t * B + blockIdx.x + lane * G]) : -1;
    };
    uint32_t b0 = 0, e0 = 0, b1 = 0, e1 = 0;
    int y0 = 0, y1 = 0;
    auto load_win = [&](int32_t id, uint32_t &b, uint32_t &e, int &y) {
      b = 0u; e = 0u; y = 0;
      if (id >= 0) {
        b = __ldg(&p.rp16[id]);
        e = __ldg(&p.rp16[id + 1]);
        y = (int)__ldg(&p.label[id]);
      }
    };
    load_win(load_id(0), b0, e0, y0);   // window of step t      (stage C input)
    load_win(load_id(1), b1, e1, y1);   // window of step t + 1  (stage B)
    int32_t id_next = load_id(2);       // sample id of step t + 2 (stage A)
    for (int64_t t = 0; t < S; ++t) {
      const int st = (int)(t % kStages);
      if (t >= kStages) {
        mbar_wait(&sm.empty[st], (unsigned)(((t / kStages) - 1) & 1), p.abort_flag, p.timeout_cycles);
        if (*(volatile int *)p.abort_flag) return;  // the barrier-synchronised warps gave up (watchdog)
      }
      auto &mt = sm.meta[st];
      // lay the rows out: exclusive scans over the CTA's rows of pairs and chunks
      const int len = (lane < n_r) ? (int)(e0 - b0) * 2 : 0;
      const int nch = (len + kChunkPairs - 1) / kChunkPairs;
      int ps = len, cs = nch;  // inclusive warp scans
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, ps, o), c2 = __shfl_up_sync(0xffffffffu, cs, o);
        if (lane >= o) { ps += a; cs += c2; }
      }
      const int my_pair = ps - len, my_chunk = cs - nch;
      const bool listed = (my_chunk + nch) <= kMaxChunks;           // prefix property: later rows miss too
      const bool in_ring = listed && (my_pair + len) <= kStagePairs;
      if (lane < n_r) {
        mt.row_y[lane] = y0;
        mt.row_b[lane] = b0;
        mt.row_len[lane] = len;
        mt.row_first[lane] = (short)my_chunk;
        mt.row_nch[lane] = (short)(listed ? nch : -1);
        if (listed) {
          for (int c = 0; c < nch; ++c) {
            const int n = min(kChunkPairs, len - c * kChunkPairs);
            mt.ch_off[my_chunk + c] = in_ring ? (uint32_t)(my_pair + c * kChunkPairs)
                                              : (kChunkGlobal | (b0 * 2u + (uint32_t)(c * kChunkPairs)));
            mt.ch_n[my_chunk + c] = (short)n;
            mt.ch_row[my_chunk + c] = (short)lane;
          }
        }
      }
      const unsigned my_bytes = (lane < n_r && in_ring) ? (unsigned)len * 8u : 0u;
      unsigned ring_bytes = my_bytes;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ring_bytes += __shfl_xor_sync(0xffffffffu, ring_bytes, o);
      // chunks actually written to the list: everything up to the first row that did not fit it
      int listed_chunks = (lane < n_r && listed) ? (my_chunk + nch) : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) listed_chunks = max(listed_chunks, __shfl_xor_sync(0xffffffffu, listed_chunks, o));
      if (lane == 0) {
        mt.n_rows = n_r;
        mt.n_chunks = listed_chunks;
      }
      __syncwarp();  // every lane's metadata is written before lane 0 arrives on the full barrier
      if (lane == 0) {
        if (ring_bytes) mbar_expect_tx(&sm.full[st], ring_bytes);
        else mbar_arrive(&sm.full[st]);  // metadata only: complete the phase
      }
      __syncwarp();
      if (my_bytes) bulk_g2s(&sm.ring[st][my_pair], p.pairs + (size_t)b0 * 2, my_bytes, &sm.full[st]);
      // advance the register pipeline
      b0 = b1; e0 = e1; y0 = y1;
      load_win(id_next, b1, e1, y1);
      id_next = load_id(t + 3);
    }
    return;
  }

  if constexpr (kMode == 4) {
    // =======================================================================================================
    // EXPERIMENTAL (DSGD_P2P_MODE=4; written after the round's GPU budget ran out, never run): mode 3 with COLUMN
    // OWNERSHIP.  Mode 3 stores every rank's whole dense gradient into every peer ((K-1) * 756 KB per rank and step:
    // 5.3 MB at K = 8, >= 5.9 us of NVLink time) and every rank repeats the K-way reduction for all columns.  Here rank r
    // owns columns [r * cpr, (r + 1) * cpr), cpr = ceil(dim / K):
    //   everybody : push this CTA's column slice of g_{T-1} as LL words (tag T), each column ONLY to its owner
    //   owners    : (update warps) per owned column: the K replies of step T-1 (own from local g_{T-1}, the peers' from
    //               the receive area), regularized on their own support and folded in rank order, W_T = W_{T-1} -
    //               lr*sum/K, stored as an LL word (tag T+1) into the weight buffer of parity T of EVERY rank
    //   collectors: (update warps, after their owner duty) read W_T over the CTA's slice of ALL columns as the words
    //               arrive: partials of c_T and ||W_T||^2 in the same order on every rank; zero g_{T+1}'s buffer.  Nobody
    //               leaves interval T before every owner delivered all of W_T: that is the flow control that keeps
    //               two receive parities sufficient.
    //   consumers : as in mode 3 -- one LL gather of W_T[col] per non-zero, spinning on the tag
    // The packed counter slot [dim] (hinge + 2^32 * samples) still goes to every peer: every rank reports the loss.
    // Bytes per rank and step: 2 * (K-1)/K * 756 KB (1.3 MB at K = 8) instead of (K-1) * 756 KB.
    // =======================================================================================================
    const int K = p.world, me = p.rank;
    const int slice = (p.dim + 1 + G - 1) / G;   // columns per CTA, plus ONE counter slot [dim] = hinge + 2^32 * samples
    const int j_lo = min(blockIdx.x * slice, p.dim + 1), j_hi = min(j_lo + slice, p.dim + 1);
    const int par_stride = p.xstride, snd_stride = 2 * p.xstride;
    const double lr = p.lr, kd = (double)K;
    const int64_t base = p.step_base;
    unsigned phase = 0;
    const int cpr = (p.dim + K - 1) / K;                                        // columns per owner
    const int r_lo = min(me * cpr, p.dim), r_hi = min(r_lo + cpr, p.dim);      // this rank's columns
    const int oslice = (cpr + G - 1) / G;                                       // ... of which this CTA's
    const int o_lo = min(r_lo + (int)blockIdx.x * oslice, r_hi), o_hi = min(o_lo + oslice, r_hi);
    const int ut = (int)threadIdx.x - kCons * 32;                               // index among the update threads (< 0: consumer)

    for (int64_t T = base; T <= base + S; ++T) {
      const int64_t t = T - base;
      const bool first = (T == base), last = (T == base + S);
      const unsigned long long *LWprev = p.xllw[me][(T + 1) & 1];  // LL words of W_{T-1}, tag T
      unsigned long long *LWcur = p.xllw[me][T & 1];               // LL words of W_T, tag T+1
      const double *Gprev = p.xg[(T + 2) % 3];                // g_{T-1}
      double *Gcur = p.xg[T % 3];
      double *Gzero = p.xg[(T + 1) % 3];
      const int parp = (int)((T + 1) & 1);                    // receive parity of step T-1
      const unsigned gtag = (unsigned)T;                      // g words of step T-1 carry tag T
      const unsigned wtag = (unsigned)(T + 1);                // W_T words carry tag T+1
      const unsigned long long *rcv = reinterpret_cast<const unsigned long long *>(p.xrecv[me]) + 2 * (size_t)parp * par_stride;
      const double *part_prev = p.partial + (size_t)((T + 1) & 1) * G * 2;   // partials of W_{T-1}
      double *part_cur = p.partial + (size_t)(T & 1) * G * 2;
      const unsigned c_par = (unsigned)((t >> 1) & 1);
      bool ok = true;
      auto spin_ll = [&](const unsigned long long *src, unsigned tag, double &v) {
        unsigned spins = 0;
        const long long t0 = clock64();
        while (!ll_try_load(src, tag, v)) {
          if ((++spins & 255u) == 0u && (clock64() - t0 > p.timeout_cycles || *(volatile int *)p.abort_flag)) {
            *(volatile int *)p.abort_flag = 1;
            ok = false;
            v = 0.0;
            break;
          }
        }
      };

      if (warp == 0) DSGD_TL(0);
      // ---- push g_{T-1} (every sync warp; one column per thread) ----
      if (!first) {
        for (int j = j_lo + threadIdx.x; j < j_hi; j += kSyncThreads) {
          const double v = __ldcg(&Gprev[j]);
          const size_t slot = 2 * ((size_t)me * snd_stride + (size_t)parp * par_stride + j);
          if (j == p.dim) {                                     // counters: to everybody
            for (int k = 0; k < K; ++k)
              if (k != me) ll_store(reinterpret_cast<unsigned long long *>(p.xrecv[k]) + slot, v, gtag);
          } else {
            const int o = j / cpr;                              // gradient entries: to the column's owner only
            if (o != me) ll_store(reinterpret_cast<unsigned long long *>(p.xrecv[o]) + slot, v, gtag);
          }
        }
      }

      if (warp == 0) DSGD_TL(1);
      // ---- c_{T-1}: summed by update warp 0, handed to every sync warp of the CTA through shared memory ----
      double c_prev = 0.0;                                    // of W_{T-1}
      if (warp == kCons) {
        double nrm_prev = 0.0;
        if (!first) {
          if (T - 1 == base) {
            c_prev = p.scal[kScalC];                          // W_base came from the host: k_prepare / previous launch
            nrm_prev = p.scal[kScalNrm2];
          } else {
            double sd, sn;
            sum_partials2(part_prev, G, lane, sd, sn);
            c_prev = p.lambda * 2.0 * sd;
            nrm_prev = sn;
          }
        }
        if (lane == 0) {
          sm.c_val[t & 1] = c_prev;
          sm.nrm_val[t & 1] = nrm_prev;                       // ||W_{T-1}||^2 for the loss of step T-1
          mbar_arrive(&sm.c_bar[t & 1]);
        }
        __syncwarp();
      } else if (is_upd) {                                      // consumers never need c here
        mbar_wait(&sm.c_bar[t & 1], c_par, p.abort_flag, p.timeout_cycles);
        c_prev = sm.c_val[t & 1];
      }
      if (warp == 0) DSGD_TL(2);
      // ---- first interval: W_base (plain doubles from the host, identical on every rank) goes into LL form locally ----
      double pd = 0.0, pn = 0.0;
      if (first) {
        const int j = j_lo + threadIdx.x;                     // slice <= kSyncThreads columns (checked on the host)
        if (j < j_hi) {
          if (j < p.dim) ll_store(LWcur + 2 * (size_t)j, __ldcg(&p.wbuf[0][j]), wtag);
          Gzero[j] = 0.0;
        }
      } else if (is_upd) {
        const bool add_c = (c_prev != 0.0) && (fabs(c_prev) > kEps);
        // ---- owner duty: reduce the K replies of every owned column of this CTA, update, deliver W_T to every rank ----
        for (int jo = o_lo + ut; jo < o_hi; jo += kUpd * 32) {
          double raw[kMaxWorld];
          bool got[kMaxWorld];
          double wn = 0.0;
          bool got_w = ll_try_load(LWprev + 2 * (size_t)jo, gtag, wn);            // own word of W_{T-1}[jo] carries tag T
#pragma unroll
          for (int k = 0; k < kMaxWorld; ++k) {
            got[k] = true;
            raw[k] = 0.0;
            if (k < K) {
              if (k == me) raw[k] = __ldcg(&Gprev[jo]);
              else got[k] = ll_try_load(rcv + 2 * ((size_t)k * snd_stride + jo), gtag, raw[k]);
            }
          }
          if (!got_w) spin_ll(LWprev + 2 * (size_t)jo, gtag, wn);
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < kMaxWorld; ++k) {
            if (k < K) {
              if (!got[k]) spin_ll(rcv + 2 * ((size_t)k * snd_stride + jo), gtag, raw[k]);
              double v = filt(raw[k]);
              if (v != 0.0 && add_c) v = filt(v + c_prev);
              s = (k == 0) ? v : filt(s + v);
            }
          }
          if (s != 0.0) {
            const double mean = filt(s / kd);
            const double step = filt(mean * lr);
            wn = filt(wn - step);
          }
          for (int k = 0; k < K; ++k) ll_store(p.xllw[k][T & 1] + 2 * (size_t)jo, wn, wtag);   // [me] is LWcur
        }
        // ---- collector duty: W_T over the CTA's slice of ALL columns, as the owners' words arrive ----
        for (int j = j_lo + ut; j < j_hi; j += kUpd * 32) {
          if (j == p.dim) {
            double s = 0.0;                                     // packed counters of step T-1: plain sum over the ranks
            for (int k = 0; k < K; ++k) {
              double v;
              if (k == me) v = __ldcg(&Gprev[j]);
              else spin_ll(rcv + 2 * ((size_t)k * snd_stride + j), gtag, v);
              s += v;
            }
            if (p.losses) {  // loss of step T-1 on W_{T-1}
              const double ns = floor(s / 4294967296.0);
              p.losses[t - 1] = p.lambda * sm.nrm_val[t & 1] + (s - ns * 4294967296.0) / ns;
            }
          } else {
            double wn;
            spin_ll(LWcur + 2 * (size_t)j, wtag, wn);
            pd += filt(wn * __ldg(&p.d[j]));
            pn += wn * wn;
          }
          Gzero[j] = 0.0;
        }
      }
      pd = warp_sum(pd);
      pn = warp_sum(pn);
      if (lane == 0) { sm.red_all[warp][0] = pd; sm.red_all[warp][1] = pn; }   // summed by thread 0 before the grid barrier
      if (warp == 0) DSGD_TL(3);

      if (is_cons) {
        if (!last) {
          const int st = (int)(t % kStages);
          auto &mt = sm.meta[st];
          mbar_wait(&sm.full[st], (unsigned)((t / kStages) & 1), p.abort_flag, p.timeout_cycles);
          if (warp == 0) DSGD_TL(4);
          const int n_ch = mt.n_chunks;
          const uint2 *ring = &sm.ring[st][0];
          for (int c = warp; c < n_ch; c += kCons) {
            const uint32_t off = mt.ch_off[c];
            const int n = mt.ch_n[c];
            const uint2 *src = (off & kChunkGlobal) ? (p.pairs + (off & ~kChunkGlobal)) : (ring + off);
            uint2 pr[4];
            double wv[4];
            bool got[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int k = u * 32 + lane;
              pr[u] = (k < n) ? src[k] : make_uint2(0u, 0u);   // col 0 / val 0: inert, still a valid gather
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) got[u] = ll_try_load(LWcur + 2 * (size_t)pr[u].x, wtag, wv[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (!got[u]) spin_ll(LWcur + 2 * (size_t)pr[u].x, wtag, wv[u]);
            double acc = 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += filt(filt((double)__uint_as_float(pr[u].y)) * wv[u]);
            acc = warp_sum(acc);
            if (lane == 0) mt.part[c] = acc;
          }
          if (warp == 0) DSGD_TL(5);
          named_bar_sync(2, kCons * 32);
          unsigned hinge = 0;
          for (int c = warp; c < n_ch; c += kCons) {
            const int row = mt.ch_row[c];
            const int firstc = mt.row_first[row], nch = mt.row_nch[row];
            double dot = 0.0;
            for (int i = 0; i < nch; ++i) dot += mt.part[firstc + i];
            const int yi = mt.row_y[row];
            const double y = (double)yi;
            if (c == firstc && lane == 0) hinge += (unsigned)(1 - yi * ((dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0)));
            if (!(y * dot < 0.0)) {
              const uint32_t off = mt.ch_off[c];
              const int n = mt.ch_n[c];
              const uint2 *src = (off & kChunkGlobal) ? (p.pairs + (off & ~kChunkGlobal)) : (ring + off);
              for (int k = lane; k < n; k += 32) {
                const uint2 pr = src[k];
                const double gvv = filt(filt((double)__uint_as_float(pr.y)) * y);
                if (gvv != 0.0) atomicAdd(&Gcur[pr.x], gvv);
              }
            }
          }
          for (int m = warp; m < mt.n_rows; m += kCons) {
            const int nch = mt.row_nch[m];
            if (nch == 0) {
              if (lane == 0) hinge += 1u;
            } else if (nch < 0) {  // row outside the chunk list: whole row from global memory
              const uint2 *grow = p.pairs + (size_t)mt.row_b[m] * 2;
              const int len = mt.row_len[m];
              double acc = 0.0;
              for (int k = lane; k < len; k += 32) {
                const uint2 pr = __ldg(&grow[k]);
                double wv;
                spin_ll(LWcur + 2 * (size_t)pr.x, wtag, wv);
                acc += filt(filt((double)__uint_as_float(pr.y)) * wv);
              }
              const double dot = warp_sum(acc);
              const int yi = mt.row_y[m];
              const double y = (double)yi;
              if (lane == 0) hinge += (unsigned)(1 - yi * ((dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0)));
              if (!(y * dot < 0.0))
                for (int k = lane; k < len; k += 32) {
                  const uint2 pr = __ldg(&grow[k]);
                  const double gvv = filt(filt((double)__uint_as_float(pr.y)) * y);
                  if (gvv != 0.0) atomicAdd(&Gcur[pr.x], gvv);
                }
            }
          }
          if (lane == 0 && hinge) atomicAdd(&sm.hinge_acc, hinge);
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.empty[st]);
          if (warp == 0) DSGD_TL(8);
        }
      }
      if (!ok) *(volatile int *)&sm.ok = 0;
      // ---- grid barrier T (the CTA's hinge total and batch ride in slot [dim] of g_T) ----
      named_bar_sync(3, kSyncThreads);
      if (*(volatile int *)&sm.ok == 0) { *(volatile int *)p.abort_flag = 1; }
      if (threadIdx.x == 0 && !first) {   // per-CTA partials of c_T, ||W_T||^2: warps in index order (deterministic)
        double sd = 0.0, sn = 0.0;
#pragma unroll
        for (int i = 0; i < kCons + kUpd; ++i) { sd += sm.red_all[i][0]; sn += sm.red_all[i][1]; }
        part_cur[2 * blockIdx.x] = sd;
        part_cur[2 * blockIdx.x + 1] = sn;
      }
      if (threadIdx.x == 0 && !last) {
        const unsigned h = sm.hinge_acc;
        if (h) { atomicAdd(&Gcur[p.dim], (double)h); sm.hinge_acc = 0u; }
        if (blockIdx.x == 0) atomicAdd(&Gcur[p.dim], (double)B * 4294967296.0);
      }
      ++phase;
      if (!grid_barrier(p.bar, phase * (unsigned)G, p.abort_flag, p.timeout_cycles, &sm.ok, kSyncThreads,
                        (p.tl && blockIdx.x == 0 && t < 256) ? p.tl + t * 16 + 6 : nullptr))
        return;
      if (*(volatile int *)p.abort_flag) return;
    }
    // epilogue: W_{base+S} sits in LL form (tag base+S+1) in llw[(base+S) & 1]; publish it as plain resident weights
    {
      const unsigned long long *LW = p.xllw[me][(base + S) & 1];
      const unsigned wtag = (unsigned)(base + S + 1);
      const int n_all = G * kSyncThreads;
      for (int j = blockIdx.x * kSyncThreads + threadIdx.x; j < p.dim; j += n_all) {
        double wv = 0.0;
        ll_try_load(LW + 2 * (size_t)j, wtag, wv);             // complete: written before the last grid barrier
        p.w_out[j] = wv;
        p.w32_out[j] = (float)wv;
      }
      if (blockIdx.x == 0 && warp == 0 && S > 0) {
        double sd, sn;
        sum_partials2(p.partial + (size_t)((base + S) & 1) * G * 2, G, lane, sd, sn);
        if (lane == 0) { p.scal[kScalC] = p.lambda * 2.0 * sd; p.scal[kScalNrm2] = sn; }
      }
    }
    return;
  }

  if constexpr (kMode == 3) {
    // =======================================================================================================
    // world > 1, one grid barrier per step, weights handed from the update warps to the consumers as LL words.
    // Interval I_T (between grid barrier T-1 and T), W_T = weights step T differentiates at:
    //   everybody : push this CTA's column slice of g_{T-1} to every peer's receive area as LL words (tag T)
    //   updaters  : per column of the CTA's slice: wait for the K replies of step T-1 (own from local g_{T-1}, the
    //               peers' from the receive area), regularize each on its own support and fold them in rank order
    //               (core/Master.scala:194), W_T = W_{T-1} - lr*sum/K, published as an LL word (tag T+1)
    //               into the weight buffer of parity T; partials of c_T, ||W_T||^2; loss of step T-1; zero g_{T+1}'s buffer
    //   consumers : ONE 16-byte gather per non-zero -- the LL word of W_T[col] (spinning on its tag if the column's
    //               update has not landed yet); x.W_T; gate; RED y*x into g_T
    // The K-way reduction is done once per column instead of once per gathered non-zero, and nothing but the
    // updaters ever waits for c.  A peer can be at most one interval ahead: two parities suffice everywhere.
    // =======================================================================================================
    const int K = p.world, me = p.rank;
    const int slice = (p.dim + 1 + G - 1) / G;   // columns per CTA, plus ONE counter slot [dim] = hinge + 2^32 * samples
    const int j_lo = min(blockIdx.x * slice, p.dim + 1), j_hi = min(j_lo + slice, p.dim + 1);
    const int par_stride = p.xstride, snd_stride = 2 * p.xstride;
    const double lr = p.lr, kd = (double)K;
    const int64_t base = p.step_base;
    unsigned phase = 0;

    for (int64_t T = base; T <= base + S; ++T) {
      const int64_t t = T - base;
      const bool first = (T == base), last = (T == base + S);
      const unsigned long long *LWprev = p.llw[(T + 1) & 1];  // LL words of W_{T-1}, tag T
      unsigned long long *LWcur = p.llw[T & 1];               // LL words of W_T, tag T+1
      const double *Gprev = p.xg[(T + 2) % 3];                // g_{T-1}
      double *Gcur = p.xg[T % 3];
      double *Gzero = p.xg[(T + 1) % 3];
      const int parp = (int)((T + 1) & 1);                    // receive parity of step T-1
      const unsigned gtag = (unsigned)T;                      // g words of step T-1 carry tag T
      const unsigned wtag = (unsigned)(T + 1);                // W_T words carry tag T+1
      const unsigned long long *rcv = reinterpret_cast<const unsigned long long *>(p.xrecv[me]) + 2 * (size_t)parp * par_stride;
      const double *part_prev = p.partial + (size_t)((T + 1) & 1) * G * 2;   // partials of W_{T-1}
      double *part_cur = p.partial + (size_t)(T & 1) * G * 2;
      const unsigned c_par = (unsigned)((t >> 1) & 1);
      bool ok = true;
      auto spin_ll = [&](const unsigned long long *src, unsigned tag, double &v) {
        unsigned spins = 0;
        const long long t0 = clock64();
        while (!ll_try_load(src, tag, v)) {
          if ((++spins & 255u) == 0u && (clock64() - t0 > p.timeout_cycles || *(volatile int *)p.abort_flag)) {
            *(volatile int *)p.abort_flag = 1;
            ok = false;
            v = 0.0;
            break;
          }
        }
      };

      if (warp == 0) DSGD_TL(0);
      // ---- push g_{T-1} (every sync warp; one column per thread) ----
      if (!first) {
        for (int j = j_lo + threadIdx.x; j < j_hi; j += kSyncThreads) {
          const double v = __ldcg(&Gprev[j]);
          for (int k = 0; k < K; ++k)
            if (k != me)
              ll_store(reinterpret_cast<unsigned long long *>(p.xrecv[k]) + 2 * ((size_t)me * snd_stride + (size_t)parp * par_stride + j), v, gtag);
        }
      }

      if (warp == 0) DSGD_TL(1);
      // ---- c_{T-1}: summed by update warp 0, handed to every sync warp of the CTA through shared memory ----
      double c_prev = 0.0;                                    // of W_{T-1}
      if (warp == kCons) {
        double nrm_prev = 0.0;
        if (!first) {
          if (T - 1 == base) {
            c_prev = p.scal[kScalC];                          // W_base came from the host: k_prepare / previous launch
            nrm_prev = p.scal[kScalNrm2];
          } else {
            double sd, sn;
            sum_partials2(part_prev, G, lane, sd, sn);
            c_prev = p.lambda * 2.0 * sd;
            nrm_prev = sn;
          }
        }
        if (lane == 0) {
          sm.c_val[t & 1] = c_prev;
          sm.nrm_val[t & 1] = nrm_prev;                       // ||W_{T-1}||^2 for the loss of step T-1
          mbar_arrive(&sm.c_bar[t & 1]);
        }
        __syncwarp();
      } else {
        mbar_wait(&sm.c_bar[t & 1], c_par, p.abort_flag, p.timeout_cycles);
        c_prev = sm.c_val[t & 1];
      }
      if (warp == 0) DSGD_TL(2);
      // ---- column update, one column per sync thread: all K replies and W_{T-1}[j] are requested at once ----
      double pd = 0.0, pn = 0.0;
      {
        const bool add_c = (c_prev != 0.0) && (fabs(c_prev) > kEps);
        const int j = j_lo + threadIdx.x;                     // slice <= kSyncThreads columns (checked on the host)
        if (j < j_hi) {
          if (first) {
            // W_base arrives as plain doubles from the host (wbuf): publish it in LL form, no update pending
            if (j < p.dim) ll_store(LWcur + 2 * (size_t)j, __ldcg(&p.wbuf[0][j]), wtag);
          } else {
            double raw[kMaxWorld];
            bool got[kMaxWorld];
            double wn = 0.0;
            bool got_w = true;
            if (j < p.dim) got_w = ll_try_load(LWprev + 2 * (size_t)j, gtag, wn);   // W_{T-1}[j] carries tag T
#pragma unroll
            for (int k = 0; k < kMaxWorld; ++k) {
              got[k] = true;
              raw[k] = 0.0;
              if (k < K) {
                if (k == me) raw[k] = __ldcg(&Gprev[j]);
                else got[k] = ll_try_load(rcv + 2 * ((size_t)k * snd_stride + j), gtag, raw[k]);
              }
            }
            if (!got_w) spin_ll(LWprev + 2 * (size_t)j, gtag, wn);
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < kMaxWorld; ++k) {
              if (k < K) {
                if (!got[k]) spin_ll(rcv + 2 * ((size_t)k * snd_stride + j), gtag, raw[k]);
                if (j == p.dim) {
                  s += raw[k];                                // packed counters: plain sum
                } else {
                  double v = filt(raw[k]);
                  if (v != 0.0 && add_c) v = filt(v + c_prev);
                  s = (k == 0) ? v : filt(s + v);
                }
              }
            }
            if (j == p.dim) {
              if (p.losses) {  // loss of step T-1 on W_{T-1}
                const double ns = floor(s / 4294967296.0);
                p.losses[t - 1] = p.lambda * sm.nrm_val[t & 1] + (s - ns * 4294967296.0) / ns;
              }
            } else {
              if (s != 0.0) {
                const double mean = filt(s / kd);
                const double step = filt(mean * lr);
                wn = filt(wn - step);
              }
              ll_store(LWcur + 2 * (size_t)j, wn, wtag);
              pd = filt(wn * __ldg(&p.d[j]));
              pn = wn * wn;
            }
          }
          Gzero[j] = 0.0;
        }
        pd = warp_sum(pd);
        pn = warp_sum(pn);
        if (lane == 0) { sm.red_all[warp][0] = pd; sm.red_all[warp][1] = pn; }   // summed by thread 0 before the grid barrier
      }
      if (warp == 0) DSGD_TL(3);

      if (is_cons) {
        if (!last) {
          const int st = (int)(t % kStages);
          auto &mt = sm.meta[st];
          mbar_wait(&sm.full[st], (unsigned)((t / kStages) & 1), p.abort_flag, p.timeout_cycles);
          if (warp == 0) DSGD_TL(4);
          const int n_ch = mt.n_chunks;
          const uint2 *ring = &sm.ring[st][0];
          for (int c = warp; c < n_ch; c += kCons) {
            const uint32_t off = mt.ch_off[c];
            const int n = mt.ch_n[c];
            const uint2 *src = (off & kChunkGlobal) ? (p.pairs + (off & ~kChunkGlobal)) : (ring + off);
            uint2 pr[4];
            double wv[4];
            bool got[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int k = u * 32 + lane;
              pr[u] = (k < n) ? src[k] : make_uint2(0u, 0u);   // col 0 / val 0: inert, still a valid gather
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) got[u] = ll_try_load(LWcur + 2 * (size_t)pr[u].x, wtag, wv[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (!got[u]) spin_ll(LWcur + 2 * (size_t)pr[u].x, wtag, wv[u]);
            double acc = 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += filt(filt((double)__uint_as_float(pr[u].y)) * wv[u]);
            acc = warp_sum(acc);
            if (lane == 0) mt.part[c] = acc;
          }
          if (warp == 0) DSGD_TL(5);
          named_bar_sync(2, kCons * 32);
          unsigned hinge = 0;
          for (int c = warp; c < n_ch; c += kCons) {
            const int row = mt.ch_row[c];
            const int firstc = mt.row_first[row], nch = mt.row_nch[row];
            double dot = 0.0;
            for (int i = 0; i < nch; ++i) dot += mt.part[firstc + i];
            const int yi = mt.row_y[row];
            const double y = (double)yi;
            if (c == firstc && lane == 0) hinge += (unsigned)(1 - yi * ((dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0)));
            if (!(y * dot < 0.0)) {
              const uint32_t off = mt.ch_off[c];
              const int n = mt.ch_n[c];
              const uint2 *src = (off & kChunkGlobal) ? (p.pairs + (off & ~kChunkGlobal)) : (ring + off);
              for (int k = lane; k < n; k += 32) {
                const uint2 pr = src[k];
                const double gvv = filt(filt((double)__uint_as_float(pr.y)) * y);
                if (gvv != 0.0) atomicAdd(&Gcur[pr.x], gvv);
              }
            }
          }
          for (int m = warp; m < mt.n_rows; m += kCons) {
            const int nch = mt.row_nch[m];
            if (nch == 0) {
              if (lane == 0) hinge += 1u;
            } else if (nch < 0) {  // row outside the chunk list: whole row from global memory
              const uint2 *grow = p.pairs + (size_t)mt.row_b[m] * 2;
              const int len = mt.row_len[m];
              double acc = 0.0;
              for (int k = lane; k < len; k += 32) {
                const uint2 pr = __ldg(&grow[k]);
                double wv;
                spin_ll(LWcur + 2 * (size_t)pr.x, wtag, wv);
                acc += filt(filt((double)__uint_as_float(pr.y)) * wv);
              }
              const double dot = warp_sum(acc);
              const int yi = mt.row_y[m];
              const double y = (double)yi;
              if (lane == 0) hinge += (unsigned)(1 - yi * ((dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0)));
              if (!(y * dot < 0.0))
                for (int k = lane; k < len; k += 32) {
                  const uint2 pr = __ldg(&grow[k]);
                  const double gvv = filt(filt((double)__uint_as_float(pr.y)) * y);
                  if (gvv != 0.0) atomicAdd(&Gcur[pr.x], gvv);
                }
            }
          }
          if (lane == 0 && hinge) atomicAdd(&sm.hinge_acc, hinge);
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.empty[st]);
          if (warp == 0) DSGD_TL(8);
        }
      }
      if (!ok) *(volatile int *)&sm.ok = 0;
      // ---- grid barrier T (the CTA's hinge total and batch ride in slot [dim] of g_T) ----
      named_bar_sync(3, kSyncThreads);
      if (*(volatile int *)&sm.ok == 0) { *(volatile int *)p.abort_flag = 1; }
      if (threadIdx.x == 0 && !first) {   // per-CTA partials of c_T, ||W_T||^2: warps in index order (deterministic)
        double sd = 0.0, sn = 0.0;
#pragma unroll
        for (int i = 0; i < kCons + kUpd; ++i) { sd += sm.red_all[i][0]; sn += sm.red_all[i][1]; }
        part_cur[2 * blockIdx.x] = sd;
        part_cur[2 * blockIdx.x + 1] = sn;
      }
      if (threadIdx.x == 0 && !last) {
        const unsigned h = sm.hinge_acc;
        if (h) { atomicAdd(&Gcur[p.dim], (double)h); sm.hinge_acc = 0u; }
        if (blockIdx.x == 0) atomicAdd(&Gcur[p.dim], (double)B * 4294967296.0);
      }
      ++phase;
      if (!grid_barrier(p.bar, phase * (unsigned)G, p.abort_flag, p.timeout_cycles, &sm.ok, kSyncThreads,
                        (p.tl && blockIdx.x == 0 && t < 256) ? p.tl + t * 16 + 6 : nullptr))
        return;
      if (*(volatile int *)p.abort_flag) return;
    }
    // epilogue: W_{base+S} sits in LL form (tag base+S+1) in llw[(base+S) & 1]; publish it as plain resident weights
    {
      const unsigned long long *LW = p.llw[(base + S) & 1];
      const unsigned wtag = (unsigned)(base + S + 1);
      const int n_all = G * kSyncThreads;
      for (int j = blockIdx.x * kSyncThreads + threadIdx.x; j < p.dim; j += n_all) {
        double wv = 0.0;
        ll_try_load(LW + 2 * (size_t)j, wtag, wv);             // complete: written before the last grid barrier
        p.w_out[j] = wv;
        p.w32_out[j] = (float)wv;
      }
      if (blockIdx.x == 0 && warp == 0 && S > 0) {
        double sd, sn;
        sum_partials2(p.partial + (size_t)((base + S) & 1) * G * 2, G, lane, sd, sn);
        if (lane == 0) { p.scal[kScalC] = p.lambda * 2.0 * sd; p.scal[kScalNrm2] = sn; }
      }
    }
    return;
  }

  if constexpr (kMode == 2) {
    // =======================================================================================================
    // world > 1, ONE grid barrier per step: the cross-GPU form of the "on the fly" scheme of the single-GPU loop.
    // Interval I_T (between grid barrier T-1 and T), W_T = weights step T differentiates at:
    //   everybody : push this CTA's column slice of g_{T-1} to every peer's receive area as LL words (tag T)
    //   consumers : W_T[col] = update(W_{T-1}[col], replies of all K workers for step T-1 at col) computed on the
    //               fly -- own reply from local g_{T-1}, the peers' from the receive area (waiting on the tag if a
    //               word has not landed yet); x.W_T; gate; RED y*x into g_T
    //   updaters  : the same reduction over their slice -> W_T buffer; partials of c_T, ||W_T||^2; loss of step T-1;
    //               zero the buffer g_{T+1} will use
    // A peer can be at most one interval ahead (it needs this rank's g_T before it can finish I_{T+1}), so two
    // receive parities suffice.  The first interval of a launch has no pending update (W_base is materialised).
    // =======================================================================================================
    const int K = p.world, me = p.rank;
    const int slice = (p.dim + 1 + G - 1) / G;   // columns per CTA, plus ONE counter slot [dim] = hinge + 2^32 * samples
    const int j_lo = min(blockIdx.x * slice, p.dim + 1), j_hi = min(j_lo + slice, p.dim + 1);
    const int par_stride = p.xstride, snd_stride = 2 * p.xstride;
    const double lr = p.lr, kd = (double)K;
    const int64_t base = p.step_base;
    unsigned phase = 0;

    // reply of every worker at column j for the step whose tag is `tag` (own from Gp, peers' from the receive area),
    // regularized on its own support and folded in rank order (core/Master.scala:194)
    auto reduce_replies = [&](int j, const double *Gp, const unsigned long long *rcv, unsigned tag, double c, bool add_c,
                              bool &ok) -> double {
      double s = 0.0;
      for (int k = 0; k < K; ++k) {
        double raw;
        if (k == me) {
          raw = __ldcg(&Gp[j]);
        } else {
          unsigned spins = 0;
          const long long t0 = clock64();
          while (!ll_try_load(rcv + 2 * ((size_t)k * snd_stride + j), tag, raw)) {
            if ((++spins & 255u) == 0u && (clock64() - t0 > p.timeout_cycles || *(volatile int *)p.abort_flag)) {
              *(volatile int *)p.abort_flag = 1;
              ok = false;
              raw = 0.0;
              break;
            }
          }
        }
        if (j == p.dim) {
          s += raw;                                           // packed counters: plain sum
        } else {
          double v = filt(raw);
          if (v != 0.0 && add_c) v = filt(v + c);
          s = (k == 0) ? v : filt(s + v);
        }
      }
      return s;
    };
    auto updated = [&](double wv, double s) -> double {      // w - lr * (sum / K), with the constructor filters
      if (s != 0.0) {
        const double mean = filt(s / kd);
        const double step = filt(mean * lr);
        wv = filt(wv - step);
      }
      return wv;
    };

    for (int64_t T = base; T <= base + S; ++T) {
      const int64_t t = T - base;
      const bool first = (T == base), last = (T == base + S);
      const double *Wprev = p.wbuf[(T + 1) & 1];              // W_{T-1}
      double *Wcur = p.wbuf[T & 1];                           // W_T (already materialised when `first`)
      const double *Gprev = p.xg[(T + 2) % 3];                // g_{T-1}
      double *Gcur = p.xg[T % 3];
      double *Gzero = p.xg[(T + 1) % 3];
      const int parp = (int)((T + 1) & 1);                    // receive parity of step T-1
      const unsigned tag = (unsigned)T;                       // words of step T-1 carry tag T
      const unsigned long long *rcv = reinterpret_cast<const unsigned long long *>(p.xrecv[me]) + 2 * (size_t)parp * par_stride;
      const double *part_prev = p.partial + (size_t)((T + 1) & 1) * G * 2;   // partials of W_{T-1}
      double *part_cur = p.partial + (size_t)(T & 1) * G * 2;
      const unsigned c_par = (unsigned)((t >> 1) & 1);
      bool ok = true;

      // ---- push g_{T-1} (every sync warp; one column per thread) ----
      if (!first) {
        for (int j = j_lo + threadIdx.x; j < j_hi; j += kSyncThreads) {
          const double v = __ldcg(&Gprev[j]);
          for (int k = 0; k < K; ++k)
            if (k != me)
              ll_store(reinterpret_cast<unsigned long long *>(p.xrecv[k]) + 2 * ((size_t)me * snd_stride + (size_t)parp * par_stride + j), v, tag);
        }
      }

      if (is_cons) {
        if (!last) {
          const int st = (int)(t % kStages);
          auto &mt = sm.meta[st];
          mbar_wait(&sm.full[st], (unsigned)((t / kStages) & 1), p.abort_flag, p.timeout_cycles);
          const int n_ch = mt.n_chunks;
          const uint2 *ring = &sm.ring[st][0];
          double c_prev = 0.0;
          bool add_c = false, have_c = false;
          auto get_c = [&]() {
            if (!have_c) {
              mbar_wait(&sm.c_bar[t & 1], c_par, p.abort_flag, p.timeout_cycles);
              c_prev = sm.c_val[t & 1];
              add_c = (c_prev != 0.0) && (fabs(c_prev) > kEps);
              have_c = true;
            }
          };
          auto weight_at = [&](unsigned col) -> double {    // W_T[col]
            if (first) return __ldcg(&Wcur[col]);
            const double s = reduce_replies((int)col, Gprev, rcv, tag, c_prev, add_c, ok);
            return updated(__ldcg(&Wprev[col]), s);
          };
          for (int c = warp; c < n_ch; c += kCons) {
            const uint32_t off = mt.ch_off[c];
            const int n = mt.ch_n[c];
            const uint2 *src = (off & kChunkGlobal) ? (p.pairs + (off & ~kChunkGlobal)) : (ring + off);
            // the chunk's gathers are issued in rounds so that their L2 latencies overlap: W_{T-1} and the own reply
            // for all four pairs of a lane first, then one round of four LL loads per peer (rank order = fold order)
            uint2 pr[4];
            double wv[4], sacc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int k = u * 32 + lane;
              pr[u] = (k < n) ? src[k] : make_uint2(0u, 0u);   // col 0 / val 0: inert, still a valid gather
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) wv[u] = __ldcg(&(first ? Wcur : Wprev)[pr[u].x]);
            if (!first) {
              get_c();
              for (int k = 0; k < K; ++k) {
                double raw[4];
                if (k == me) {
#pragma unroll
                  for (int u = 0; u < 4; ++u) raw[u] = __ldcg(&Gprev[pr[u].x]);
                } else {
                  const unsigned long long *rk = rcv + 2 * (size_t)k * snd_stride;
                  bool got[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) got[u] = ll_try_load(rk + 2 * (size_t)pr[u].x, tag, raw[u]);
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    if (!got[u]) {
                      unsigned spins = 0;
                      const long long t0 = clock64();
                      while (!ll_try_load(rk + 2 * (size_t)pr[u].x, tag, raw[u])) {
                        if ((++spins & 255u) == 0u && (clock64() - t0 > p.timeout_cycles || *(volatile int *)p.abort_flag)) {
                          *(volatile int *)p.abort_flag = 1;
                          ok = false;
                          raw[u] = 0.0;
                          break;
                        }
                      }
                    }
                  }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  double v = filt(raw[u]);
                  if (v != 0.0 && add_c) v = filt(v + c_prev);
                  sacc[u] = (k == 0) ? v : filt(sacc[u] + v);
                }
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) wv[u] = updated(wv[u], sacc[u]);
            }
            double acc = 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += filt(filt((double)__uint_as_float(pr[u].y)) * wv[u]);
            acc = warp_sum(acc);
            if (lane == 0) mt.part[c] = acc;
          }
          named_bar_sync(2, kCons * 32);
          unsigned hinge = 0;
          for (int c = warp; c < n_ch; c += kCons) {
            const int row = mt.ch_row[c];
            const int firstc = mt.row_first[row], nch = mt.row_nch[row];
            double dot = 0.0;
            for (int i = 0; i < nch; ++i) dot += mt.part[firstc + i];
            const int yi = mt.row_y[row];
            const double y = (double)yi;
            if (c == firstc && lane == 0) hinge += (unsigned)(1 - yi * ((dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0)));
            if (!(y * dot < 0.0)) {
              const uint32_t off = mt.ch_off[c];
              const int n = mt.ch_n[c];
              const uint2 *src = (off & kChunkGlobal) ? (p.pairs + (off & ~kChunkGlobal)) : (ring + off);
              for (int k = lane; k < n; k += 32) {
                const uint2 pr = src[k];
                const double gvv = filt(filt((double)__uint_as_float(pr.y)) * y);
                if (gvv != 0.0) atomicAdd(&Gcur[pr.x], gvv);
              }
            }
          }
          for (int m = warp; m < mt.n_rows; m += kCons) {
            const int nch = mt.row_nch[m];
            if (nch == 0) {
              if (lane == 0) hinge += 1u;
            } else if (nch < 0) {  // row outside the chunk list: whole row from global memory
              const uint2 *grow = p.pairs + (size_t)mt.row_b[m] * 2;
              const int len = mt.row_len[m];
              if (!first) get_c();
              double acc = 0.0;
              for (int k = lane; k < len; k += 32) {
                const uint2 pr = __ldg(&grow[k]);
                acc += filt(filt((double)__uint_as_float(pr.y)) * weight_at(pr.x));
              }
              const double dot = warp_sum(acc);
              const int yi = mt.row_y[m];
              const double y = (double)yi;
              if (lane == 0) hinge += (unsigned)(1 - yi * ((dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0)));
              if (!(y * dot < 0.0))
                for (int k = lane; k < len; k += 32) {
                  const uint2 pr = __ldg(&grow[k]);
                  const double gvv = filt(filt((double)__uint_as_float(pr.y)) * y);
                  if (gvv != 0.0) atomicAdd(&Gcur[pr.x], gvv);
                }
            }
          }
          if (lane == 0 && hinge) atomicAdd(&sm.hinge_acc, hinge);
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.empty[st]);
        }
      } else {
        // ---- update warps ----
        const int uw = warp - kCons;
        double c_prev = 0.0, nrm_prev = 0.0;                  // of W_{T-1}
        if (uw == 0) {
          if (!first) {
            if (T - 1 == base) {
              c_prev = p.scal[kScalC];                        // W_base came from the host: k_prepare / previous launch
              nrm_prev = p.scal[kScalNrm2];
            } else {
              double sd, sn;
              sum_partials2(part_prev, G, lane, sd, sn);
              c_prev = p.lambda * 2.0 * sd;
              nrm_prev = sn;
            }
          }
          if (lane == 0) {
            sm.c_val[t & 1] = c_prev;
            sm.nrm_val[t & 1] = nrm_prev;                     // ||W_{T-1}||^2 for the loss of step T-1
            mbar_arrive(&sm.c_bar[t & 1]);
          }
          __syncwarp();
        } else {
          mbar_wait(&sm.c_bar[t & 1], c_par, p.abort_flag, p.timeout_cycles);
          c_prev = sm.c_val[t & 1];
        }
        const bool add_c = (c_prev != 0.0) && (fabs(c_prev) > kEps);
        double pd = 0.0, pn = 0.0;
        const int ut = threadIdx.x - kCons * 32;              // 0 .. kUpd*32-1
        for (int j = j_lo + ut; j < j_hi; j += kUpd * 32) {
          if (!first) {
            const double s = reduce_replies(j, Gprev, rcv, tag, c_prev, add_c, ok);
            if (j == p.dim) {
              if (p.losses) {  // loss of step T-1 on W_{T-1}
                const double ns = floor(s / 4294967296.0);
                p.losses[t - 1] = p.lambda * sm.nrm_val[t & 1] + (s - ns * 4294967296.0) / ns;
              }
            } else {
              const double wn = updated(__ldcg(&Wprev[j]), s);
              Wcur[j] = wn;
              pd += filt(wn * __ldg(&p.d[j]));
              pn += wn * wn;
            }
          }
          Gzero[j] = 0.0;
        }
        pd = warp_sum(pd);
        pn = warp_sum(pn);
        if (lane == 0) { sm.red[uw][0] = pd; sm.red[uw][1] = pn; }
        named_bar_sync(1, kUpd * 32);
        if (uw == 0 && lane == 0 && !first) {
          double sd = 0.0, sn = 0.0;
#pragma unroll
          for (int i = 0; i < kUpd; ++i) { sd += sm.red[i][0]; sn += sm.red[i][1]; }
          part_cur[2 * blockIdx.x] = sd;
          part_cur[2 * blockIdx.x + 1] = sn;
        }
      }
      if (!ok) *(volatile int *)&sm.ok = 0;
      // ---- grid barrier T (the CTA's hinge total and batch ride in slot [dim] of g_T) ----
      named_bar_sync(3, kSyncThreads);
      if (*(volatile int *)&sm.ok == 0) { *(volatile int *)p.abort_flag = 1; }
      if (threadIdx.x == 0 && !last) {
        const unsigned h = sm.hinge_acc;
        if (h) { atomicAdd(&Gcur[p.dim], (double)h); sm.hinge_acc = 0u; }
        if (blockIdx.x == 0) atomicAdd(&Gcur[p.dim], (double)B * 4294967296.0);
      }
      ++phase;
      if (!grid_barrier(p.bar, phase * (unsigned)G, p.abort_flag, p.timeout_cycles, &sm.ok, kSyncThreads)) return;
      if (*(volatile int *)p.abort_flag) return;
    }
    // epilogue: W_{base+S} is complete in wbuf[(base+S) & 1]; its c and norm come from the partials just written
    {
      const double *Wfin = p.wbuf[(base + S) & 1];
      const int n_all = G * kSyncThreads;
      for (int j = blockIdx.x * kSyncThreads + threadIdx.x; j < p.dim; j += n_all) {
        const double wv = __ldcg(&Wfin[j]);
        p.w_out[j] = wv;
        p.w32_out[j] = (float)wv;
      }
      if (blockIdx.x == 0 && warp == 0 && S > 0) {
        double sd, sn;
        sum_partials2(p.partial + (size_t)((base + S) & 1) * G * 2, G, lane, sd, sn);
        if (lane == 0) { p.scal[kScalC] = p.lambda * 2.0 * sd; p.scal[kScalNrm2] = sn; }
      }
    }
    return;
  }

  if constexpr (kMode == 1) {
    // =======================================================================================================
    // world > 1, two grid barriers per step (kept for A/B; DSGD_P2P_TWO_BARRIERS=1).  Step T (global number):
    //   [A] consumers: x.W_T from the materialised W_T buffer; gate; RED y*x into this rank's g_T
    //       -- grid barrier 1 --
    //   [push] CTA b owns a contiguous slice of the columns on EVERY rank: it copies its slice of g_T into each
    //       peer's receive area with plain remote stores (one-way NVLink traffic, no round trip), fences at
    //       system scope and raises ITS flag word on each peer.
    //   [B] CTA b waits for the K-1 flags of the peers' CTA b only, then reduces its slice from LOCAL memory in
    //       rank order (the master's left fold over replies, core/Master.scala:194), regularizing each reply on
    //       its own support: W_{T+1} = W_T - lr * sum / K; partials of c_{T+1}, ||W_{T+1}||^2;
    //       zeroes the buffer g_{T+1} will use.
    //       -- grid barrier 2 --
    // Every rank computes the full W_{T+1} itself with the same operations in the same order: replicas stay
    // bit-identical, no broadcast exists.  The cross-GPU critical path is one store + fence + one flag store.
    // =======================================================================================================
    const int K = p.world, me = p.rank;
    // columns per CTA, plus ONE counter slot [dim] = hinge total + 2^32 * sample count (exact in fp64)
    const int slice = (p.dim + 1 + G - 1) / G;
    const int j_lo = min(blockIdx.x * slice, p.dim + 1), j_hi = min(j_lo + slice, p.dim + 1);
    const int n_all = G * kSyncThreads;
    const int a0 = blockIdx.x * kSyncThreads + threadIdx.x;
    const int par_stride = p.xstride, snd_stride = 2 * p.xstride;
    const double lr = p.lr, kd = (double)K;
    unsigned phase = 0;
    double c_cur = p.scal[kScalC], nrm_cur = p.scal[kScalNrm2];  // of W_base, left by the previous launch / k_prepare
    for (int64_t t = 0; t < S; ++t) {
      const int64_t T = p.step_base + t;
      const double *Wt = p.wbuf[T & 1];
      double *Wn = p.wbuf[(T + 1) & 1];
      double *Gme = p.xg[T % 3];
      double *Gzero = p.xg[(T + 1) % 3];
      const int par = (int)(T & 1);
      double *part_cur = p.partial + (size_t)((T + 1) & 1) * G * 2;
      if (warp == 0) DSGD_TL(0);
      if (is_cons) {
        const int st = (int)(t % kStages);
        auto &mt = sm.meta[st];
        mbar_wait(&sm.full[st], (unsigned)((t / kStages) & 1), p.abort_flag, p.timeout_cycles);
        if (warp == 0) DSGD_TL(1);
        const int n_ch = mt.n_chunks;
        const uint2 *ring = &sm.ring[st][0];
        for (int c = warp; c < n_ch; c += kCons) {
          const uint32_t off = mt.ch_off[c];
          const int n = mt.ch_n[c];
          const uint2 *src = (off & kChunkGlobal) ? (p.pairs + (off & ~kChunkGlobal)) : (ring + off);
          double acc = 0.0;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = u * 32 + lane;
            if (k < n) {
              const uint2 pr = src[k];
              acc += filt(filt((double)__uint_as_float(pr.y)) * __ldcg(&Wt[pr.x]));
            }
          }
          acc = warp_sum(acc);
          if (lane == 0) mt.part[c] = acc;
        }
        named_bar_sync(2, kCons * 32);
        unsigned hinge = 0;
        for (int c = warp; c < n_ch; c += kCons) {
          const int row = mt.ch_row[c];
          const int first = mt.row_first[row], nch = mt.row_nch[row];
          double dot = 0.0;
          for (int i = 0; i < nch; ++i) dot += mt.part[first + i];
          const int yi = mt.row_y[row];
          const double y = (double)yi;
          if (c == first && lane == 0) hinge += (unsigned)(1 - yi * ((dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0)));
          if (!(y * dot < 0.0)) {
            const uint32_t off = mt.ch_off[c];
            const int n = mt.ch_n[c];
            const uint2 *src = (off & kChunkGlobal) ? (p.pairs + (off & ~kChunkGlobal)) : (ring + off);
            for (int k = lane; k < n; k += 32) {
              const uint2 pr = src[k];
              const double gvv = filt(filt((double)__uint_as_float(pr.y)) * y);
              if (gvv != 0.0) atomicAdd(&Gme[pr.x], gvv);
            }
          }
        }
        for (int m = warp; m < mt.n_rows; m += kCons) {
          const int nch = mt.row_nch[m];
          if (nch == 0) {
            if (lane == 0) hinge += 1u;
          } else if (nch < 0) {  // row outside the chunk list: whole row from global memory
            const uint2 *grow = p.pairs + (size_t)mt.row_b[m] * 2;
            const int len = mt.row_len[m];
            double acc = 0.0;
            for (int k = lane; k < len; k += 32) {
              const uint2 pr = __ldg(&grow[k]);
              acc += filt(filt((double)__uint_as_float(pr.y)) * __ldcg(&Wt[pr.x]));
            }
            const double dot = warp_sum(acc);
            const int yi = mt.row_y[m];
            const double y = (double)yi;
            if (lane == 0) hinge += (unsigned)(1 - yi * ((dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0)));
            if (!(y * dot < 0.0))
              for (int k = lane; k < len; k += 32) {
                const uint2 pr = __ldg(&grow[k]);
                const double gvv = filt(filt((double)__uint_as_float(pr.y)) * y);
                if (gvv != 0.0) atomicAdd(&Gme[pr.x], gvv);
              }
          }
        }
        if (lane == 0 && hinge) atomicAdd(&sm.hinge_acc, hinge);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[st]);
        if (warp == 0) DSGD_TL(2);
      }
      // ---- grid barrier 1: this rank's g_T is complete (the CTA's hinge total and batch ride in slots dim, dim+1) ----
      named_bar_sync(3, kSyncThreads);
      if (threadIdx.x == 0) {
        const unsigned h = sm.hinge_acc;
        if (h) { atomicAdd(&Gme[p.dim], (double)h); sm.hinge_acc = 0u; }
        if (blockIdx.x == 0) atomicAdd(&Gme[p.dim], (double)B * 4294967296.0);
      }
      ++phase;
      if (!grid_barrier(p.bar, phase * (unsigned)G, p.abort_flag, p.timeout_cycles, &sm.ok, kSyncThreads)) return;
      if (warp == 0) DSGD_TL(3);
      // ---- push: my slice of g_T (and the packed counter slot) into every peer's receive area as LL words ----
      const unsigned tag = (unsigned)(T + 1);
      for (int j = j_lo + threadIdx.x; j < j_hi; j += kSyncThreads) {
        const double v = __ldcg(&Gme[j]);
        for (int k = 0; k < K; ++k)
          if (k != me)
            ll_store(reinterpret_cast<unsigned long long *>(p.xrecv[k]) + 2 * ((size_t)me * snd_stride + (size_t)par * par_stride + j), v, tag);
      }
      if (warp == 0) DSGD_TL(4);
      // ---- [B]: every thread waits for ITS elements from every peer (tag == step), reduces in rank order, updates ----
      const bool add_c = (c_cur != 0.0) && (fabs(c_cur) > kEps);
      const unsigned long long *rcv = reinterpret_cast<const unsigned long long *>(p.xrecv[me]) + 2 * (size_t)par * par_stride;
      double pd = 0.0, pn = 0.0;
      bool ok = true;
      const long long t0 = clock64();
      for (int j = j_lo + threadIdx.x; j < j_hi; j += kSyncThreads) {
        double s = 0.0;
        for (int k = 0; k < K; ++k) {
          double raw;
          if (k == me) {
            raw = __ldcg(&Gme[j]);
          } else {
            unsigned spins = 0;
            while (!ll_try_load(rcv + 2 * ((size_t)k * snd_stride + j), tag, raw)) {
              if ((++spins & 255u) == 0u && (clock64() - t0 > p.timeout_cycles || *(volatile int *)p.abort_flag)) {
                *(volatile int *)p.abort_flag = 1;
                ok = false;
                raw = 0.0;
                break;
              }
            }
          }
          if (j == p.dim) {                                   // packed counters: plain sum
            s += raw;
          } else {
            double v = filt(raw);
            if (v != 0.0 && add_c) v = filt(v + c_cur);       // regularize on this worker's own support
            s = (k == 0) ? v : filt(s + v);                     // Vec.sum: left fold over the replies
          }
        }
        if (j == p.dim) {
          if (p.losses) {  // loss of step T on W_T: lambda*||W_T||^2 + (all ranks' hinge) / (all ranks' samples)
            const double ns = floor(s / 4294967296.0);
            p.losses[t] = p.lambda * nrm_cur + (s - ns * 4294967296.0) / ns;
          }
        } else {
          double wn = __ldcg(&Wt[j]);
          if (s != 0.0) {
            const double mean = filt(s / kd);
            const double step = filt(mean * lr);
            wn = filt(wn - step);
          }
          Wn[j] = wn;
          pd += filt(wn * __ldg(&p.d[j]));
          pn += wn * wn;
        }
        Gzero[j] = 0.0;
      }
      if (warp == 0) DSGD_TL(5);
      if (!ok) *(volatile int *)&sm.ok = 0;
      if (warp == 0) DSGD_TL(8);
      // per-CTA partials of c_{T+1} and ||W_{T+1}||^2 in a fixed order: warp -> shared -> thread 0
      pd = warp_sum(pd);
      pn = warp_sum(pn);
      __shared__ double red_all[kCons + kUpd][2];
      if (lane == 0) { red_all[warp][0] = pd; red_all[warp][1] = pn; }
      named_bar_sync(3, kSyncThreads);
      if (*(volatile int *)&sm.ok == 0) return;
      if (threadIdx.x == 0) {
        double sd = 0.0, sn = 0.0;
#pragma unroll
        for (int i = 0; i < kCons + kUpd; ++i) { sd += red_all[i][0]; sn += red_all[i][1]; }
        part_cur[2 * blockIdx.x] = sd;
        part_cur[2 * blockIdx.x + 1] = sn;
      }
      // ---- grid barrier 2: W_{T+1} and its partials are complete ----
      if (warp == 0) DSGD_TL(9);
      ++phase;
      if (!grid_barrier(p.bar, phase * (unsigned)G, p.abort_flag, p.timeout_cycles, &sm.ok, kSyncThreads)) return;
      if (warp == 0) DSGD_TL(10);
      if (warp == 0) {
        double sd, sn;
        sum_partials2(part_cur, G, lane, sd, sn);
        if (lane == 0) { sm.c_val[0] = p.lambda * 2.0 * sd; sm.c_val[1] = sn; }
      }
      named_bar_sync(3, kSyncThreads);
      c_cur = sm.c_val[0];
      nrm_cur = sm.c_val[1];
    }
    // epilogue: publish W_{base+S} as the resident weights
    const double *Wfin = p.wbuf[(p.step_base + S) & 1];
    for (int j = a0; j < p.dim; j += n_all) {
      const double wv = __ldcg(&Wfin[j]);
      p.w_out[j] = wv;
      p.w32_out[j] = (float)wv;
    }
    if (a0 == 0) { p.scal[kScalC] = c_cur; p.scal[kScalNrm2] = nrm_cur; }
    return;
  }

  const int n_upd = G * kUpd * 32;
  const int u0 = blockIdx.x * kUpd * 32 + (threadIdx.x - kCons * 32);
  const double k_den = p.k_den, lr = p.lr;
  unsigned phase = 0;

  for (int64_t t = 0; t <= S; ++t) {
    const double *Wprev = p.wbuf[(t + 1) & 1];
    double *Wcur = p.wbuf[t & 1];
    const double *Gprev = p.gbuf[(t + 2) % 3];
    double *Gcur = p.gbuf[t % 3];
    double *Gzero = p.gbuf[(t + 1) % 3];
    const double *part_prev = (kOpt & 2) ? p.push + ((size_t)((t + 1) & 1) * G + blockIdx.x) * G * 2   // this CTA's private copy
                                         : p.partial + (size_t)((t + 1) & 1) * G * 2;
    double *part_cur = (kOpt & 2) ? p.push + (size_t)(t & 1) * G * G * 2 : p.partial + (size_t)(t & 1) * G * 2;
    const unsigned c_par = (unsigned)((t >> 1) & 1);

    if (is_cons) {
      if (warp == 0) DSGD_TL(0);
      if (t < S) {
        const int st = (int)(t % kStages);
        auto &mt = sm.meta[st];
        mbar_wait(&sm.full[st], (unsigned)((t / kStages) & 1), p.abort_flag, p.timeout_cycles);
        if (warp == 0) DSGD_TL(1);
        const int n_ch = mt.n_chunks;
        const uint2 *ring = &sm.ring[st][0];
        double c_prev = 0.0;
        bool add_c = false, have_c = false;
        auto get_c = [&]() {  // c_{t-1}: summed by update warp 0 of this CTA, handed over through shared memory
          if (!have_c) {
            mbar_wait(&sm.c_bar[t & 1], c_par, p.abort_flag, p.timeout_cycles);
            c_prev = sm.c_val[t & 1];
            add_c = (c_prev != 0.0) && (fabs(c_prev) > kEps);
            have_c = true;
          }
        };
        // ---- pass 1: partial dots of this warp's chunks ----
        [[maybe_unused]] unsigned hinge_early = 0;
        for (int c = warp; c < n_ch; c += kCons) {
          const uint32_t off = mt.ch_off[c];
          const int n = mt.ch_n[c];
          const uint2 *src = (off & kChunkGlobal) ? (p.pairs + (off & ~kChunkGlobal)) : (ring + off);
          uint2 pr[4];
          double wv[4], gv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = u * 32 + lane;
            pr[u] = (k < n) ? src[k] : make_uint2(0u, 0u);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = u * 32 + lane;
            wv[u] = 0.0;
            gv[u] = 0.0;
            if (k < n) {
              wv[u] = __ldcg(&Wprev[pr[u].x]);
              gv[u] = __ldcg(&Gprev[pr[u].x]);
            }
          }
          get_c();
          double acc = 0.0;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double xv = filt((double)__uint_as_float(pr[u].y));
            const double wt = apply_update(wv[u], gv[u], c_prev, add_c, k_den, lr);
            acc += filt(xv * wt);  // (x * w).sum  (math/Vec.scala:58)
          }
          acc = warp_sum(acc);
          if constexpr (kOpt & 4) {
            // a row that is ONE chunk (<= 128 pairs: ~85 % of the rows) is complete in this warp: gate and scatter from
            // the registers that still hold its pairs, no partial, no second pass (same dot: 0.0 + acc == acc)
            const int row1 = mt.ch_row[c];
            if (mt.row_nch[row1] == 1) {
              const int yi = mt.row_y[row1];
              const double y = (double)yi;
              if (lane == 0) {
                const int pred = (acc > 0.0) ? -1 : ((acc < 0.0) ? 1 : 0);
                hinge_early += (unsigned)(1 - yi * pred);
              }
              if (!(y * acc < 0.0)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const double gvv = filt(filt((double)__uint_as_float(pr[u].y)) * y);   // lanes past the chunk hold val 0
                  if (gvv != 0.0) atomicAdd(&Gcur[pr[u].x], gvv);
                }
              }
              continue;
            }
          }
          if (lane == 0) mt.part[c] = acc;
        }
        if (warp == 0) DSGD_TL(2);
        named_bar_sync(2, kCons * 32);
        // ---- pass 2: row dot (chunk partials in order), prediction, gate, scatter ----
        unsigned hinge = 0;
        if constexpr (kOpt & 4) hinge = hinge_early;
        for (int c = warp; c < n_ch; c += kCons) {
          const int row = mt.ch_row[c];
          const int first = mt.row_first[row], nch = mt.row_nch[row];
          if constexpr (kOpt & 4) {
            if (nch == 1) continue;   // done in pass 1
          }
          double dot = 0.0;
          for (int i = 0; i < nch; ++i) dot += mt.part[first + i];
          const int yi = mt.row_y[row];
          const double y = (double)yi;
          if (c == first && lane == 0) {
            const int pred = (dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0);
            hinge += (unsigned)(1 - yi * pred);
          }
          if (!(y * dot < 0.0)) {  // SparseSVM.scala:28
            const uint32_t off = mt.ch_off[c];
            const int n = mt.ch_n[c];
            const uint2 *src = (off & kChunkGlobal) ? (p.pairs + (off & ~kChunkGlobal)) : (ring + off);
            for (int k = lane; k < n; k += 32) {
              const uint2 pr = src[k];
              const double gvv = filt(filt((double)__uint_as_float(pr.y)) * y);
              if (gvv != 0.0) atomicAdd(&Gcur[pr.x], gvv);
            }
          }
        }
        // rows outside the chunk list: empty rows (dot 0 -> prediction 0, hinge 1, nothing to scatter) and, if a
        // step ever overflows the chunk list, whole rows straight from global memory, one warp per row
        for (int m = warp; m < mt.n_rows; m += kCons) {
          const int nch = mt.row_nch[m];
          if (nch == 0) {
            if (lane == 0) hinge += 1u;
          } else if (nch < 0) {
            const uint2 *grow = p.pairs + (size_t)mt.row_b[m] * 2;
            const int len = mt.row_len[m];
            get_c();
            double acc = 0.0;
            for (int k = lane; k < len; k += 32) {
              const uint2 pr = __ldg(&grow[k]);
              const double wt = apply_update(__ldcg(&Wprev[pr.x]), __ldcg(&Gprev[pr.x]), c_prev, add_c, k_den, lr);
              acc += filt(filt((double)__uint_as_float(pr.y)) * wt);
            }
            const double dot = warp_sum(acc);
            const int yi = mt.row_y[m];
            const double y = (double)yi;
            if (lane == 0) {
              const int pred = (dot > 0.0) ? -1 : ((dot < 0.0) ? 1 : 0);
              hinge += (unsigned)(1 - yi * pred);
            }
            if (!(y * dot < 0.0)) {
              for (int k = lane; k < len; k += 32) {
                const uint2 pr = __ldg(&grow[k]);
                const double gvv = filt(filt((double)__uint_as_float(pr.y)) * y);
                if (gvv != 0.0) atomicAdd(&Gcur[pr.x], gvv);
              }
            }
          }
        }
        if (lane == 0 && hinge) atomicAdd(&sm.hinge_acc, hinge);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[st]);
        if (warp == 0) DSGD_TL(3);
      }
    } else {
      const int uw = warp - kCons;
      if (uw == 0) DSGD_TL(8);
      // c_{t-1} = 2*lambda*(W_{t-1} . d): at t == 0 g_{-1} is all zero, so its value is irrelevant
      double c_prev = 0.0;
      if (uw == 0) {
        double sn = 0.0;
        if (t > 0) {
          double sd;
          sum_partials2(part_prev, G, lane, sd, sn);
          c_prev = p.lambda * 2.0 * sd;
        }
        if (lane == 0) {  // hand c over first: the consumers of this CTA are waiting for it
          sm.c_val[t & 1] = c_prev;
          mbar_arrive(&sm.c_bar[t & 1]);
        }
        DSGD_TL(9);
        // loss of step t-1 = lambda*||W_{t-1}||^2 + hinge_{t-1}/batch  (SparseSVM.scala:20-23)
        if (t > 0 && p.losses && blockIdx.x == 0 && lane == 0)
          p.losses[t - 1] = p.lambda * sn + (double)__ldcg(&p.hinge[t - 1]) / (double)B;
      } else {
        mbar_wait(&sm.c_bar[t & 1], c_par, p.abort_flag, p.timeout_cycles);
        c_prev = sm.c_val[t & 1];
      }
      const bool add_c = (c_prev != 0.0) && (fabs(c_prev) > kEps);
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
      if (lane == 0) { sm.red[uw][0] = pd; sm.red[uw][1] = pn; }
      named_bar_sync(1, kUpd * 32);
      if constexpr (kOpt & 2) {
        // every update thread sums the CTA's kUpd pairs (same order) and stores them into the areas of the CTAs it serves
        double sd = 0.0, sn = 0.0;
#pragma unroll
        for (int i = 0; i < kUpd; ++i) { sd += sm.red[i][0]; sn += sm.red[i][1]; }
        for (int dest = threadIdx.x - kCons * 32; dest < G; dest += kUpd * 32)
          *reinterpret_cast<double2 *>(part_cur + ((size_t)dest * G + blockIdx.x) * 2) = make_double2(sd, sn);
      } else if (uw == 0 && lane == 0) {
        double sd = 0.0, sn = 0.0;
#pragma unroll
        for (int i = 0; i < kUpd; ++i) { sd += sm.red[i][0]; sn += sm.red[i][1]; }
        part_cur[2 * blockIdx.x] = sd;
        part_cur[2 * blockIdx.x + 1] = sn;
      }
      if (uw == 0) DSGD_TL(10);
    }
    // ---- grid barrier t (the CTA's hinge total rides in front of the arrival) ----
    named_bar_sync(3, kSyncThreads);
    if (threadIdx.x == 0) {
      const unsigned h = sm.hinge_acc;
      if (h) {
        atomicAdd(&p.hinge[t < S ? t : 0], h);
        sm.hinge_acc = 0u;
      }
    }
    ++phase;
    long long *tl_slot = nullptr;
    bool tl_ns = false;
    if (p.tl) {
      if (t >= 100 && t < 104) { tl_slot = p.tl + 4096 + ((t - 100) * 160 + blockIdx.x) * 2; tl_ns = true; }
      else if (blockIdx.x == 0 && t < 256) tl_slot = p.tl + t * 16 + 6;
    }
    if constexpr (kOpt & 1) {
      if (!grid_barrier_flags(p.bar, p.bar_flags, phase, (unsigned)G, p.abort_flag, p.timeout_cycles, &sm.ok, kSyncThreads,
                              tl_slot, tl_ns)) return;
    } else {
      if (!grid_barrier(p.bar, phase * (unsigned)G, p.abort_flag, p.timeout_cycles, &sm.ok, kSyncThreads, tl_slot, tl_ns)) return;
    }
  }

  // ---- epilogue: W_S is complete in wbuf[S & 1]; publish it as the resident weights, clear g_{S-1} ----------
  if (is_upd) {
    const double *Wfin = p.wbuf[S & 1];
    double *Glast = p.gbuf[(S + 2) % 3];
    for (int j = u0; j < p.dim; j += n_upd) {
      const double wv = __ldcg(&Wfin[j]);
      p.w_out[j] = wv;
      p.w32_out[j] = (float)wv;
      Glast[j] = 0.0;
    }
    if (blockIdx.x == 0 && warp == kCons) {
      const double *part = (kOpt & 2) ? p.push + (size_t)(S & 1) * G * G * 2   // CTA 0's private copy
                                      : p.partial + (size_t)(S & 1) * G * 2;
      double sd, sn;
      sum_partials2(part, G, lane, sd, sn);
      if (lane == 0) {
        p.scal[kScalC] = p.lambda * 2.0 * sd;
        p.scal[kScalNrm2] = sn;
      }
    }
  }
}

}  // namespace dsgd
