// Device side of K5 (block triangular solve on the Cholesky factor): the per-warp tape
// ring and the tile-task loop.  Shared by the stand-alone preconditioner kernel (sptrsv.cu)
// and the per-agent RTR cluster kernel (rtr_cluster.cu).  Tape format: trsv.h.
#pragma once
#include <cooperative_groups.h>

#include "kernels.cuh"
#include "trsv.h"

namespace dcora {
namespace {

constexpr int RING_SLOTS = 8;   // records per warp ring (one TMA / one mbarrier each)
constexpr int L2_AHEAD = 16;    // records of L2 prefetch distance per warp (~18 KB)

template <int DH>
struct TapeRec {
  static constexpr int BS = DH * DH;
  static constexpr int W = TAPE_REC_HEADER + 8 * BS;  // words per record
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int CL>
__device__ __forceinline__ void phase_sync() {
  if constexpr (CL == 1) __syncthreads();
  else cooperative_groups::this_cluster().sync();
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(pred));
  return pred != 0;
}

// Values every lane holds identically, re-materialised from lane 0: this is what lets the
// compiler keep them in uniform registers, so the bulk-copy instructions (which take
// uniform operands) need no per-lane "waterfall" loop.
__device__ __forceinline__ int uniform32(int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ unsigned long long uniform64(unsigned long long v) {
  const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)(v & 0xffffffffull), 0);
  const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), 0);
  return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ int lo32(unsigned long long u) { return (int)(unsigned)(u & 0xffffffffull); }
__device__ __forceinline__ int hi32(unsigned long long u) { return (int)(unsigned)(u >> 32); }

// Loads through 32-bit shared-space addresses (the ring's slots are addressed from a
// uniform base + lane offset; generic pointers would make the compiler rebuild the shared
// window address from special registers inside the loop).
__device__ __forceinline__ unsigned long long lds64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// Per-warp ring of RING_SLOTS tape records in shared memory, filled by bulk asynchronous
// copies (TMA) that complete on one mbarrier per slot -- no LSU work at all for the tape,
// which matters: with 16 warps per SM the kernel is bound by the LSU's wavefronts (panel
// gathers + shared-memory reads), see the notes at the end of this file.  The tape is static data: record i
// is requested as soon as record i - RING_SLOTS has been consumed, regardless of barriers,
// and record i + L2_AHEAD is pulled into L2 at the same time, so that the ring's own
// requests are L2 hits (8 records in flight alone would cap a warp's stream at
// ring bytes / HBM latency).
template <int RECW>
struct TapeRing {
  const unsigned long long *src;  // this warp's tape (global), warp-uniform
  unsigned long long *ring;       // RING_SLOTS * RECW words (shared)
  uint32_t ring_s, mbar_s;        // shared-space addresses
  int nrec;                       // records on the tape
  int issued, ready;              // record counters (warp-uniform)
  unsigned base;                  // records that went through these mbarriers before
  int dbg = 0;                    // experiment switches (DCORA_K5_DBG), 0 in production
  unsigned long long l2_first;    // L2 policy of the tape stream: evict first (read once)
  long long *itrace = nullptr;    // per-iteration clock64 stamps of one warp (debug aid)
  int itrace_left = 0;

  __device__ __forceinline__ void setup(unsigned long long *ring_, unsigned long long *mbar_) {
    ring = ring_;
    ring_s = smem_u32(ring_);
    mbar_s = smem_u32(mbar_);
    base = 0;
    nrec = issued = ready = 0;
    // the tape is read exactly once per solve: keep it from evicting the T / X rows the
    // gathers re-read from L2 (ncu: 10.3 GB of DRAM traffic for 6.3 GB of algorithmic bytes
    // at full batch without the hint)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(l2_first));
    if (elect_one()) {
      for (int s = 0; s < RING_SLOTS; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar_s + s * 8) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  __device__ __forceinline__ void issue(int i) {
    const unsigned slot = (base + (unsigned)i) & (RING_SLOTS - 1);
    const uint32_t mb = mbar_s + slot * 8;
    const uint32_t dst = ring_s + slot * (RECW * 8);
    const unsigned long long *g = src + (size_t)i * RECW;
    if (elect_one()) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                   :: "r"(mb), "r"(RECW * 8) : "memory");
      if (dbg & 16)
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            :: "r"(dst), "l"(g), "r"(RECW * 8), "r"(mb) : "memory");
      else
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
            "[%0], [%1], %2, [%3], %4;"
            :: "r"(dst), "l"(g), "r"(RECW * 8), "r"(mb), "l"(l2_first) : "memory");
      if (i + L2_AHEAD < nrec && (dbg & 4))
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                     :: "l"(g + (size_t)L2_AHEAD * RECW), "r"(RECW * 8) : "memory");
    }
  }
  // start streaming a tape (the previous one must have been drained)
  __device__ __forceinline__ void start(const unsigned long long *tape, int nrec_) {
    base += (unsigned)nrec;
    src = tape;
    nrec = nrec_;
    issued = ready = 0;
    const int npf = nrec < RING_SLOTS + L2_AHEAD ? nrec : RING_SLOTS + L2_AHEAD;
    if (elect_one() && (dbg & 4))
      for (int c = RING_SLOTS; c < npf; ++c)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                     :: "l"(src + (size_t)c * RECW), "r"(RECW * 8) : "memory");
    const int depth = (dbg & 2) ? 4 : RING_SLOTS;
    const int n0 = nrec < depth ? nrec : depth;
    for (; issued < n0; ++issued) issue(issued);
  }
  // make record i readable (calls are monotone in i)
  __device__ __forceinline__ void wait(int i) {
    while (ready <= i) {
      const unsigned gi = base + (unsigned)ready;
      const uint32_t mb = mbar_s + (gi & (RING_SLOTS - 1)) * 8;
      const uint32_t parity = (gi / RING_SLOTS) & 1;
      uint32_t ok = 0;
      do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
      } while (!ok);
      ++ready;
    }
  }
  // Non-blocking probe of record i's barrier (i == ready): the result is consumed an
  // iteration later by confirm(), so the barrier's round trip overlaps with arithmetic.
  __device__ __forceinline__ uint32_t probe(int i) const {
    const unsigned gi = base + (unsigned)i;
    const uint32_t mb = mbar_s + (gi & (RING_SLOTS - 1)) * 8;
    const uint32_t parity = (gi / RING_SLOTS) & 1;
    uint32_t ok = 0;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
    return ok;
  }
  __device__ __forceinline__ void confirm(int i, uint32_t ok) {
    if (ready <= i) {
      if (ok && ready == i) ++ready;
      else wait(i);
    }
  }
  // shared-space address of record i
  __device__ __forceinline__ uint32_t rec_s(int i) const {
    return ring_s + ((base + (unsigned)i) & (RING_SLOTS - 1)) * (RECW * 8);
  }
  // record i is consumed: its slot takes record i + RING_SLOTS
  __device__ __forceinline__ void refill() {
    __syncwarp();
    if (issued < nrec) {
      issue(issued);
      ++issued;
    }
  }
  // wait for every copy still in flight (before the ring is reused or the CTA exits)
  __device__ __forceinline__ void drain() { if (issued > 0) wait(issued - 1); }
  __device__ __forceinline__ const unsigned long long *rec(int i) const {
    return ring + ((base + (unsigned)i) & (RING_SLOTS - 1)) * RECW;
  }
};

// per-iteration cycle stamps of one warp (tools/trsv_trace.py with DCORA_K5_DBG=8): compiled
// in only with -DDCORA_K5_TRACE_ITERATIONS (the four tests per record cost ~10 %)
#ifdef DCORA_K5_TRACE_ITERATIONS
#define DCORA_K5_STAMP() \
  do { if (rd.itrace && rd.itrace_left > 0) *rd.itrace++ = clock64(); } while (0)
#define DCORA_K5_STAMP_END() \
  do { if (rd.itrace && rd.itrace_left > 0) { *rd.itrace++ = clock64(); --rd.itrace_left; } } while (0)
#else
#define DCORA_K5_STAMP() do {} while (0)
#define DCORA_K5_STAMP_END() do {} while (0)
#endif

__device__ __forceinline__ unsigned lds32(uint32_t addr) {
  unsigned v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// Executes every round of one warp's tape (format: trsv.h).  Lane (g, a) of the warp --
// group g = lane / 4, a = lane % 4 -- owns COLUMN a of its group's tile: per record it
// needs one row of the source panel (R doubles) and DH tape words, and accumulates all DH
// rows of the product; the 4 lanes of a group (and the `gpt` groups of a task) are summed
// once per round.  A cluster barrier follows the last round of every slot.
template <int R, int DH, int CL>
__device__ __forceinline__ void trsv_run(TapeRing<TapeRec<DH>::W> &rd, double *T, double *X,
                                         int row0, unsigned long long *&trace) {
  static_assert(TAPE_AHEAD == 4, "the row window below holds TAPE_AHEAD rows");
  // The source rows are gathered from L2 PFD records ahead of their use (register ring,
  // statically indexed by unrolling).  Which rows: every record names the rows of the
  // record TAPE_AHEAD further on, so the window (w0..w3) always holds the rows of the
  // current and the next three records and no future record has to have landed for a gather
  // to be issued.
  constexpr int PFD = R <= 8 ? 3 : 1;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, a = lane & 3;
  uint32_t cw_off = (uint32_t)g * 4;
  uint32_t m_off = DH == 4 ? (uint32_t)(TAPE_REC_HEADER + lane * 2) * 8
                           : (uint32_t)(TAPE_REC_HEADER + (g * DH + a) * DH) * 8;
  // keep the lane constants in registers (the compiler would otherwise rebuild them from
  // the thread index -- a slow special-register read -- inside the loop)
  asm volatile("" : "+r"(cw_off), "+r"(m_off));
  const int nrec = rd.nrec;
  int w0 = row0 + a, w1 = w0, w2 = w0, w3 = w0;
  int rec = 0;
  while (rec < nrec) {
    rd.wait(rec);
    const uint32_t hs = rd.rec_s(rec);
    const unsigned long long ctl = lds64(hs + 4 * 8);
    const int dpk = (int)lds32(hs + 5 * 8 + cw_off);
    const int niter = lo32(ctl), cw = hi32(ctl);
    const int gpt = cw & 0xff, type = (cw >> 8) & 0xff;
    const bool last_of_slot = (cw >> 16) & 1;
    const int r0 = dpk >> 3, nrows = dpk & 7;
    const double *src = type == TAPE_BWD ? X : T;
    double *dst = type == TAPE_FWD ? T : X;
    const bool accumulate = type != TAPE_MID;
    const bool writer = dpk >= 0 && a < nrows && (g & (gpt - 1)) == 0;
    const size_t o = (size_t)(r0 + a) * R;
    // the destination's current value (accumulating tasks): independent of everything
    // below, so its latency hides behind the loop
    double old[R];
#pragma unroll
    for (int q = 0; q < R; ++q) old[q] = 0.0;
    if (accumulate && writer) {
#pragma unroll
      for (int q = 0; q < R; ++q) old[q] = dst[o + q];
    }
    double acc[4][R];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < R; ++q) acc[i][q] = 0.0;
    double yq[PFD][R];
#pragma unroll
    for (int s2 = 0; s2 < PFD; ++s2)
#pragma unroll
      for (int q = 0; q < R; ++q) yq[s2][q] = 0.0;
    auto gather = [&](int row, double (&dstv)[R]) {
      if (a < DH && !(rd.dbg & 1)) {
        const double *yj = src + (size_t)row * R;
#pragma unroll
        for (int q = 0; q < R; ++q) dstv[q] = yj[q];
      }
    };
    const int last = rec + niter - 1;
    if (niter > 0) gather(w0, yq[0]);
    if constexpr (PFD > 1) { if (niter > 1) gather(w1, yq[1]); }
    if constexpr (PFD > 2) { if (niter > 2) gather(w2, yq[2]); }
    // the barrier of the next record is probed one iteration early
    uint32_t ok = rec < last ? rd.probe(rec + 1) : 0u;
    for (int it0 = 0; it0 < niter; it0 += PFD) {
#pragma unroll
      for (int u = 0; u < PFD; ++u) {
        const int ri = rec + it0 + u;
        if (ri <= last) {
          DCORA_K5_STAMP();
          if (ri > rec) rd.confirm(ri, ok);
          DCORA_K5_STAMP();
          const uint32_t rs = rd.rec_s(ri);
          const int w4 = (int)lds32(rs + cw_off) + a;   // row of record ri + TAPE_AHEAD
          const uint32_t ms = rs + m_off;
          double m[DH];
          if (a < DH) {
            if constexpr (DH == 4) {
              const double2 m01 = lds_f64x2(ms), m23 = lds_f64x2(ms + 512);
              m[0] = m01.x; m[1] = m01.y; m[2] = m23.x; m[3] = m23.y;
            } else {
#pragma unroll
              for (int i = 0; i < DH; ++i) m[i] = lds_f64(ms + 8 * i);
            }
          } else {
#pragma unroll
            for (int i = 0; i < DH; ++i) m[i] = 0.0;
          }
          if (ri < last) ok = rd.probe(ri + 1);
#pragma unroll
          for (int i = 0; i < DH; ++i)
#pragma unroll
            for (int q = 0; q < R; ++q) acc[i][q] = fma(m[i], yq[u][q], acc[i][q]);
          DCORA_K5_STAMP();
          if (ri + PFD <= last) gather(PFD == 3 ? w3 : PFD == 2 ? w2 : w1, yq[u]);
          w0 = w1; w1 = w2; w2 = w3; w3 = w4;
          rd.refill();
          DCORA_K5_STAMP_END();
        }
      }
    }
    rec += niter;
    // sum over the 4 lanes of the group, leaving row a in lane a
    double row[R];
    {
      double v[2][R];
      const bool hi2 = (a & 2) != 0, hi1 = (a & 1) != 0;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int q = 0; q < R; ++q) {
          const double mine = hi2 ? acc[2 + hh][q] : acc[hh][q];
          const double send = hi2 ? acc[hh][q] : acc[2 + hh][q];
          v[hh][q] = mine + __shfl_xor_sync(0xffffffffu, send, 2);
        }
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const double mine = hi1 ? v[1][q] : v[0][q];
        const double send = hi1 ? v[0][q] : v[1][q];
        row[q] = mine + __shfl_xor_sync(0xffffffffu, send, 1);
      }
    }
    for (int og = 4; og < 4 * gpt; og <<= 1) {
#pragma unroll
      for (int q = 0; q < R; ++q) row[q] += __shfl_xor_sync(0xffffffffu, row[q], og);
    }
    if (writer) {
      if (accumulate) {
#pragma unroll
        for (int q = 0; q < R; ++q) dst[o + q] = old[q] - row[q];
      } else {
#pragma unroll
        for (int q = 0; q < R; ++q) dst[o + q] = row[q];
      }
    }
    if (last_of_slot) {
      phase_sync<CL>();
      if (trace) *trace++ = globaltimer_ns();
    }
  }
}

// T[f] = V[column of factor row f] for the agent's factor rows [f0, f1), dealt to `nt`
// threads; the reverse with TO_COLS (OUT[column] = X[f]).  Row-wise: one index load and R
// independent element moves per row.
template <int R, bool TO_COLS>
__device__ __forceinline__ void trsv_permute_rows(long long f0, long long f1, int tid, int nt,
                                                  const int *__restrict__ frow_col,
                                                  const double *in, double *out) {
  for (long long f = f0 + tid; f < f1; f += nt) {
    const size_t c = (size_t)frow_col[f] * R, ff = (size_t)f * R;
    double v[R];
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = in[(TO_COLS ? ff : c) + q];
#pragma unroll
    for (int q = 0; q < R; ++q) out[(TO_COLS ? c : ff) + q] = v[q];
  }
}

// ---- What bounds this kernel (B200, config-5 agent: 16810 rows, 2 x 1.4 M tile values, 25
// barrier slots, 64 warps = 4 CTAs x 16 warps; per-iteration clock64 stamps, DCORA_K5_DBG=8;
// ncu source view; SASS control words) -----------------------------------------------------
//  * ~1000 cycles per record and warp in every variant tried, i.e. ~65 cycles of some
//    SM-wide resource per warp-iteration: the LSU.  A record costs 5 scattered 8-byte panel
//    gathers (8 lane groups x 160 bytes: 8-16 wavefronts each), the shared-memory reads of
//    the tile column, and -- when the tape or the gathers are staged through shared memory
//    with cp.async -- their write side too.
//  * register prefetch of the panel rows PFD records ahead is defeated by the compiler: all
//    in-flight gathers share ONE scoreboard (decoded control words: every gather `wr 5`, the
//    first DFMA of an iteration waits on SB5), so it waits for the newest gather.
//  * variants measured (one agent alone / its colour class of 28, microseconds per solve):
//      TMA ring, records name their own rows (look-ahead wait on future records)  243 / 281
//      TMA ring, rows named TAPE_AHEAD records ahead (this file)                   256 / 282
//      cp.async ring for the tape, register gathers                               286 / 307
//      cp.async for tape and panel rows (commit groups, no scoreboard aliasing)   388 / 393
//      ... the same with sector-coalesced gathers + conflict-free tile layout     345 / 353
//    i.e. every byte moved through the LSU a second time costs more than the stall it removes.
//  * switching the gathers off entirely (DCORA_K5_DBG=1) gives 207 / 236; the bulk L2 prefetch
//    changes nothing (the ring's own requests already hide HBM).

}  // namespace
}  // namespace dcora
