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
// copies (TMA) that complete on one mbarrier per slot.  The tape is static data: record i
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

  __device__ __forceinline__ void setup(unsigned long long *ring_, unsigned long long *mbar_) {
    ring = ring_;
    ring_s = smem_u32(ring_);
    mbar_s = smem_u32(mbar_);
    base = 0;
    nrec = issued = ready = 0;
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
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          :: "r"(dst), "l"(g), "r"(RECW * 8), "r"(mb) : "memory");
      if (i + L2_AHEAD < nrec && !(dbg & 4))
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
    if (elect_one() && !(dbg & 4))
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

// Executes every round of one warp's tape (format: trsv.h).  Lane (g, a) of the warp --
// group g = lane / 4, a = lane % 4 -- owns COLUMN a of its group's tile: per iteration it
// needs one row of the source panel (R doubles) and DH tape words, and accumulates all DH
// rows of the product; the 4 lanes of a group (and the `gpt` groups of a task) are summed
// once per round.  A cluster barrier follows the last round of every slot.
template <int R, int DH, int CL>
__device__ __forceinline__ void trsv_run(TapeRing<TapeRec<DH>::W> &rd, double *T, double *X,
                                         unsigned long long *&trace) {
  // The source rows are gathered from L2 PFD iterations ahead of their use (register ring,
  // statically indexed by unrolling): 3 deep for r <= 8, 1 deep for higher ranks.
  constexpr int PFD = R <= 8 ? 3 : 1;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, a = lane & 3;
  uint32_t cw_off = (uint32_t)(g >> 1) * 8;
  uint32_t c_shift = (g & 1) ? 32u : 0u;
  uint32_t m_off = (uint32_t)(TAPE_REC_HEADER + (g * DH + a) * DH) * 8;
  // keep the lane constants in registers (the compiler would otherwise rebuild them from
  // the thread index -- a slow special-register read -- inside the loop)
  asm volatile("" : "+r"(cw_off), "+r"(c_shift), "+r"(m_off));
  const int nrec = rd.nrec;
  int rec = 0;
  while (rec < nrec) {
    rd.wait(rec);
    const uint32_t hs = rd.rec_s(rec);
    const unsigned long long ctl = lds64(hs + 4 * 8);
    const unsigned long long d0 = lds64(hs + (5 + g) * 8);
    const int niter = lo32(ctl), cw = hi32(ctl);
    const int gpt = cw & 0xff, type = (cw >> 8) & 0xff;
    const bool last_of_slot = (cw >> 16) & 1;
    const int r0 = lo32(d0), nrows = hi32(d0);
    const double *src = type == TAPE_BWD ? X : T;
    double *dst = type == TAPE_FWD ? T : X;
    const bool accumulate = type != TAPE_MID;
    const bool writer = r0 >= 0 && a < nrows && (g & (gpt - 1)) == 0;
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
    // source row of this lane's group in (ready) record ri
    auto src_row = [&](int ri) {
      const unsigned long long cwd = lds64(rd.rec_s(ri) + cw_off);
      return (int)(unsigned)(cwd >> c_shift) + a;
    };
    auto gather = [&](int row, double (&dstv)[R]) {
      if (a < DH && !(rd.dbg & 1)) {
        const double *yj = src + (size_t)row * R;
#pragma unroll
        for (int q = 0; q < R; ++q) dstv[q] = yj[q];
      }
    };
    const int last = rec + niter - 1;
#pragma unroll
    for (int s2 = 0; s2 < PFD; ++s2)
      if (s2 < niter) { rd.wait(rec + s2); gather(src_row(rec + s2), yq[s2]); }
    // the barrier of the next look-ahead record is probed one iteration early
    uint32_t ok = (rec + PFD <= last) ? rd.probe(rec + PFD) : 0u;
    for (int it0 = 0; it0 < niter; it0 += PFD) {
#pragma unroll
      for (int u = 0; u < PFD; ++u) {
        const int ri = rec + it0 + u;
        if (ri <= last) {
          const bool more = ri + PFD <= last;
          int nrow = 0;
          if (more) { rd.confirm(ri + PFD, ok); nrow = src_row(ri + PFD); }
          const uint32_t ms = rd.rec_s(ri) + m_off;
          double m[DH];
          if (a < DH) {
            if constexpr (DH == 4) {
              const double2 m01 = lds_f64x2(ms), m23 = lds_f64x2(ms + 16);
              m[0] = m01.x; m[1] = m01.y; m[2] = m23.x; m[3] = m23.y;
            } else {
#pragma unroll
              for (int i = 0; i < DH; ++i) m[i] = lds_f64(ms + 8 * i);
            }
          } else {
#pragma unroll
            for (int i = 0; i < DH; ++i) m[i] = 0.0;
          }
          if (ri + PFD + 1 <= last) ok = rd.probe(ri + PFD + 1);
#pragma unroll
          for (int i = 0; i < DH; ++i)
#pragma unroll
            for (int q = 0; q < R; ++q) acc[i][q] = fma(m[i], yq[u][q], acc[i][q]);
          if (more) gather(nrow, yq[u]);
          rd.refill();
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

}  // namespace
}  // namespace dcora
