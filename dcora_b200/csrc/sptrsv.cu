// K5: preconditioner application  OUT = proj_X( (Q + reg I)^-1 V )  as a BLOCK
// triangular solve on the host-built Cholesky factor (replaces
// graph_->preconditioner()->solve(INVEC^T)^T, src/QuadraticProblem.cpp:79).
//
// Schedule (trsv_host.cpp): the factor is grouped by the nested-dissection
// blocks; per level of the (log-depth) block elimination tree
//   phase A  t_B = rhs_B - M[B, solved] y     (d+1)x(d+1) off-diagonal blocks
//   phase B  y_B = D_B t_B                    dense packed triangular inverse
// forward with M = L, then backward with M = L^T, in ONE launch.
//
// One thread-block CLUSTER per agent (up to 8 CTAs x 8 warps on 8 SMs of one
// GPC); phases are separated by the hardware cluster barrier, so agents never
// wait for each other.  Every warp executes its own pre-compiled TAPE (trsv.h):
// the factor values arrive in consumption order through a per-warp shared-
// memory ring that is filled by bulk asynchronous copies (cp.async.bulk +
// mbarrier, i.e. TMA) running up to 8 KB ahead of the consumer and straight
// through the level barriers; only the gathers from the solution panel wait
// for a barrier.  The r right-hand sides are processed together; the panel is
// (factor row, r) row-major, the same memory shape as every other vector.
#include <cooperative_groups.h>

#include <mutex>

#include "kernels.cuh"
#include "trsv.h"

namespace cg = cooperative_groups;

namespace dcora {

void DeviceFactor::release() {
  cudaFree(tape); cudaFree(tape_off); cudaFree(tape_chunks); cudaFree(nslots);
  cudaFree(arow0); cudaFree(frow_col);
  *this = DeviceFactor();
}

namespace {

constexpr int TRSV_THREADS = 256;
constexpr int TRSV_NW = TRSV_THREADS / 32;      // warps per CTA
constexpr int CHW = TAPE_CHUNK_WORDS;           // words per chunk (2 KB)
constexpr int NCH = 4;                          // chunks per ring (8 KB per warp): a phase-A
                                                // iteration is ~1 KB, so 2 KB chunks halve the
                                                // ring bookkeeping (arm + bulk copy) per iteration
constexpr int RING_WORDS = NCH * CHW;
constexpr int L2_AHEAD = 8;                     // chunks (16 KB) of L2 prefetch distance per warp
constexpr int TRSV_SMEM = TRSV_NW * (RING_WORDS * 8 + NCH * 8);
static_assert((RING_WORDS & (RING_WORDS - 1)) == 0, "ring size must be a power of two");

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int CL>
__device__ __forceinline__ void phase_sync() {
  if constexpr (CL == 1) __syncthreads();
  else cg::this_cluster().sync();
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// Per-warp reader of a tape through the shared-memory ring.
struct TapeReader {
  const unsigned long long *src;  // this warp's tape (global)
  unsigned long long *ring;       // RING_WORDS words (shared)
  uint32_t ring_s, mbar_s;        // shared-space addresses
  int total;                      // chunks on the tape
  int issued, ready, freed;       // chunk counters (warp-uniform)
  int lane;

  __device__ __forceinline__ void issue_one() {
    // lane 0: arm the slot's mbarrier and start the bulk copy of chunk `issued`
    const int slot = issued & (NCH - 1);
    const uint32_t mb = mbar_s + slot * 8;
    const uint32_t dst = ring_s + slot * (CHW * 8);
    const unsigned long long *g = src + (size_t)issued * CHW;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(mb), "r"(CHW * 8) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(dst), "l"(g), "r"(CHW * 8), "r"(mb) : "memory");
    // The ring holds 8 KB per warp: at an HBM round trip of > 1 us that alone would cap a
    // warp's stream at a few GB/s.  The tape is static, so the chunk L2_AHEAD chunks further
    // on is pulled into L2 now; by the time the ring asks for it, it is an L2 hit.
    if (issued + L2_AHEAD < total)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                   :: "l"(g + (size_t)L2_AHEAD * CHW), "r"(CHW * 8) : "memory");
  }
  __device__ __forceinline__ void init(const unsigned long long *tape, int nchunks,
                                       unsigned long long *ring_, unsigned long long *mbar_,
                                       int lane_) {
    src = tape; ring = ring_; total = nchunks; lane = lane_;
    ring_s = smem_u32(ring_); mbar_s = smem_u32(mbar_);
    issued = 0; ready = 0; freed = 0;
    if (lane == 0) {
      for (int s = 0; s < NCH; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar_s + s * 8) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int n0 = total < NCH ? total : NCH;
    if (lane == 0) {
      const int np = total < NCH + L2_AHEAD ? total : NCH + L2_AHEAD;
      for (int c = NCH; c < np; ++c)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                     :: "l"(src + (size_t)c * CHW), "r"(CHW * 8) : "memory");
      for (; issued < n0; ++issued) issue_one();
    }
    issued = n0;
  }
  // make words [.., upto) readable
  __device__ __forceinline__ void ensure(int upto) {
    while (ready * CHW < upto) {
      const uint32_t mb = mbar_s + (ready & (NCH - 1)) * 8;
      const uint32_t parity = (ready / NCH) & 1;
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
  // words below `pos` are consumed: recycle their chunk.  Callers advance by less
  // than one chunk between calls, so at most one chunk is freed per call.
  __device__ __forceinline__ void release(int pos) {
    if (freed < pos / CHW) {
      __syncwarp();
      if (issued < total) {
        if (lane == 0) issue_one();
        ++issued;
      }
      ++freed;
    }
  }
  // wait for every copy still in flight (before the CTA may exit)
  __device__ __forceinline__ void drain() { ensure(issued * CHW); }
  __device__ __forceinline__ unsigned long long word(int i) const {
    return ring[i & (RING_WORDS - 1)];
  }
};

__device__ __forceinline__ int lo32(unsigned long long u) { return (int)(unsigned)(u & 0xffffffffull); }
__device__ __forceinline__ int hi32(unsigned long long u) { return (int)(unsigned)(u >> 32); }

// Executes every slot of one warp's tape (format: trsv.h).  Lane (g, a) of the
// warp -- group g = lane / 4, a = lane % 4 -- owns COLUMN a of its group's
// tile: per iteration it needs one row of the source panel (R doubles) and DH
// tape words, and accumulates all DH rows of the product; the 4 lanes of a
// group (and the `gpt` groups of a task) are summed once per round.
template <int R, int DH, int CL>
__device__ __forceinline__ void trsv_run(TapeReader &rd, int nslots, double *T, double *X,
                                         unsigned long long *&trace) {
  constexpr int BS = DH * DH;
  constexpr int ITW = 4 + 8 * BS;  // words per iteration
  // The source rows are gathered from L2 PFD iterations ahead of their use
  // (register ring, statically indexed by unrolling): 3 deep while the
  // registers allow 2 CTAs per SM (r <= 8), 1 deep for higher ranks.
  constexpr int PFD = R <= 8 ? 3 : 1;
  const int lane = rd.lane;
  const int g = lane >> 2, a = lane & 3;
  const int cw_off = g >> 1;
  const bool c_hi = (g & 1) != 0;
  const int m_off = 4 + (g * DH + a) * DH;
  int pos = 0;
  for (int sl = 0; sl < nslots; ++sl) {
    rd.ensure(pos + TAPE_SLOT_HEADER);
    const int nrounds = lo32(rd.word(pos));
    pos += TAPE_SLOT_HEADER;
    for (int rnd = 0; rnd < nrounds; ++rnd) {
      rd.ensure(pos + TAPE_ROUND_HEADER);
      const unsigned long long rh = rd.word(pos);
      const unsigned long long d0 = rd.word(pos + 1 + g);
      const int niter = lo32(rh), gpt = hi32(rh) & 0xff, type = hi32(rh) >> 8;
      const int r0 = lo32(d0), nrows = hi32(d0);
      pos += TAPE_ROUND_HEADER;
      const double *src = type == TAPE_BWD ? X : T;
      double *dst = type == TAPE_FWD ? T : X;
      const bool accumulate = type != TAPE_MID;
      const bool writer = r0 >= 0 && a < nrows && (g & (gpt - 1)) == 0;
      const size_t o = (size_t)(r0 + a) * R;
      // the destination's current value (accumulating tasks): independent of
      // everything below, so its latency hides behind the loop
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
      const int base = pos;
      auto gather = [&](int itn, double (&dstv)[R]) {
        const int wpos = base + itn * ITW;
        rd.ensure(wpos + 4);
        const unsigned long long cw = rd.word(wpos + cw_off);
        const int c = c_hi ? hi32(cw) : lo32(cw);
        if (a < DH) {
          const double *yj = src + (size_t)(c + a) * R;
#pragma unroll
          for (int q = 0; q < R; ++q) dstv[q] = yj[q];
        }
      };
#pragma unroll
      for (int s2 = 0; s2 < PFD; ++s2)
        if (s2 < niter) gather(s2, yq[s2]);
      for (int it0 = 0; it0 < niter; it0 += PFD) {
#pragma unroll
        for (int u = 0; u < PFD; ++u) {
          const int it = it0 + u;
          if (it < niter) {
            rd.ensure(pos + ITW);
            double m[DH];
            if (a < DH) {
              const int mb = (pos + m_off) & (RING_WORDS - 1);
              if constexpr (DH == 4) {
                const double2 *mp = reinterpret_cast<const double2 *>(rd.ring + mb);
                const double2 m01 = mp[0], m23 = mp[1];
                m[0] = m01.x; m[1] = m01.y; m[2] = m23.x; m[3] = m23.y;
              } else {
#pragma unroll
                for (int i = 0; i < DH; ++i)
                  m[i] = __longlong_as_double((long long)rd.ring[(pos + m_off + i) & (RING_WORDS - 1)]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < DH; ++i) m[i] = 0.0;
            }
#pragma unroll
            for (int i = 0; i < DH; ++i)
#pragma unroll
              for (int q = 0; q < R; ++q) acc[i][q] = fma(m[i], yq[u][q], acc[i][q]);
            if (it + PFD < niter) gather(it + PFD, yq[u]);
            pos += ITW;
            rd.release(pos);
          }
        }
      }
      // sum over the 4 lanes of the group, leaving row a in lane a
      double row[R];
      {
        double v[2][R];
        const bool hi2 = (a & 2) != 0, hi1 = (a & 1) != 0;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double mine = hi2 ? acc[2 + h][q] : acc[h][q];
            const double send = hi2 ? acc[h][q] : acc[2 + h][q];
            v[h][q] = mine + __shfl_xor_sync(0xffffffffu, send, 2);
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
      rd.release(pos);
    }
    phase_sync<CL>();
    if (trace) *trace++ = globaltimer_ns();
  }
  // bring pos to the reader's notion for drain(): nothing to do
}

// One cluster per agent.  T <- permuted right-hand side, the slots of the
// tape, then OUT <- solution X in internal column order.
template <int R, int DH, int CL>
__global__ void __launch_bounds__(TRSV_THREADS)
trsv_tape_kernel(const int *__restrict__ flags, int mask_flag,
                 const int *__restrict__ agent_list,
                 const unsigned long long *__restrict__ tape,
                 const long long *__restrict__ tape_off, const int *__restrict__ tape_chunks,
                 const int *__restrict__ nslots, const long long *__restrict__ arow0,
                 const int *__restrict__ frow_col, const double *__restrict__ V, double *T,
                 double *X, double *OUT, unsigned long long *trace_buf, int trace_agent) {
  const int slot = blockIdx.x / CL;
  const int agent = agent_list ? agent_list[slot] : slot;
  if (mask_flag >= 0 && flags[agent * NF + mask_flag] == 0) return;  // whole cluster
  const int crank = blockIdx.x - slot * CL;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = crank * TRSV_NW + warp;
  constexpr int nwarps = CL * TRSV_NW;
  extern __shared__ __align__(128) unsigned long long trsv_smem[];
  unsigned long long *ring = trsv_smem + warp * RING_WORDS;
  unsigned long long *mbar = trsv_smem + TRSV_NW * RING_WORDS + warp * NCH;
  const size_t wi = (size_t)agent * nwarps + gwarp;
  TapeReader rd;
  rd.init(tape + tape_off[wi], tape_chunks[wi], ring, mbar, lane);  // the tape starts streaming
  // optional phase trace (debug / profiling aid): thread 0 of the traced
  // agent's first CTA stamps the end of every slot
  unsigned long long *trace =
      (trace_buf && agent == trace_agent && gwarp == 0 && lane == 0) ? trace_buf : nullptr;
  if (trace) *trace++ = globaltimer_ns();
  // T[f] = V[column of factor row f]
  const long long f0 = arow0[agent], f1 = arow0[agent + 1];
  const int tid = crank * TRSV_THREADS + threadIdx.x;
  constexpr int NT = CL * TRSV_THREADS;
  for (long long e = (long long)f0 * R + tid; e < f1 * R; e += NT) {
    const long long f = e / R;
    const int q = (int)(e - f * R);
    T[e] = V[(size_t)frow_col[f] * R + q];
  }
  phase_sync<CL>();
  if (trace) *trace++ = globaltimer_ns();
  trsv_run<R, DH, CL>(rd, nslots[agent], T, X, trace);
  for (long long e = (long long)f0 * R + tid; e < f1 * R; e += NT) {
    const long long f = e / R;
    const int q = (int)(e - f * R);
    OUT[(size_t)frow_col[f] * R + q] = X[e];
  }
  if (trace) { *trace++ = globaltimer_ns(); *trace = 0ull; }
  rd.drain();
}

template <typename T>
cudaError_t up(T **dst, const std::vector<T> &src) {
  cudaError_t e = cudaMalloc(dst, std::max<size_t>(1, src.size()) * sizeof(T));
  if (e != cudaSuccess) return e;
  if (src.empty()) return cudaSuccess;
  return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

}  // namespace

int trsv_cluster_for(double rows_per_agent) {
  // enough CTAs to spread one agent's factor over several SMs; a single CTA
  // for tiny agents (the level barriers become __syncthreads)
  return rows_per_agent < 400.0 ? 1 : rows_per_agent < 1500.0 ? 2
         : rows_per_agent < 4000.0 ? 4 : 8;
}
int trsv_warps_per_cta() { return TRSV_NW; }

int upload_factor(DeviceMatrices *M, const std::vector<HostFactor> &HF,
                  std::vector<AgentTape> &AT, int cluster, const HostLayout &H) {
  DeviceFactor &F = M->F;
  const int A = H.A;
  const int nwarps = cluster * TRSV_NW;
  int64_t nnzL = 0, stored = 0, words = 0;
  int max_levels = 0;
  std::vector<long long> off((size_t)A * nwarps), arow0((size_t)A + 1);
  std::vector<int> chunks((size_t)A * nwarps), nslots((size_t)A), frow_col((size_t)H.K);
  F.crit_iters = F.ideal_iters = F.crit_rounds = 0;
  for (int a = 0; a < A; ++a) {
    nnzL += HF[a].nnz;
    stored += AT[a].stored_values;
    if (AT[a].nwarps != nwarps) {
      set_error("upload_factor: tape built for another cluster size");
      return DCORA_ERR_ARG;
    }
    for (int w = 0; w < nwarps; ++w) {
      off[(size_t)a * nwarps + w] = words;
      chunks[(size_t)a * nwarps + w] = (int)(AT[a].warp_tape[w].size() / TAPE_CHUNK_WORDS);
      words += (int64_t)AT[a].warp_tape[w].size();
    }
    nslots[a] = AT[a].nslots;
    arow0[a] = H.coff[a];
    const int k = (int)(H.coff[a + 1] - H.coff[a]);
    for (int f = 0; f < k; ++f) frow_col[(size_t)H.coff[a] + f] = (int)H.coff[a] + HF[a].perm[f];
    max_levels = std::max(max_levels, std::max(AT[a].nlev[0], AT[a].nlev[1]));
    F.crit_iters = std::max(F.crit_iters, AT[a].crit_iters);
    F.crit_rounds = std::max(F.crit_rounds, AT[a].crit_rounds);
    F.ideal_iters = std::max(F.ideal_iters, AT[a].ideal_iters);
    F.max_slots = std::max(F.max_slots, AT[a].nslots);
  }
  arow0[A] = H.coff[A];
  DCORA_CUDA(cudaMalloc(&F.tape, std::max<int64_t>(1, words) * sizeof(unsigned long long)));
  for (int a = 0; a < A; ++a) {
    for (int w = 0; w < nwarps; ++w) {
      const std::vector<uint64_t> &tp = AT[a].warp_tape[w];
      if (!tp.empty())
        DCORA_CUDA(cudaMemcpy(F.tape + off[(size_t)a * nwarps + w], tp.data(),
                              tp.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    AT[a] = AgentTape();  // release host memory early
  }
  DCORA_CUDA(up(&F.tape_off, off));
  DCORA_CUDA(up(&F.tape_chunks, chunks));
  DCORA_CUDA(up(&F.nslots, nslots));
  DCORA_CUDA(up(&F.arow0, arow0));
  DCORA_CUDA(up(&F.frow_col, frow_col));
  F.present = true;
  F.nnzL = nnzL;
  F.stored_values = stored / 2;
  F.tape_bytes = words * 8;
  F.A = A;
  F.max_levels = max_levels;
  F.cluster = cluster;
  return DCORA_OK;
}

// cudaFuncSetAttribute once per (instantiation, device); handles of different
// devices may be driven from different host threads
template <int R, int DH, int CL>
static cudaError_t trsv_set_smem(int device, int smem_bytes) {
  static std::mutex mtx;
  static unsigned long long done = 0;  // bit d: set on device d
  std::lock_guard<std::mutex> lock(mtx);
  if (device < 64 && (done & (1ull << device))) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(trsv_tape_kernel<R, DH, CL>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e == cudaSuccess && device < 64) done |= 1ull << device;
  return e;
}

template <int R, int DH, int CL>
static int launch_trsv(dcora_problem *p, const double *V, double *T, double *X,
                       double *OUT, int mask_flag, unsigned long long *trace,
                       int trace_agent) {
  DeviceMatrices &M = *p->M;
  const DeviceFactor &F = M.F;
  cudaLaunchConfig_t cfg = {};
  const ActiveSet *as = active_set(p, mask_flag);
  if (as && as->n_agents == 0) return DCORA_OK;
  cfg.gridDim = dim3((unsigned)((as ? as->n_agents : M.A) * CL));
  cfg.blockDim = dim3(TRSV_THREADS);
  constexpr int smem_bytes = TRSV_SMEM;
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = p->stream;
  DCORA_CUDA((trsv_set_smem<R, DH, CL>(M.device, smem_bytes)));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  DCORA_CUDA(cudaLaunchKernelEx(&cfg, trsv_tape_kernel<R, DH, CL>, (const int *)p->flags,
                                mask_flag, (const int *)(as ? as->agents : nullptr),
                                (const unsigned long long *)F.tape,
                                (const long long *)F.tape_off, (const int *)F.tape_chunks,
                                (const int *)F.nslots, (const long long *)F.arow0,
                                (const int *)F.frow_col, V, T, X, OUT, trace, trace_agent));
  p->launches++;
  return DCORA_OK;
}

int launch_precond(dcora_problem *p, const double *X, const double *V, double *OUT,
                   int mask_flag, unsigned long long *trace, int trace_agent,
                   const double *dot_with) {
  DeviceMatrices &M = *p->M;
  const DeviceFactor &F = M.F;
  if (!F.present) {
    // Graph::hasPreconditioner() false: the reference only projects
    // (QuadraticProblem.cpp:78-83)
    if (!X) return launch_copy(p, OUT, V, mask_flag);
    return launch_project_tangent(p, X, V, OUT, mask_flag, dot_with);
  }
  double *panel = p->vec(WS_TRSV_PANEL);
  double *tmp = p->vec(WS_TRSV_TMP);
  double *sol = p->vec(WS_TRSV_SOL);
  if (!panel || !sol || !tmp) return DCORA_ERR_ALLOC;
  DCORA_DISPATCH_RD(p->r, M.d, {
    switch (F.cluster) {
      case 1: DCORA_TRY((launch_trsv<R, DH, 1>(p, V, tmp, panel, sol, mask_flag, trace, trace_agent))); break;
      case 2: DCORA_TRY((launch_trsv<R, DH, 2>(p, V, tmp, panel, sol, mask_flag, trace, trace_agent))); break;
      case 4: DCORA_TRY((launch_trsv<R, DH, 4>(p, V, tmp, panel, sol, mask_flag, trace, trace_agent))); break;
      default: DCORA_TRY((launch_trsv<R, DH, 8>(p, V, tmp, panel, sol, mask_flag, trace, trace_agent))); break;
    }
  })
  if (!X) return launch_copy(p, OUT, sol, mask_flag);  // plain (Q + reg I)^-1 V
  return launch_project_tangent(p, X, sol, OUT, mask_flag, dot_with);
}

}  // namespace dcora
