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

#include "kernels.cuh"
#include "trsv.h"

namespace cg = cooperative_groups;

namespace dcora {

void DeviceFactor::release() {
  cudaFree(tape); cudaFree(tape_off); cudaFree(tape_chunks); cudaFree(nlev);
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
constexpr int TRSV_SMEM = TRSV_NW * (RING_WORDS * 8 + NCH * 8);

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
    if (lane == 0)
      for (; issued < n0; ++issued) issue_one();
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
  // words below `pos` are consumed: recycle their chunks
  __device__ __forceinline__ void release(int pos) {
    const int done = pos / CHW;
    if (freed < done) {
      __syncwarp();
      while (freed < done) {
        if (issued < total) {
          if (lane == 0) issue_one();
          ++issued;
        }
        ++freed;
      }
    }
  }
  // wait for every copy still in flight (before the CTA may exit)
  __device__ __forceinline__ void drain() { ensure(issued * CHW); }
  __device__ __forceinline__ unsigned long long word(int i) const {
    return ring[i & (RING_WORDS - 1)];
  }
  __device__ __forceinline__ double dword(int i) const {
    return __longlong_as_double((long long)ring[i & (RING_WORDS - 1)]);
  }
};

__device__ __forceinline__ int lo32(unsigned long long u) { return (int)(unsigned)(u & 0xffffffffull); }
__device__ __forceinline__ int hi32(unsigned long long u) { return (int)(unsigned)(u >> 32); }

// Walks one direction of one agent's tape.  MODE 0: forward (right-hand side
// gathered from V through the embedded permutation), MODE 1: backward (rhs =
// forward solution in Y; the result overwrites Y in place and is scattered to
// OUT through the permutation).
template <int R, int DH, int CL, int MODE>
__device__ __forceinline__ void trsv_walk(TapeReader &rd, int &pos, int nlev,
                                          const double *__restrict__ V, double *T,
                                          double *Y, double *OUT,
                                          unsigned long long *&trace, double *scratch) {
  constexpr int BS = DH * DH;
  constexpr int ITW = 4 + 8 * BS;  // words per phase-A iteration
  const int lane = rd.lane;
  const int g = lane >> 2, a = lane & 3;
  for (int lv = 0; lv < nlev; ++lv) {
    // ------------------------------------------------------------ phase A
    rd.ensure(pos + 1);
    const unsigned long long ah = rd.word(pos);
    ++pos;
    const int nrounds = lo32(ah);
    for (int rnd = 0; rnd < nrounds; ++rnd) {
      rd.ensure(pos + 17);
      const unsigned long long rh = rd.word(pos);
      const int niter = lo32(rh), gpt = hi32(rh);
      const unsigned long long d0 = rd.word(pos + 1 + 2 * g), d1 = rd.word(pos + 2 + 2 * g);
      const int r0 = lo32(d0), nrows = hi32(d0), pbase = lo32(d1);
      pos = (pos + 17 + 3) & ~3;  // the iterations start 32-byte aligned on the tape
      // lane (g, a) owns COLUMN a of its group's block: it needs one row of the
      // source pose (R doubles of Y) per block and accumulates all DH rows of
      // the product; the 4 lanes of a group are summed once per round.
      double acc[4][R];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < R; ++q) acc[i][q] = 0.0;
      // The source rows of Y are gathered from L2 PFD iterations ahead of their use
      // (register ring): 3 deep while the registers allow 2 CTAs per SM (r <= 8),
      // 1 deep for higher ranks.  (Measured: 360 -> 352 us per solve; the phase is
      // bound by the instruction stream of the tape walk, not by this latency.)
      constexpr int PFD = R <= 8 ? 3 : 1;
      double yq[PFD][R];
#pragma unroll
      for (int s2 = 0; s2 < PFD; ++s2)
#pragma unroll
        for (int q = 0; q < R; ++q) yq[s2][q] = 0.0;
      auto gather = [&](int ahead, double (&dst)[R]) {
        rd.ensure(pos + (ahead + 1) * ITW);
        const unsigned long long cw = rd.word(pos + ahead * ITW + (g >> 1));
        const int c = (g & 1) ? hi32(cw) : lo32(cw);
        if (a < DH) {
          const double *yj = Y + (size_t)(c + a) * R;
#pragma unroll
          for (int q = 0; q < R; ++q) dst[q] = yj[q];
        }
      };
#pragma unroll
      for (int s2 = 0; s2 < PFD; ++s2)
        if (niter > s2) gather(s2, yq[s2]);
      for (int it = 0; it < niter; ++it) {
        double yc[R], m[DH];
#pragma unroll
        for (int q = 0; q < R; ++q) {
          yc[q] = yq[0][q];
#pragma unroll
          for (int s2 = 0; s2 + 1 < PFD; ++s2) yq[s2][q] = yq[s2 + 1][q];
        }
        if (a < DH) {
          // lane (g, a) owns DH consecutive tape words: column a of its block
          const int base = pos + 4 + (g * DH + a) * DH;
          if constexpr (DH == 4) {
            const double2 *mp =
                reinterpret_cast<const double2 *>(rd.ring + (base & (RING_WORDS - 1)));
            const double2 m01 = mp[0], m23 = mp[1];
            m[0] = m01.x; m[1] = m01.y; m[2] = m23.x; m[3] = m23.y;
          } else {
#pragma unroll
            for (int i = 0; i < DH; ++i) m[i] = rd.dword(base + i);
          }
        } else {
#pragma unroll
          for (int i = 0; i < DH; ++i) m[i] = 0.0;
        }
        if (it + PFD < niter) gather(PFD, yq[PFD - 1]);
#pragma unroll
        for (int i = 0; i < DH; ++i)
#pragma unroll
          for (int q = 0; q < R; ++q) acc[i][q] = fma(m[i], yc[q], acc[i][q]);
        pos += ITW;
        rd.release(pos);
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
      for (int o = 4; o < 4 * gpt; o <<= 1) {
#pragma unroll
        for (int q = 0; q < R; ++q) row[q] += __shfl_xor_sync(0xffffffffu, row[q], o);
      }
      if (r0 >= 0 && a < nrows && (g & (gpt - 1)) == 0) {
        const size_t o = (size_t)(r0 + a) * R;
        const double *rhs = MODE == 0 ? V + (size_t)(pbase + a) * R : Y + o;
#pragma unroll
        for (int q = 0; q < R; ++q) T[o + q] = rhs[q] - row[q];
      }
      rd.release(pos);
    }
    if (hi32(ah) != 0) {  // the level has a phase A (agent-uniform)
      phase_sync<CL>();
      if (trace) *trace++ = globaltimer_ns() | 1ull;  // odd: end of a phase A
    }
    // ------------------------------------------------------------ phase B
    rd.ensure(pos + 1);
    const unsigned long long bh = rd.word(pos);
    ++pos;
    const int nsmall = lo32(bh), nlarge = hi32(bh);
    // small blocks: lane = row, packed triangle consumed column by column
    for (int s = 0; s < nsmall; ++s) {
      rd.ensure(pos + 2);
      const unsigned long long h0 = rd.word(pos), h1 = rd.word(pos + 1);
      const int r0 = lo32(h0), w = hi32(h0);
      const bool direct = MODE == 0 && lo32(h1) != 0;
      pos += 2;
      const int npw = (w + 1) >> 1;
      rd.ensure(pos + npw);
      int pm = 0;
      if (lane < w) {
        const unsigned long long pw = rd.word(pos + (lane >> 1));
        pm = (lane & 1) ? hi32(pw) : lo32(pw);
      }
      pos += npw;
      double tv[R], acc[R];
#pragma unroll
      for (int q = 0; q < R; ++q) { tv[q] = 0.0; acc[q] = 0.0; }
      if (lane < w) {
        const double *src = direct ? V + (size_t)pm * R : T + (size_t)(r0 + lane) * R;
#pragma unroll
        for (int q = 0; q < R; ++q) tv[q] = src[q];
      }
      // the block's right-hand side goes through this warp's shared scratch: row j is
      // then a broadcast read (R shared loads) instead of 2 R shuffles per column
      __syncwarp();
      if (lane < w) {
#pragma unroll
        for (int q = 0; q < R; ++q) scratch[lane * R + q] = tv[q];
      }
      __syncwarp();
      rd.ensure(pos + w * (w + 1) / 2);  // the whole packed triangle (<= 528 words)
      {
        int off = pos;
#pragma unroll 4
        for (int j = 0; j < w; ++j) {
          double m = 0.0;
          if (MODE == 0) {            // column j holds rows j..w-1
            if (lane >= j && lane < w) m = rd.dword(off + (lane - j));
            off += w - j;
          } else {                    // column j holds rows 0..j
            if (lane <= j) m = rd.dword(off + lane);
            off += j + 1;
          }
#pragma unroll
          for (int q = 0; q < R; ++q) acc[q] = fma(m, scratch[j * R + q], acc[q]);
        }
        pos = off;
      }
      rd.release(pos);
      if (lane < w) {
        const size_t o = (size_t)(r0 + lane) * R;
#pragma unroll
        for (int q = 0; q < R; ++q) Y[o + q] = acc[q];
        if (MODE == 1) {
          const size_t oo = (size_t)pm * R;
#pragma unroll
          for (int q = 0; q < R; ++q) OUT[oo + q] = acc[q];
        }
      }
    }
    // rows of large blocks: lanes stride along the row
    for (int l = 0; l < nlarge; ++l) {
      rd.ensure(pos + 2);
      const unsigned long long h0 = rd.word(pos), h1 = rd.word(pos + 1);
      const int src0 = lo32(h0), len = hi32(h0);
      const int dest = lo32(h1), dperm = hi32(h1);
      pos += 2;
      double acc[R];
#pragma unroll
      for (int q = 0; q < R; ++q) acc[q] = 0.0;
      const double *srcp = T + (size_t)src0 * R;
      for (int j0 = 0; j0 < len; j0 += CHW) {
        const int n = len - j0 < CHW ? len - j0 : CHW;
        rd.ensure(pos + n);
#pragma unroll 4
        for (int j = lane; j < n; j += 32) {
          const double m = rd.dword(pos + j);
          const double *sv = srcp + (size_t)(j0 + j) * R;
#pragma unroll
          for (int q = 0; q < R; ++q) acc[q] = fma(m, sv[q], acc[q]);
        }
        pos += n;
        rd.release(pos);
      }
#pragma unroll
      for (int q = 0; q < R; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
      }
      if (lane < R) {
        double sres = 0.0;
#pragma unroll
        for (int q = 0; q < R; ++q)
          if (q == lane) sres = acc[q];
        Y[(size_t)dest * R + lane] = sres;
        if (MODE == 1) OUT[(size_t)dperm * R + lane] = sres;
      }
    }
    phase_sync<CL>();
    if (trace) *trace++ = globaltimer_ns() & ~1ull;  // even: end of a phase B
  }
}

template <int R, int DH, int CL>
__global__ void __launch_bounds__(TRSV_THREADS)
trsv_tape_kernel(const int *__restrict__ flags, int mask_flag,
                 const int *__restrict__ agent_list,
                 const unsigned long long *__restrict__ tape,
                 const long long *__restrict__ tape_off, const int *__restrict__ tape_chunks,
                 const int *__restrict__ nlev, const double *__restrict__ V, double *T,
                 double *Y, double *OUT, unsigned long long *trace_buf, int trace_agent) {
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
  double *scratch = reinterpret_cast<double *>(trsv_smem + TRSV_NW * (RING_WORDS + NCH)) +
                    warp * (TRSV_SMALL_BLOCK * R);
  const size_t wi = (size_t)agent * nwarps + gwarp;
  TapeReader rd;
  rd.init(tape + tape_off[wi], tape_chunks[wi], ring, mbar, lane);
  // optional phase trace (debug / profiling aid): thread 0 of the traced
  // agent's first CTA stamps the end of every phase
  unsigned long long *trace =
      (trace_buf && agent == trace_agent && gwarp == 0 && lane == 0) ? trace_buf : nullptr;
  if (trace) *trace++ = globaltimer_ns();
  int pos = 0;
  trsv_walk<R, DH, CL, 0>(rd, pos, nlev[agent * 2 + 0], V, T, Y, OUT, trace, scratch);
  trsv_walk<R, DH, CL, 1>(rd, pos, nlev[agent * 2 + 1], V, T, Y, OUT, trace, scratch);
  if (trace) *trace = 0ull;
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
  std::vector<long long> off((size_t)A * nwarps);
  std::vector<int> chunks((size_t)A * nwarps), nlev((size_t)A * 2);
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
    nlev[a * 2 + 0] = AT[a].nlev[0];
    nlev[a * 2 + 1] = AT[a].nlev[1];
    max_levels = std::max(max_levels, std::max(AT[a].nlev[0], AT[a].nlev[1]));
  }
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
  DCORA_CUDA(up(&F.nlev, nlev));
  F.present = true;
  F.nnzL = nnzL;
  F.stored_values = stored / 2;
  F.tape_bytes = words * 8;
  F.A = A;
  F.max_levels = max_levels;
  F.cluster = cluster;
  return DCORA_OK;
}

template <int R, int DH, int CL>
static int launch_trsv(dcora_problem *p, const double *V, double *T, double *Y,
                       double *OUT, int mask_flag, unsigned long long *trace,
                       int trace_agent) {
  DeviceMatrices &M = *p->M;
  const DeviceFactor &F = M.F;
  cudaLaunchConfig_t cfg = {};
  const ActiveSet *as = active_set(p, mask_flag);
  if (as && as->n_agents == 0) return DCORA_OK;
  cfg.gridDim = dim3((unsigned)((as ? as->n_agents : M.A) * CL));
  cfg.blockDim = dim3(TRSV_THREADS);
  constexpr int smem_bytes = TRSV_SMEM + TRSV_NW * TRSV_SMALL_BLOCK * R * 8;  // + scratch
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = p->stream;
  static unsigned attr_done = 0;  // per instantiation: bit d = set on device d
  if (M.device >= 32 || !(attr_done & (1u << M.device))) {
    DCORA_CUDA(cudaFuncSetAttribute(trsv_tape_kernel<R, DH, CL>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    if (M.device < 32) attr_done |= 1u << M.device;
  }
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
                                (const int *)F.nlev, V, T, Y, OUT, trace, trace_agent));
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
