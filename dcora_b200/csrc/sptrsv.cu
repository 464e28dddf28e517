// K5: preconditioner application  OUT = proj_X( (Q + reg I)^-1 V )  as a BLOCK
// triangular solve on the host-built Cholesky factor (replaces
// graph_->preconditioner()->solve(INVEC^T)^T, src/QuadraticProblem.cpp:79).
//
// Schedule (trsv_host.cpp, trsv.h): the factor is grouped by the nested-dissection blocks
// and the inverted diagonal blocks are folded into the off-diagonal part, so that the
// forward and the backward walk need ONE barrier per level of the (log-depth) block
// elimination tree; everything is (d+1)x(d+1) tile tasks executed by one device loop
// (trsv_dev.cuh), forward, mid and backward in ONE launch.
//
// One thread-block CLUSTER per agent (64 warps: 4 CTAs x 16 warps on 4 SMs for r <= 8, so
// that the <= 32 agents of an RBCD colour class each own their SMs); slots are separated by
// the hardware cluster barrier, so agents never wait for each other.  Every warp executes
// its own pre-compiled TAPE: the factor values arrive in consumption order through a
// per-warp shared-memory ring that is filled by bulk asynchronous copies (cp.async.bulk +
// mbarrier, i.e. TMA) running ahead of the consumer and straight through the barriers;
// only the gathers from the solution panels wait for a barrier.  (The alternatives that were
// measured are listed at the end of trsv_dev.cuh.)  The r right-hand sides
// are processed together; the panels are (factor row, r) row-major, the same memory shape
// as every other vector.
#include <cstdlib>
#include <atomic>
#include <mutex>

#include "trsv_dev.cuh"

namespace dcora {

void DeviceFactor::release() {
  cudaFree(tape); cudaFree(tape_off); cudaFree(tape_chunks); cudaFree(nslots);
  cudaFree(arow0); cudaFree(frow_col); cudaFree(col_frow);
  *this = DeviceFactor();
}

namespace {

// One cluster per agent.  T <- permuted right-hand side, the rounds of the
// tape, then OUT <- solution X in internal column order.
template <int R, int DH, int CL, int NW>
__global__ void __launch_bounds__(NW * 32)
trsv_tape_kernel(const int *__restrict__ flags, int mask_flag,
                 const int *__restrict__ agent_list,
                 const unsigned long long *__restrict__ tape,
                 const long long *__restrict__ tape_off, const int *__restrict__ tape_chunks,
                 const long long *__restrict__ arow0,
                 const int *__restrict__ frow_col, const double *__restrict__ V, double *T,
                 double *X, double *OUT, unsigned long long *trace_buf, int trace_agent, int dbg) {
  constexpr int RECW = TapeRec<DH>::W;
  constexpr int NTHREADS = NW * 32;
  const int slot = blockIdx.x / CL;
  const int agent = agent_list ? agent_list[slot] : slot;
  if (mask_flag >= 0 && flags[agent * NF + mask_flag] == 0) return;  // whole cluster
  const int crank = blockIdx.x - slot * CL;
  const int warp = uniform32(threadIdx.x >> 5);
  const int gwarp = crank * NW + warp;
  constexpr int nwarps = CL * NW;
  extern __shared__ __align__(128) unsigned long long trsv_smem[];
  unsigned long long *ring = trsv_smem + (size_t)warp * RING_SLOTS * RECW;
  unsigned long long *mbar = trsv_smem + (size_t)NW * RING_SLOTS * RECW + warp * RING_SLOTS;
  const size_t wi = (size_t)agent * nwarps + gwarp;
  TapeRing<RECW> rd;
  rd.setup(ring, mbar);
  rd.dbg = dbg;
  // the tape starts streaming
  rd.start(tape + uniform64((unsigned long long)tape_off[wi]), uniform32(tape_chunks[wi]));
  // optional slot trace (debug / profiling aid): thread 0 of the traced
  // agent's first CTA stamps the end of every slot
  unsigned long long *trace =
      (trace_buf && agent == trace_agent && gwarp == 0 && (threadIdx.x & 31) == 0) ? trace_buf
                                                                                   : nullptr;
  if (trace) *trace++ = globaltimer_ns();
  if (trace && (dbg & 8)) {   // per-iteration cycle stamps of this warp, behind the slot stamps
    rd.itrace = reinterpret_cast<long long *>(trace_buf) + 64;
    rd.itrace_left = 120;
  }
  // T[f] = V[column of factor row f]
  const long long f0 = arow0[agent], f1 = arow0[agent + 1];
  const int tid = crank * NTHREADS + threadIdx.x;
  constexpr int NT = CL * NTHREADS;
  trsv_permute_rows<R, false>(f0, f1, tid, NT, frow_col, V, T);
  phase_sync<CL>();
  if (trace) *trace++ = globaltimer_ns();
  trsv_run<R, DH, CL>(rd, T, X, (int)f0, trace);
  trsv_permute_rows<R, true>(f0, f1, tid, NT, frow_col, X, OUT);
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

// warps of the largest agents: 64 (4 CTAs x 16 warps) by default, 128 (8 CTAs) when the
// caller expects few agents to run concurrently (dcora_set_trsv_max_warps: a colour class of
// <= 18 agents leaves SMs idle at 4 CTAs per agent)
static std::atomic<int> g_trsv_max_warps{64};
static int legal_big_warps(int w) { return w >= 256 ? 256 : w >= 128 ? 128 : w >= 96 ? 96 : 64; }
void set_trsv_max_warps(int w) { g_trsv_max_warps.store(legal_big_warps(w)); }

int trsv_warps_for(double rows_per_agent) {
  // enough warps to spread one agent's factor over several SMs; few for tiny
  // agents (a single CTA: the level barriers become __syncthreads)
  static const int env = getenv("DCORA_TRSV_WARPS") ? atoi(getenv("DCORA_TRSV_WARPS")) : 0;
  const int big = env > 0 ? legal_big_warps(env) : g_trsv_max_warps.load();
  return rows_per_agent < 400.0 ? 8 : rows_per_agent < 1500.0 ? 16
         : rows_per_agent < 4000.0 ? 32 : rows_per_agent < 12000.0 ? 64 : big;
}

int upload_factor(DeviceMatrices *M, const std::vector<HostFactor> &HF,
                  std::vector<AgentTape> &AT, int nwarps, const HostLayout &H) {
  DeviceFactor &F = M->F;
  const int A = H.A;
  const int recw = tape_record_words(M->d + 1);
  int64_t nnzL = 0, stored = 0, words = 0;
  int max_levels = 0;
  std::vector<long long> off((size_t)A * nwarps), arow0((size_t)A + 1);
  std::vector<int> chunks((size_t)A * nwarps), nslots((size_t)A), frow_col((size_t)H.K),
      col_frow((size_t)H.K);
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
      chunks[(size_t)a * nwarps + w] = (int)(AT[a].warp_tape[w].size() / recw);
      words += (int64_t)AT[a].warp_tape[w].size();
    }
    nslots[a] = AT[a].nslots;
    arow0[a] = H.coff[a];
    const int k = (int)(H.coff[a + 1] - H.coff[a]);
    for (int f = 0; f < k; ++f) {
      frow_col[(size_t)H.coff[a] + f] = (int)H.coff[a] + HF[a].perm[f];
      col_frow[(size_t)H.coff[a] + HF[a].perm[f]] = (int)H.coff[a] + f;
    }
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
  DCORA_CUDA(up(&F.col_frow, col_frow));
  F.present = true;
  F.nnzL = nnzL;
  F.stored_values = stored / 2;
  F.tape_bytes = words * 8;
  F.A = A;
  F.max_levels = max_levels;
  F.nwarps = nwarps;
  return DCORA_OK;
}

// cudaFuncSetAttribute once per (instantiation, device); handles of different
// devices may be driven from different host threads
template <int R, int DH, int CL, int NW>
static cudaError_t trsv_set_attrs(int device, int smem_bytes) {
  static std::mutex mtx;
  static unsigned long long done = 0;  // bit d: set on device d
  std::lock_guard<std::mutex> lock(mtx);
  if (device < 64 && (done & (1ull << device))) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(trsv_tape_kernel<R, DH, CL, NW>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e == cudaSuccess && CL > 8)
    e = cudaFuncSetAttribute(trsv_tape_kernel<R, DH, CL, NW>,
                             cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e == cudaSuccess && device < 64) done |= 1ull << device;
  return e;
}

static int k5_debug_flags() {
  static const int v = getenv("DCORA_K5_DBG") ? atoi(getenv("DCORA_K5_DBG")) : 0;
  return v;
}

template <int R, int DH, int CL, int NW>
static int launch_trsv(dcora_problem *p, const double *V, double *T, double *X,
                       double *OUT, int mask_flag, unsigned long long *trace,
                       int trace_agent) {
  DeviceMatrices &M = *p->M;
  const DeviceFactor &F = M.F;
  cudaLaunchConfig_t cfg = {};
  const ActiveSet *as = active_set(p, mask_flag);
  if (as && as->n_agents == 0) return DCORA_OK;
  cfg.gridDim = dim3((unsigned)((as ? as->n_agents : M.A) * CL));
  cfg.blockDim = dim3(NW * 32);
  constexpr int smem_bytes = NW * RING_SLOTS * (TapeRec<DH>::W * 8 + 8);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = p->stream;
  DCORA_CUDA((trsv_set_attrs<R, DH, CL, NW>(M.device, smem_bytes)));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  DCORA_CUDA(cudaLaunchKernelEx(&cfg, trsv_tape_kernel<R, DH, CL, NW>, (const int *)p->flags,
                                mask_flag, (const int *)(as ? as->agents : nullptr),
                                (const unsigned long long *)F.tape,
                                (const long long *)F.tape_off, (const int *)F.tape_chunks,
                                (const long long *)F.arow0, (const int *)F.frow_col, V, T, X,
                                OUT, trace, trace_agent, k5_debug_flags()));
  p->launches++;
  return DCORA_OK;
}

// r <= 8: 16 warps per CTA (<= 128 registers per thread); higher ranks keep 8 warps per
// CTA (up to 255 registers) and twice the CTAs -- the tapes only know global warp numbers
template <int R, int DH>
static int launch_trsv_shape(dcora_problem *p, const double *V, double *T, double *X,
                             double *OUT, int mask_flag, unsigned long long *trace,
                             int trace_agent) {
  const int nw = p->M->F.nwarps;
  if constexpr (R <= 8) {
    // 64 warps run as 8 CTAs x 8 warps (up to two CTAs per SM) rather than 4 x 16: the walk is
    // bound by LSU wavefronts per SM, and the smaller CTAs spread a solve over twice the SMs
    // whenever fewer than a full colour class is active (212 vs 266 us for one agent alone,
    // equal for 28 agents; 33.2 vs 34.9 ms per RBCD iteration).  DCORA_K5_CL8=0: 4 x 16.
    static const bool cl8 = !(getenv("DCORA_K5_CL8") && atoi(getenv("DCORA_K5_CL8")) == 0);
    if (cl8 && nw == 64)
      return launch_trsv<R, DH, 8, 8>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
    // the same for the larger shapes (non-portable cluster sizes 12 / 16): one agent alone
    // 171 vs 212 us at 96 warps, 148 vs 189 us at 128 warps (profiles/r02_k5_cta_shape_big.txt)
    const bool cl8big = cl8;
    if (cl8big && nw == 96)
      return launch_trsv<R, DH, 12, 8>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
    if (cl8big && nw == 128)
      return launch_trsv<R, DH, 16, 8>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
    switch (nw) {
      case 8: return launch_trsv<R, DH, 1, 8>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
      case 16: return launch_trsv<R, DH, 1, 16>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
      case 32: return launch_trsv<R, DH, 2, 16>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
      case 96: return launch_trsv<R, DH, 6, 16>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
      case 128: return launch_trsv<R, DH, 8, 16>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
      case 256: return launch_trsv<R, DH, 16, 16>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
      default: return launch_trsv<R, DH, 4, 16>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
    }
  } else {
    if (nw > 64) {
      set_error("K5: tapes of more than 64 warps (dcora_set_trsv_max_warps) support ranks <= 8 only");
      return DCORA_ERR_ARG;
    }
    switch (nw) {
      case 8: return launch_trsv<R, DH, 1, 8>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
      case 16: return launch_trsv<R, DH, 2, 8>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
      case 32: return launch_trsv<R, DH, 4, 8>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
      default: return launch_trsv<R, DH, 8, 8>(p, V, T, X, OUT, mask_flag, trace, trace_agent);
    }
  }
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
    DCORA_TRY((launch_trsv_shape<R, DH>(p, V, tmp, panel, sol, mask_flag, trace, trace_agent)));
  })
  if (!X) return launch_copy(p, OUT, sol, mask_flag);  // plain (Q + reg I)^-1 V
  return launch_project_tangent(p, X, sol, OUT, mask_flag, dot_with);
}

}  // namespace dcora

extern "C" void dcora_set_trsv_max_warps(int32_t warps) { dcora::set_trsv_max_warps(warps); }
