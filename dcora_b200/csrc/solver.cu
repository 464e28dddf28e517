// K3/K4: Riemannian trust-region (RTR) with truncated CG, run for a whole batch
// of agents at once on the device.
//
// Replaces ROPTLIB::RTRNewton::Run / SolversTR::tCG_TR as driven by
// QuadraticOptimizer (src/QuadraticOptimizer.cpp:28-108,234-280).  ROPTLIB is
// not vendored in the reference tree (cmake/roptlib.cmake:7-8 fetches
// yuluntian/ROPTLIB@feature/cmake); the algorithm is restated from upstream:
// theta = 1, kappa = 0.1, Acceptence_Rho = 0.1, Shrinked_tau = 0.25,
// Magnified_tau = 2, radius measured in the preconditioner norm.
//
// Every agent keeps its own trust-region state in p->scal / p->flags; vector
// kernels are masked per agent, so agents converge independently while sharing
// launches.  The host only reads back two counters per iteration.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>

#include "solver.cuh"

namespace dcora {

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum of partial slot `slot` over agent a's chunks (fixed order); valid in
// every thread of the 128-thread CTA.
__device__ double agent_sum(const int *agent_chunk0, const double *partials, int a,
                            int slot, double *sm) {
  const int c0 = agent_chunk0[a], c1 = agent_chunk0[a + 1];
  double acc = 0.0;
  for (int c = c0 + threadIdx.x; c < c1; c += 128) acc += partials[(size_t)c * NPART + slot];
  acc = warp_sum(acc);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  return sm[0] + sm[1] + sm[2] + sm[3];
}

__global__ void __launch_bounds__(128)
combine_f_kernel(const int *list, const int *agent_chunk0, const double *partials, const int *flags,
                 int mask_flag, int slot, int with_G, double *scal) {
  __shared__ double sm[4];
  const int a = list ? list[blockIdx.x] : blockIdx.x;
  if (mask_flag >= 0 && flags[a * NF + mask_flag] == 0) return;
  const double p0 = agent_sum(agent_chunk0, partials, a, 0, sm);
  const double p1 = with_G ? agent_sum(agent_chunk0, partials, a, 1, sm) : 0.0;
  if (threadIdx.x == 0) scal[a * NS + slot] = 0.5 * p0 + p1;
}

// after f1 (S_F1) and <gf,gf> (partial 0)
__global__ void __launch_bounds__(128)
rtr_init_kernel(const int *list, const int *agent_chunk0, const double *partials, int *flags,
                double *scal, double tol, double delta0, double max_delta,
                int *counter, int *counter_next) {
  __shared__ double sm[4];
  if (blockIdx.x == 0 && threadIdx.x == 0) *counter_next = 0;   // the next counting kernel's slot
  const int a = list ? list[blockIdx.x] : blockIdx.x;
  const double gg = agent_sum(agent_chunk0, partials, a, 0, sm);
  if (threadIdx.x != 0) return;
  double *s = scal + a * NS;
  int *f = flags + a * NF;
  const double ngf = sqrt(gg);
  s[S_NGF] = ngf;
  s[S_FINIT] = s[S_F1];
  s[S_GNINIT] = ngf;
  s[S_DELTA] = delta0;
  s[S_MAXDELTA] = max_delta;
  s[S_TOL] = tol;
  // QuadraticOptimizer::trustRegion early return (:54)
  const int on = f[F_USER] != 0 && !(ngf < tol);
  f[F_OUTER] = on;
  f[F_TCG] = 0; f[F_STEP] = 0; f[F_END] = 0;
  f[F_STATUS] = DCORA_TCG_MAXITER;
  f[F_NINNER] = 0; f[F_NOUTER] = 0; f[F_ACCEPT] = 0;
  if (on) atomicAdd(counter, 1);
}

// start of tCG: partial 0 = <z, r>
__global__ void __launch_bounds__(128)
tcg_init_kernel(const int *list, const int *agent_chunk0, const double *partials, int *flags,
                double *scal) {
  __shared__ double sm[4];
  const int a = list ? list[blockIdx.x] : blockIdx.x;
  int *f = flags + a * NF;
  if (f[F_OUTER] == 0) { if (threadIdx.x == 0) f[F_TCG] = 0; return; }
  const double zr = agent_sum(agent_chunk0, partials, a, 0, sm);
  if (threadIdx.x != 0) return;
  double *s = scal + a * NS;
  s[S_ZR] = zr; s[S_DPD] = zr; s[S_EPE] = 0.0; s[S_EPD] = 0.0;
  s[S_NR0] = s[S_NGF];
  f[F_TCG] = 1; f[F_STEP] = 0; f[F_END] = 0;
  f[F_STATUS] = DCORA_TCG_MAXITER;
}

// partial 0 = <delta, H delta>
__global__ void __launch_bounds__(128)
tcg_stepA_kernel(const int *list, const int *agent_chunk0, const double *partials, int *flags,
                 double *scal) {
  __shared__ double sm[4];
  const int a = list ? list[blockIdx.x] : blockIdx.x;
  int *f = flags + a * NF;
  if (f[F_TCG] == 0) { if (threadIdx.x == 0) f[F_STEP] = 0; return; }
  const double d_Hd = agent_sum(agent_chunk0, partials, a, 0, sm);
  if (threadIdx.x != 0) return;
  double *s = scal + a * NS;
  const double z_r = s[S_ZR], d_Pd = s[S_DPD], e_Pd = s[S_EPD], e_Pe = s[S_EPE];
  const double Delta = s[S_DELTA];
  const double alpha = z_r / d_Hd;
  const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
  s[S_DHD] = d_Hd;
  f[F_STEP] = 1;
  if (d_Hd <= 0.0 || e_Pe_new >= Delta * Delta) {
    const double tau =
        (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd;
    s[S_CETA] = tau;
    s[S_CR] = 0.0;
    f[F_STATUS] = d_Hd <= 0.0 ? DCORA_TCG_NEGCURVTURE : DCORA_TCG_EXCREGION;
    f[F_END] = 1;
  } else {
    s[S_EPE] = e_Pe_new;
    s[S_ALPHA] = alpha;
    s[S_CETA] = alpha;
    s[S_CR] = alpha;
    f[F_END] = 0;
  }
}

// partial 0 = <r, r> after the update
__global__ void __launch_bounds__(128)
tcg_stepB_kernel(const int *list, const int *agent_chunk0, const double *partials, int *flags,
                 double *scal, int j, int min_inner, int max_inner, int *counter,
                 int *counter_next) {
  __shared__ double sm[4];
  if (blockIdx.x == 0 && threadIdx.x == 0) *counter_next = 0;
  const int a = list ? list[blockIdx.x] : blockIdx.x;
  int *f = flags + a * NF;
  if (f[F_STEP] == 0) return;
  const double rr = agent_sum(agent_chunk0, partials, a, 0, sm);
  if (threadIdx.x != 0) return;
  double *s = scal + a * NS;
  f[F_NINNER] += 1;
  int on = 1;
  if (f[F_END]) {
    on = 0;
  } else {
    const double nr = sqrt(rr), nr0 = s[S_NR0];
    s[S_NR] = nr;
    const double theta = 1.0, kappa = 0.1;
    const double pw = pow(nr0, theta);
    if (j >= min_inner && nr <= nr0 * (pw < kappa ? pw : kappa)) {
      f[F_STATUS] = kappa < pw ? DCORA_TCG_LCON : DCORA_TCG_SCON;
      on = 0;
    } else if (j + 1 >= max_inner) {
      on = 0;  // status stays TR_MAXITER
    }
  }
  f[F_TCG] = on;
  if (on) atomicAdd(counter, 1);
}

// partial 0 = <z, r> with the new z
__global__ void __launch_bounds__(128)
tcg_stepC_kernel(const int *list, const int *agent_chunk0, const double *partials, const int *flags,
                 double *scal) {
  __shared__ double sm[4];
  const int a = list ? list[blockIdx.x] : blockIdx.x;
  if (flags[a * NF + F_TCG] == 0) return;
  const double zr = agent_sum(agent_chunk0, partials, a, 0, sm);
  if (threadIdx.x != 0) return;
  double *s = scal + a * NS;
  const double zold = s[S_ZR];
  const double beta = zr / zold;
  const double alpha = s[S_ALPHA], d_Pd = s[S_DPD], e_Pd = s[S_EPD];
  s[S_ZR] = zr;
  s[S_BETA] = beta;
  s[S_EPD] = beta * (e_Pd + alpha * d_Pd);
  s[S_DPD] = zr + beta * beta * d_Pd;
}

// S_F2 holds f(x2); partial 0 = <eta, gf1>, partial 1 = <eta, H eta>
__global__ void __launch_bounds__(128)
rtr_rho_kernel(const int *list, const int *agent_chunk0, const double *partials, int *flags,
               double *scal, int single_mode, int stable) {
  __shared__ double sm[4];
  const int a = list ? list[blockIdx.x] : blockIdx.x;
  int *f = flags + a * NF;
  if (f[F_OUTER] == 0) { if (threadIdx.x == 0) f[F_ACCEPT] = 0; return; }
  // stable: slots 0, 1 = <d, EG>, <d, dQ> (d = x2 - x1), slots 2, 3 = <eta, gf>, <eta, H eta>
  const double eg = agent_sum(agent_chunk0, partials, a, stable ? 2 : 0, sm);
  const double eh = agent_sum(agent_chunk0, partials, a, stable ? 3 : 1, sm);
  double dEG = 0.0, dQd = 0.0;
  if (stable) {
    dEG = agent_sum(agent_chunk0, partials, a, 0, sm);
    dQd = agent_sum(agent_chunk0, partials, a, 1, sm);
  }
  if (threadIdx.x != 0) return;
  double *s = scal + a * NS;
  const double f1 = s[S_F1];
  // f(x2) - f(x1) = <egrad(x1), d> + 0.5 <d Q, d> for the quadratic cost: no cancellation
  // between two costs of size 1e9 when the decrease is 1e-6
  const double f2 = stable ? f1 + (dEG + 0.5 * dQd) : s[S_F2];
  if (stable) s[S_F2] = f2;
  const double rho = (f1 - f2) / (-(eg + 0.5 * eh));
  s[S_RHO] = rho;
  double Delta = s[S_DELTA];
  const int st = f[F_STATUS];
  if (rho > 0.75) {
    if (st == DCORA_TCG_EXCREGION || st == DCORA_TCG_NEGCURVTURE)
      Delta = fmin(2.0 * Delta, s[S_MAXDELTA]);
  } else if (rho < 0.25) {
    Delta *= 0.25;
  }
  const int acc = (rho > 0.1) ||
                  (fabs(f1 - f2) / (fabs(f1) + 1.0) < 1.4901161193847656e-08 && f2 < f1);
  if (single_mode) {
    // QuadraticOptimizer.cpp:254-273: radius /= 4 after a rejection and
    // maximum_Delta = initial_Delta = radius for the next Run()
    if (!acc) { Delta = s[S_MAXDELTA] * 0.25; s[S_MAXDELTA] = Delta; }
  }
  s[S_DELTA] = Delta;
  f[F_ACCEPT] = acc;
  f[F_NOUTER] += 1;
}

// after the gradient at the accepted point: partial 0 = <gf, gf>
__global__ void __launch_bounds__(128)
rtr_accept_kernel(const int *list, const int *agent_chunk0, const double *partials, int *flags,
                  double *scal, int max_iter, int single_mode, int *counter, int *counter_next) {
  __shared__ double sm[4];
  if (blockIdx.x == 0 && threadIdx.x == 0) *counter_next = 0;
  const int a = list ? list[blockIdx.x] : blockIdx.x;
  int *f = flags + a * NF;
  if (f[F_OUTER] == 0) return;
  double gg = 0.0;
  if (f[F_ACCEPT]) gg = agent_sum(agent_chunk0, partials, a, 0, sm);
  if (threadIdx.x != 0) return;
  double *s = scal + a * NS;
  int on = 1;
  if (f[F_ACCEPT]) {
    s[S_F1] = s[S_F2];
    s[S_NGF] = sqrt(gg);
    if (s[S_NGF] < s[S_TOL]) on = 0;
    if (single_mode) on = 0;
  }
  if (f[F_NOUTER] >= max_iter) on = 0;
  f[F_OUTER] = on;
  if (on) atomicAdd(counter, 1);
}

}  // namespace

int launch_combine_f(dcora_problem *p, int slot, int with_G, int mask_flag) {
  const ActiveSet *as = active_set(p, mask_flag);
  if (as && as->n_agents == 0) return DCORA_OK;
  combine_f_kernel<<<as ? as->n_agents : p->M->A, 128, 0, p->stream>>>(
      as ? as->agents : nullptr, p->M->agent_chunk0, p->partials, p->flags, mask_flag, slot,
      with_G, p->scal);
  DCORA_LAUNCH_CHECK(p);
  return DCORA_OK;
}

int eval_f(dcora_problem *p, const double *X, double *tmp, int slot, int mask_flag) {
  DCORA_TRY(launch_spmm(p, X, nullptr, tmp, mask_flag));
  const double *xs[2] = {tmp, X};
  const double *ys[2] = {X, p->G};
  DCORA_TRY(launch_dots(p, p->has_G ? 2 : 1, xs, ys, mask_flag));
  return launch_combine_f(p, slot, p->has_G ? 1 : 0, mask_flag);
}

// "How many agents go on?" counters.  The host must know them to stop launching, but it
// must not wait for them: every count lives in its own slot of a small ring, is copied to
// pinned memory behind the kernel that produces it, and is looked at one loop iteration
// LATER -- by then the copy has long completed while the GPU is still busy with the work
// queued in between.  The price is a handful of launches that find every agent switched
// off when a loop ends (they exit at once).
struct CountRing {
  dcora_problem *p;
  int *dev, *host;
  int next = 0;
  bool fresh = true;
  explicit CountRing(dcora_problem *p_)
      : p(p_), dev(p_->flags + (size_t)NF * p_->M->A + 4),
        host(p_->h_flags + (size_t)NF * p_->M->A + 4) {}
  // a zeroed slot for the next counting kernel: the first one is cleared here, every later
  // one by the counting kernel before it (*dnext; its previous read-back was enqueued
  // DCORA_COUNT_RING - 1 counting kernels earlier on the same stream)
  int open(int **dptr, int **dnext) {
    const int s = next;
    next = (next + 1) % DCORA_COUNT_RING;
    if (fresh) {
      if (cudaMemsetAsync(dev + s, 0, sizeof(int), p->stream) != cudaSuccess) return -1;
      fresh = false;
    }
    *dptr = dev + s;
    *dnext = dev + next;
    return s;
  }
  // after the counting kernel: start the read-back
  int post(int s) {
    if (!p->count_ev[s] &&
        cudaEventCreateWithFlags(&p->count_ev[s], cudaEventDisableTiming) != cudaSuccess)
      return DCORA_ERR_CUDA;
    if (cudaMemcpyAsync(host + s, dev + s, sizeof(int), cudaMemcpyDeviceToHost, p->stream) !=
            cudaSuccess ||
        cudaEventRecord(p->count_ev[s], p->stream) != cudaSuccess)
      return DCORA_ERR_CUDA;
    return DCORA_OK;
  }
  int get(int s, int *out) {
    if (cudaEventSynchronize(p->count_ev[s]) != cudaSuccess) return DCORA_ERR_CUDA;
    *out = host[s];
    return DCORA_OK;
  }
};

// Riemannian Hessian-vector product at x1: out = proj(V*Q - V*sym(Y^T EG))
// with_dot: also leaves <v, out> in partial slot 0
static int rhess(dcora_problem *p, const double *x1, const double *S, const double *v,
                 double *tmp, double *out, int mask, bool with_dot = false) {
  DCORA_TRY(launch_spmm(p, v, nullptr, tmp, mask));
  return launch_rhess_correct(p, x1, S, v, tmp, out, mask, with_dot);
}

// EG, gf, S at x1 for the agents selected by `mask`; leaves <gf,gf> in partial 0.
// `XQ` (optional) holds x1*Q from the cost evaluation at the same point
// (eval_f): the Euclidean gradient is then XQ + G, bit-identical to the fused
// SpMM epilogue, without a second pass over Q.
static int gradient_at(dcora_problem *p, const double *x1, double *EG, double *gf,
                       double *S, int mask, const double *XQ = nullptr) {
  if (XQ != nullptr)  // one fused pass: EG, gf, S and <gf, gf>
    return launch_gradient_fused(p, x1, XQ, p->has_G ? p->G : nullptr, EG, gf, S, mask);
  DCORA_TRY(launch_spmm(p, x1, p->has_G ? p->G : nullptr, EG, mask));
  DCORA_TRY(launch_project_tangent(p, x1, EG, gf, mask));
  DCORA_TRY(launch_sym_ytg(p, x1, EG, S, mask));
  const double *xs[1] = {gf};
  return launch_dots(p, 1, xs, xs, mask);
}

namespace {
__global__ void __launch_bounds__(128)
set_user_flags_kernel(int A, const uint8_t *__restrict__ mask, int *__restrict__ flags) {
  const int a = blockIdx.x * 128 + threadIdx.x;
  if (a >= A) return;
  flags[a * NF + F_USER] = mask ? (mask[a] ? 1 : 0) : 1;
  flags[a * NF + F_ALL] = 1;
}

template <typename T>
cudaError_t up_list(T **dst, const std::vector<T> &src) {
  cudaError_t e = cudaMalloc(dst, std::max<size_t>(1, src.size()) * sizeof(T));
  if (e != cudaSuccess || src.empty()) return e;
  return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}
}  // namespace

// flags[a][F_USER] <- active[a] (NULL: all agents) and the matching compacted
// launch lists (built once per distinct subset and cached: an RBCD sweep cycles
// through the same few colour classes).
int set_user_mask(dcora_problem *p, const uint8_t *active) {
  DeviceMatrices &M = *p->M;
  const int A = M.A;
  bool all = active == nullptr;
  if (!all) {
    all = true;
    for (int a = 0; a < A && all; ++a) all = active[a] != 0;
  }
  if (all) {
    p->act.reset();
  } else {
    const std::string key(reinterpret_cast<const char *>(active), (size_t)A);
    auto it = p->act_cache.find(key);
    const void *cpl = p->C.get();
    if (it == p->act_cache.end() || it->second->coupling_of != cpl) {
      auto as = std::make_shared<ActiveSet>();
      std::vector<int> agents, chunks, rem, crem;
      std::vector<int2> tiles;
      std::vector<uint8_t> mask(active, active + A);
      const int ppc = spmm_rows_per_cta(M.d);
      for (int a = 0; a < A; ++a) {
        if (!active[a]) continue;
        agents.push_back(a);
        for (int c = M.agent_chunk0_h[a]; c < M.agent_chunk0_h[a + 1]; ++c) chunks.push_back(c);
        for (int i0 = M.poff[a]; i0 < M.poff[a + 1]; i0 += ppc)
          tiles.push_back(make_int2(i0, std::min(i0 + ppc, M.poff[a + 1])));
      }
      for (size_t s = 0; s < M.rem_agent_h.size(); ++s)
        if (active[M.rem_agent_h[s]]) rem.push_back((int)s);
      if (p->C)
        for (size_t s = 0; s < p->C->rem_agent_h.size(); ++s)
          if (active[p->C->rem_agent_h[s]]) crem.push_back((int)s);
      as->n_agents = (int)agents.size(); as->n_chunks = (int)chunks.size();
      as->n_tiles = (int)tiles.size(); as->n_rem = (int)rem.size(); as->n_crem = (int)crem.size();
      as->coupling_of = cpl;
      DCORA_CUDA(up_list(&as->agents, agents));
      DCORA_CUDA(up_list(&as->chunks, chunks));
      DCORA_CUDA(up_list(&as->tiles, tiles));
      DCORA_CUDA(up_list(&as->rem_rows, rem));
      DCORA_CUDA(up_list(&as->crem_rows, crem));
      DCORA_CUDA(up_list(&as->mask, mask));
      if (p->act_cache.size() > 64) p->act_cache.clear();
      p->act_cache[key] = as;
      p->act = as;
    } else {
      p->act = it->second;
    }
  }
  set_user_flags_kernel<<<(A + 127) / 128, 128, 0, p->stream>>>(
      A, p->act ? p->act->mask : nullptr, p->flags);
  DCORA_LAUNCH_CHECK(p);
  return DCORA_OK;
}

namespace {
// DCORA_TIMING=1: wall-clock split of one rtr_solve into its phases (adds a
// stream synchronise per mark: profiling aid only, never on in benchmarks).
struct PhaseTimer {
  bool on;
  cudaStream_t st;
  std::chrono::steady_clock::time_point t0;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  explicit PhaseTimer(cudaStream_t s) : on(getenv("DCORA_TIMING") != nullptr), st(s) {
    if (on) { cudaStreamSynchronize(st); t0 = std::chrono::steady_clock::now(); }
  }
  void mark(int slot) {
    if (!on) return;
    cudaStreamSynchronize(st);
    const auto t1 = std::chrono::steady_clock::now();
    acc[slot] += std::chrono::duration<double, std::micro>(t1 - t0).count();
    t0 = t1;
  }
  void report(int outer, int inner) {
    if (!on) return;
    fprintf(stderr, "[dcora timing] init %.0f | tcg-setup(precond) %.0f | tcg-iter(spmm..stepB) %.0f | "
                    "tcg-iter(precond..) %.0f | candidate+rho %.0f | accept+grad %.0f | readback %.0f us"
                    "  (outer %d, tcg %d)\n",
            acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6], outer, inner);
  }
};
}  // namespace

int rtr_solve_device(dcora_problem *p, double *x1, const dcora_ropt_params *prm,
                     const uint8_t *active, dcora_ropt_result *results) {
  DeviceMatrices &M = *p->M;
  PhaseTimer pt(p->stream);
  int n_outer_done = 0, n_inner_done = 0;
  const int A = M.A;
  double *x2 = p->vec(WS_X2), *EG = p->vec(WS_EG), *gf = p->vec(WS_GF);
  double *S = p->vec(WS_S), *eta = p->vec(WS_ETA), *delta = p->vec(WS_DELTA);
  double *Hd = p->vec(WS_HD), *r = p->vec(WS_R), *z = p->vec(WS_Z);
  double *t0 = p->vec(WS_T0), *xq = p->vec(WS_T1), *Heta = p->vec(WS_HETA);
  if (!x2 || !EG || !gf || !S || !eta || !delta || !Hd || !r || !z || !t0 || !xq || !Heta)
    return DCORA_ERR_ALLOC;
  const auto tstart = std::chrono::steady_clock::now();
  DCORA_TRY(set_user_mask(p, active));
  // per-agent control kernels run over the caller's subset only
  const int nA = p->act ? p->act->n_agents : A;
  const int *alist = p->act ? p->act->agents : nullptr;
  if (nA == 0) {
    DCORA_TRY(read_scalars(p));
    if (results)
      for (int a = 0; a < A; ++a) { results[a] = dcora_ropt_result(); }
    return DCORA_OK;
  }

  // initial statistics: fInit, gradNormInit (QuadraticOptimizer.cpp:30-31)
  DCORA_TRY(eval_f(p, x1, xq, S_F1, F_USER));
  DCORA_TRY(gradient_at(p, x1, EG, gf, S, F_USER, xq));
  CountRing ring(p);
  const int single_mode = prm->RTR_iterations == 1;
  const int max_outer = single_mode ? 12 : prm->RTR_iterations;
  const double Delta0 = prm->RTR_initial_radius;
  int *dcnt = nullptr, *dnext = nullptr;
  int outer_slot = ring.open(&dcnt, &dnext);
  if (outer_slot < 0) return DCORA_ERR_CUDA;
  rtr_init_kernel<<<nA, 128, 0, p->stream>>>(alist, M.agent_chunk0, p->partials, p->flags,
                                            p->scal, prm->gradnorm_tol, Delta0,
                                            single_mode ? Delta0 : 5.0 * Delta0, dcnt, dnext);
  DCORA_LAUNCH_CHECK(p);
  DCORA_TRY(ring.post(outer_slot));
  pt.mark(0);

  if (prm->method == 1) {
    // one Riemannian gradient-descent step (QuadraticOptimizer.cpp:124-201)
    const double *g = gf;
    if (prm->RGD_use_preconditioner) {
      DCORA_TRY(launch_precond(p, x1, gf, z, F_USER));
      g = z;
    }
    DCORA_TRY(launch_axpby(p, eta, -1, -prm->RGD_stepsize, g, -1, 0.0, nullptr, F_USER));
    DCORA_TRY(launch_retract(p, x1, eta, x2, prm->retraction, F_USER));
    DCORA_TRY(launch_copy(p, x1, x2, F_USER));
  }

  for (int it = 0; it < max_outer && prm->method != 1; ++it) {
    // ---- tCG (SolversTR::tCG_TR), eta0 = 0
    // r = gf, eta = 0, Heta = 0.  H eta is accumulated with eta (H eta += c H delta, the
    // Hessian is linear on the tangent space), which saves the SpMM of a separate H*eta
    // for the model decrease
    DCORA_TRY(launch_tcg_begin(p, gf, r, eta, Heta, F_OUTER));
    DCORA_TRY(launch_precond(p, x1, r, z, F_OUTER, nullptr, 0, r));  // + <z, r>
    tcg_init_kernel<<<nA, 128, 0, p->stream>>>(alist, M.agent_chunk0, p->partials, p->flags, p->scal);
    DCORA_LAUNCH_CHECK(p);
    DCORA_TRY(launch_axpby(p, delta, -1, -1.0, z, -1, 0.0, nullptr, F_TCG));
    // how many agents entered this outer iteration?  (counted by the kernel before the
    // four launches above: if none did, they found every agent switched off)
    int n_outer = 0;
    DCORA_TRY(ring.get(outer_slot, &n_outer));
    if (n_outer == 0) break;
    pt.mark(1);
    ++n_outer_done;
    int prev_slot = -1;
    for (int j = 0; j < prm->RTR_tCG_iterations; ++j) {
      DCORA_TRY(rhess(p, x1, S, delta, t0, Hd, F_TCG, true));  // + <delta, H delta>
      tcg_stepA_kernel<<<nA, 128, 0, p->stream>>>(alist, M.agent_chunk0, p->partials, p->flags, p->scal);
      DCORA_LAUNCH_CHECK(p);
      // eta += c delta, Heta += c Hd, r += c_r Hd and <r, r> in one pass
      DCORA_TRY(launch_tcg_update(p, S_CETA, S_CR, delta, Hd, eta, Heta, r, F_STEP));
      const int slot = ring.open(&dcnt, &dnext);
      if (slot < 0) return DCORA_ERR_CUDA;
      tcg_stepB_kernel<<<nA, 128, 0, p->stream>>>(alist, M.agent_chunk0, p->partials, p->flags,
                                                 p->scal, j, 0, prm->RTR_tCG_iterations, dcnt, dnext);
      DCORA_LAUNCH_CHECK(p);
      DCORA_TRY(ring.post(slot));
      pt.mark(2);
      // the PREVIOUS iteration's count (this iteration's is still in flight): zero means
      // the tCG of every agent had ended and the launches above exited at once
      if (prev_slot >= 0) {
        int n_tcg = 0;
        DCORA_TRY(ring.get(prev_slot, &n_tcg));
        if (n_tcg == 0) break;
      }
      ++n_inner_done;
      prev_slot = slot;
      DCORA_TRY(launch_precond(p, x1, r, z, F_TCG, nullptr, 0, r));  // + <z, r>
      tcg_stepC_kernel<<<nA, 128, 0, p->stream>>>(alist, M.agent_chunk0, p->partials, p->flags, p->scal);
      DCORA_LAUNCH_CHECK(p);
      // delta = -z + beta * delta
      DCORA_TRY(launch_axpby(p, delta, -1, -1.0, z, S_BETA, 1.0, delta, F_TCG));
      pt.mark(3);
    }
    // ---- candidate, model decrease, rho (SolversTR::Run)
    DCORA_TRY(launch_retract(p, x1, eta, x2, prm->retraction, F_OUTER));
    DCORA_TRY(eval_f(p, x2, xq, S_F2, F_OUTER));  // xq = x2*Q, reused by the gradient
    if (prm->stable_decrease) {
      // d = x2 - x1 (t0), dQ = x2 Q - x1 Q = xq - (EG - G) (Hd); both vectors are free here
      DCORA_TRY(launch_axpby(p, t0, -1, 1.0, x2, -1, -1.0, x1, F_OUTER));
      DCORA_TRY(launch_axpby(p, Hd, -1, 1.0, xq, -1, -1.0, EG, F_OUTER));
      if (p->has_G) DCORA_TRY(launch_axpby(p, Hd, -1, 1.0, Hd, -1, 1.0, p->G, F_OUTER));
      const double *xs[4] = {t0, t0, eta, eta}, *ys[4] = {EG, Hd, gf, Heta};
      DCORA_TRY(launch_dots(p, 4, xs, ys, F_OUTER));
    } else {
      const double *xs[2] = {eta, eta}, *ys[2] = {gf, Heta};
      DCORA_TRY(launch_dots(p, 2, xs, ys, F_OUTER));
    }
    rtr_rho_kernel<<<nA, 128, 0, p->stream>>>(alist, M.agent_chunk0, p->partials, p->flags, p->scal,
                                             single_mode, prm->stable_decrease ? 1 : 0);
    DCORA_LAUNCH_CHECK(p);
    pt.mark(4);
    DCORA_TRY(launch_copy(p, x1, x2, F_ACCEPT));
    DCORA_TRY(gradient_at(p, x1, EG, gf, S, F_ACCEPT, xq));
    outer_slot = ring.open(&dcnt, &dnext);
    if (outer_slot < 0) return DCORA_ERR_CUDA;
    rtr_accept_kernel<<<nA, 128, 0, p->stream>>>(alist, M.agent_chunk0, p->partials, p->flags, p->scal,
                                                max_outer, single_mode, dcnt, dnext);
    DCORA_LAUNCH_CHECK(p);
    DCORA_TRY(ring.post(outer_slot));
    pt.mark(5);
    if (prm->time_bound_s > 0.0) {
      // the reference's wall-clock bound is checked against finished work
      DCORA_CUDA(cudaStreamSynchronize(p->stream));
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - tstart).count();
      if (el > prm->time_bound_s) break;
    }
  }

  if (prm->method == 1) {
    // statistics after the RGD step
    DCORA_TRY(eval_f(p, x1, xq, S_F1, F_USER));
    DCORA_TRY(gradient_at(p, x1, EG, gf, S, F_USER, xq));
    const int dst[1] = {S_TMP0};
    DCORA_TRY(launch_reduce_scalars(p, 1, dst));
  }
  DCORA_TRY(read_scalars(p));
  pt.mark(6);
  pt.report(n_outer_done, n_inner_done);
  const double ms =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tstart).count();
  if (results) {
    for (int a = 0; a < A; ++a) {
      const double *s = p->h_scal + a * NS;
      const int *f = p->h_flags + a * NF;
      dcora_ropt_result &R = results[a];
      R.success = f[F_USER] ? 1 : 0;
      R.tCGStatus = f[F_STATUS];
      R.fInit = s[S_FINIT];
      R.gradNormInit = s[S_GNINIT];
      R.fOpt = s[S_F1];
      R.gradNormOpt = prm->method == 1 ? std::sqrt(s[S_TMP0]) : s[S_NGF];
      R.elapsedMs = ms;
      R.outer_iterations = f[F_NOUTER];
      R.inner_iterations = f[F_NINNER];
      R.accepted = f[F_ACCEPT];
      R.reserved = 0;
    }
  }
  return DCORA_OK;
}

}  // namespace dcora

using namespace dcora;

extern "C" int dcora_rtr_solve(dcora_problem *p, const double *X0, double *Xopt,
                               const dcora_ropt_params *params, const uint8_t *active,
                               dcora_ropt_result *results) {
  if (!p || !X0 || !Xopt) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  dcora_ropt_params prm;
  if (params) prm = *params; else dcora_ropt_params_default(&prm);
  double *x1 = nullptr;
  const bool track = p->tracking && p->xe && pointer_on_device(X0) && pointer_on_device(Xopt);
  // in place on the tracked buffer: only the updated agents' columns change
  const bool in_place_tracked = track && X0 == Xopt && X0 == p->tracked;
  if (track) {
    DCORA_TRY(acquire_iterate(p, X0, nullptr, &x1));   // the iterate lives in p->xe
  } else {
    x1 = p->vec(WS_X1);
    if (!x1) return DCORA_ERR_ALLOC;
    DCORA_TRY(import_matrix(p, X0, x1, p->r));
  }
  const int rc_solve = rtr_cluster_applicable(p, &prm)
                           ? rtr_solve_cluster(p, x1, &prm, active, results)
                           : rtr_solve_device(p, x1, &prm, active, results);
  if (track) p->tracked = nullptr;   // xe changed; re-established by the write-back below
  if (rc_solve < 0) return rc_solve;
  if (in_place_tracked) {
    DCORA_TRY(export_active_columns(p, x1, Xopt, F_USER));
    DCORA_CUDA(cudaStreamSynchronize(p->stream));
    p->tracked = Xopt;
    return DCORA_OK;
  }
  const int rc_exp = export_matrix(p, x1, Xopt, p->r);
  if (rc_exp >= 0 && track) p->tracked = Xopt;
  return rc_exp;
}
