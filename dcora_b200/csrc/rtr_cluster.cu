// Host side of the per-agent RTR cluster kernel (rtr_cluster.cuh): when it applies, argument
// set-up, dispatch to the per-rank translation units (rtr_cluster_inst.cu), result read-back.
#include <chrono>
#include <cstdlib>

#include "rtr_cluster.cuh"

namespace dcora {

int launch_cluster_r2(dcora_problem *, const ClusterArgs &, int, int);
int launch_cluster_r3(dcora_problem *, const ClusterArgs &, int, int);
int launch_cluster_r4(dcora_problem *, const ClusterArgs &, int, int);
int launch_cluster_r5(dcora_problem *, const ClusterArgs &, int, int);
int launch_cluster_r6(dcora_problem *, const ClusterArgs &, int, int);
int launch_cluster_r7(dcora_problem *, const ClusterArgs &, int, int);
int launch_cluster_r8(dcora_problem *, const ClusterArgs &, int, int);

// The one-launch solve is used for agents that fit one CTA (< 1500 columns: 8 or 16 warps).
bool rtr_cluster_applicable(const dcora_problem *p, const dcora_ropt_params *prm) {
  static const char *force = getenv("DCORA_RTR_PATH");   // "host": never
  if (prm->method != 0 || prm->stable_decrease) return false;
  if (p->r < p->M->d || p->r < CLUSTER_MIN_RANK || p->r > CLUSTER_MAX_RANK) return false;
  if (force && force[0] == 'h') return false;
  const DeviceMatrices &M = *p->M;
  if (M.F.present) return M.F.nwarps <= 16;
  int64_t kmax = 0;
  for (int a = 0; a < M.A; ++a) kmax = std::max<int64_t>(kmax, M.coff[a + 1] - M.coff[a]);
  return trsv_warps_for((double)kmax) <= 16;
}

int rtr_solve_cluster(dcora_problem *p, double *x1, const dcora_ropt_params *prm,
                      const uint8_t *active, dcora_ropt_result *results) {
  DeviceMatrices &M = *p->M;
  const int A = M.A;
  const auto tstart = std::chrono::steady_clock::now();
  DCORA_TRY(set_user_mask(p, active));
  const int nA = p->act ? p->act->n_agents : A;
  if (nA > 0) {
    ClusterArgs a = {};
    a.browptr = M.browptr; a.bcol = M.bcol; a.prowcol = M.prowcol; a.bval = M.bval;
    a.rem_row = M.rem_row; a.rem_rowptr = M.rem_rowptr; a.rem_col = M.rem_col;
    a.rem_assign = M.rem_assign; a.rem_val = M.rem_val;
    a.chunks = M.chunks; a.agent_chunk0 = M.agent_chunk0; a.chunk_pose0 = M.chunk_pose0;
    a.chunk_rem = M.chunk_rem;
    a.has_factor = M.F.present ? 1 : 0;
    a.tape = M.F.tape; a.tape_off = M.F.tape_off; a.tape_nrec = M.F.tape_chunks;
    a.arow0 = M.F.arow0; a.frow_col = M.F.frow_col; a.col_frow = M.F.col_frow;
    a.agent_list = p->act ? p->act->agents : nullptr;
    a.flags = p->flags; a.scal = p->scal; a.partials = p->partials;
    a.x1 = x1; a.x2 = p->vec(WS_X2); a.EG = p->vec(WS_EG); a.gf = p->vec(WS_GF);
    a.S = p->vec(WS_S); a.eta = p->vec(WS_ETA); a.delta = p->vec(WS_DELTA);
    a.Hd = p->vec(WS_HD); a.r = p->vec(WS_R); a.z = p->vec(WS_Z); a.t0 = p->vec(WS_T0);
    a.xq = p->vec(WS_T1); a.Heta = p->vec(WS_HETA);
    a.T = p->vec(WS_TRSV_TMP); a.Xp = p->vec(WS_TRSV_PANEL);
    if (!a.x2 || !a.EG || !a.gf || !a.S || !a.eta || !a.delta || !a.Hd || !a.r || !a.z ||
        !a.t0 || !a.xq || !a.Heta || !a.T || !a.Xp)
      return DCORA_ERR_ALLOC;
    a.G = p->has_G ? p->G : nullptr;
    a.single_mode = prm->RTR_iterations == 1;
    a.max_outer = a.single_mode ? 12 : prm->RTR_iterations;
    a.max_inner = prm->RTR_tCG_iterations;
    a.retraction = prm->retraction;
    a.tol = prm->gradnorm_tol;
    a.delta0 = prm->RTR_initial_radius;
    a.max_delta = a.single_mode ? a.delta0 : 5.0 * a.delta0;
    a.time_bound_ns = prm->time_bound_s > 0.0 ? (unsigned long long)(prm->time_bound_s * 1e9) : 0ull;
    int nw = M.F.present ? M.F.nwarps : 8;
    if (!M.F.present) {
      int64_t kmax = 0;
      for (int ag = 0; ag < A; ++ag) kmax = std::max<int64_t>(kmax, M.coff[ag + 1] - M.coff[ag]);
      nw = trsv_warps_for((double)kmax);
    }
    int rc = DCORA_ERR_ARG;
    switch (p->r) {
      case 2: rc = launch_cluster_r2(p, a, nA, nw); break;
      case 3: rc = launch_cluster_r3(p, a, nA, nw); break;
      case 4: rc = launch_cluster_r4(p, a, nA, nw); break;
      case 5: rc = launch_cluster_r5(p, a, nA, nw); break;
      case 6: rc = launch_cluster_r6(p, a, nA, nw); break;
      case 7: rc = launch_cluster_r7(p, a, nA, nw); break;
      case 8: rc = launch_cluster_r8(p, a, nA, nw); break;
      default: set_error("rtr_solve_cluster: rank outside the built range");
    }
    if (rc < 0) return rc;
  }
  DCORA_TRY(read_scalars(p));
  const double ms =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tstart).count();
  if (results) {
    for (int ag = 0; ag < A; ++ag) {
      const double *s = p->h_scal + ag * NS;
      const int *f = p->h_flags + ag * NF;
      dcora_ropt_result &Rr = results[ag];
      const bool on = f[F_USER] != 0 && nA > 0;
      Rr = dcora_ropt_result();
      Rr.success = on ? 1 : 0;
      if (!on) continue;
      Rr.tCGStatus = f[F_STATUS];
      Rr.fInit = s[S_FINIT];
      Rr.gradNormInit = s[S_GNINIT];
      Rr.fOpt = s[S_F1];
      Rr.gradNormOpt = s[S_NGF];
      Rr.outer_iterations = f[F_NOUTER];
      Rr.inner_iterations = f[F_NINNER];
      Rr.accepted = f[F_ACCEPT];
      Rr.elapsedMs = ms;
    }
  }
  return DCORA_OK;
}

}  // namespace dcora
