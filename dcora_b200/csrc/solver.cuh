// RTR / tCG driver and host<->device staging helpers shared between
// problem.cu, solver.cu and certify.cu.
#pragma once
#include "kernels.cuh"

namespace dcora {

int import_matrix(dcora_problem *p, const double *src, double *dst, int r);
int export_matrix(dcora_problem *p, const double *src, double *dst, int r);
int read_scalars(dcora_problem *p);
bool pointer_on_device(const void *ptr);
// Internal-layout view of the caller's iterate X (r x K): the tracked copy in
// p->xe when X is the tracked device buffer, else X imported into `buf` (or into
// p->xe, which then becomes the tracked copy, when tracking is on).
int acquire_iterate(dcora_problem *p, const double *X, double *buf, double **out);
// Reference-layout write-back of the active agents' columns only
int export_active_columns(dcora_problem *p, const double *src, double *dst, int mask_flag);

// scal[slot] = 0.5*partial0 (+ partial1 if with_G) per agent
int launch_combine_f(dcora_problem *p, int slot, int with_G, int mask_flag);

// f(X) per agent into scal slot; tmp receives X*Q
int eval_f(dcora_problem *p, const double *X, double *tmp, int slot, int mask_flag);

// flags[a][F_USER] <- active[a] (NULL: all agents)
int set_user_mask(dcora_problem *p, const uint8_t *active);

// RTR / tCG of the selected agents (x1: internal layout, updated in place):
// per-operation launches over the whole GPU (solver.cu) ...
int rtr_solve_device(dcora_problem *p, double *x1, const dcora_ropt_params *prm,
                     const uint8_t *active, dcora_ropt_result *results);
// ... or one launch, one cluster per agent (rtr_cluster.cu)
int rtr_solve_cluster(dcora_problem *p, double *x1, const dcora_ropt_params *prm,
                      const uint8_t *active, dcora_ropt_result *results);
bool rtr_cluster_applicable(const dcora_problem *p, const dcora_ropt_params *prm);

}  // namespace dcora
