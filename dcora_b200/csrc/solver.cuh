// RTR / tCG driver and host<->device staging helpers shared between
// problem.cu, solver.cu and certify.cu.
#pragma once
#include "kernels.cuh"

namespace dcora {

int import_matrix(dcora_problem *p, const double *src, double *dst, int r);
int export_matrix(dcora_problem *p, const double *src, double *dst, int r);
int read_scalars(dcora_problem *p);

// scal[slot] = 0.5*partial0 (+ partial1 if with_G) per agent
int launch_combine_f(dcora_problem *p, int slot, int with_G, int mask_flag);

// f(X) per agent into scal slot; tmp receives X*Q
int eval_f(dcora_problem *p, const double *X, double *tmp, int slot, int mask_flag);

// flags[a][F_USER] <- active[a] (NULL: all agents)
int set_user_mask(dcora_problem *p, const uint8_t *active);

}  // namespace dcora
