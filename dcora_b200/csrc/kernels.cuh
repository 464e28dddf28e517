// Launch wrappers of the sm_100a kernels (kernels.cu).  All pointers are
// device pointers in the INTERNAL layout (see layout.cpp); all launches go to
// p->stream and bump p->launches.
#pragma once
#include "common.cuh"

namespace dcora {

constexpr int NPART = 4;  // dot-product slots per chunk in p->partials

// per-agent scalar slots (p->scal[a*NS + i])
enum Scal : int {
  S_F1 = 0, S_F2, S_NGF, S_DELTA, S_RHO, S_ZR, S_DPD, S_EPD, S_EPE, S_ALPHA,
  S_BETA, S_NR0, S_NR, S_DHD, S_CETA, S_CR, S_FINIT, S_GNINIT, S_TMP0, S_TMP1,
  S_TMP2, S_TMP3, S_MAXDELTA, S_TOL, NS
};
// per-agent int flags (p->flags[a*NF + i])
enum Flag : int {
  F_OUTER = 0,   // agent still iterating the RTR outer loop
  F_TCG,         // agent's tCG still running
  F_STEP,        // agent takes the eta/r update of this tCG iteration
  F_STATUS,      // tCG status
  F_NINNER, F_NOUTER, F_ACCEPT, F_ALL /* always 1 */, F_USER /* caller mask */,
  F_END,         // tCG hit the boundary / negative curvature this iteration
  NF
};

// A block-sparse operand of the SpMM kernels: BSR over output pose blocks +
// scalar CSR remainder (see layout.cpp).
struct SpmmMat {
  bool coupling = false;  // the cross-agent coupling matrix (rbcd.cu) rather than Q
  int np = 0;
  const int *browptr = nullptr, *bcol = nullptr, *prowcol = nullptr, *pose_agent = nullptr;
  const double *bval = nullptr;
  int64_t nrem_rows = 0;
  const int *rem_row = nullptr, *rem_rowptr = nullptr, *rem_col = nullptr, *rem_agent = nullptr;
  const uint8_t *rem_assign = nullptr;
  const double *rem_val = nullptr;
};
// Launch lists for a masked launch: the caller's agent subset when the mask is a
// per-solve flag (every such flag is a subset of F_USER), null for the whole
// batch.
inline const ActiveSet *active_set(const dcora_problem *p, int mask_flag) {
  return (mask_flag >= 0 && mask_flag != F_ALL) ? p->act.get() : nullptr;
}
int launch_spmm_mat(dcora_problem *p, const SpmmMat &A, const double *X,
                    const double *G, double *Y, int mask_flag);
// Y = X*Q (+G).  mask_flag < 0: all agents.
int launch_spmm(dcora_problem *p, const double *X, const double *G, double *Y,
                int mask_flag);
// Manifold ops (per chunk; mask as above)
int launch_project_tangent(dcora_problem *p, const double *X, const double *V,
                           double *OUT, int mask_flag, const double *dot_with = nullptr);
// OUT = proj_X(HV - V*sym(Y^T EG)) with S = sym(Y^T EG) precomputed
int launch_rhess_correct(dcora_problem *p, const double *X, const double *S,
                         const double *V, const double *HV, double *OUT,
                         int mask_flag, bool with_dot = false);
int launch_gradient_fused(dcora_problem *p, const double *X, const double *XQ,
                          const double *G, double *EG, double *gf, double *S,
                          int mask_flag);
int launch_tcg_begin(dcora_problem *p, const double *gf, double *r, double *eta, double *Heta,
                     int mask_flag);
int launch_tcg_update(dcora_problem *p, int slot_eta, int slot_r, const double *delta,
                      const double *Hd, double *eta, double *Heta, double *r, int mask_flag);
// S[pose] = sym(Y^T EG) (d*d per pose block, 1 per sphere column): stored
// column-aligned in a vector of the same shape (first d*d / 1 entries used)
int launch_sym_ytg(dcora_problem *p, const double *X, const double *EG, double *S,
                   int mask_flag);
int launch_retract(dcora_problem *p, const double *X, const double *V, double *OUT,
                   int kind, int mask_flag);
int launch_project_manifold(dcora_problem *p, const double *M, double *OUT,
                            int mask_flag);
// partial dot products: slot s of chunk c <- <x_s, y_s> over the chunk (n<=NPART)
int launch_dots(dcora_problem *p, int n, const double *const *x,
                const double *const *y, int mask_flag);
// out = a*x + b*y with per-agent coefficients read from scal slots (or
// constants when slot < 0 -> uses ca / cb); y may be null (b ignored).
int launch_axpby(dcora_problem *p, double *out, int slot_a, double ca,
                 const double *x, int slot_b, double cb, const double *y,
                 int mask_flag);
int launch_copy(dcora_problem *p, double *out, const double *x, int mask_flag);
// layout permutation between reference and internal column order
int launch_permute(dcora_problem *p, const double *in, double *out, int r,
                   bool to_internal);
int launch_permute_cols(dcora_problem *p, const int *perm, int64_t ncols,
                        const double *in, double *out, int r, bool to_internal);
int launch_reduce_scalars(dcora_problem *p, int nslots, const int *dst_slots);

// preconditioner application z = proj_X( (Q+reg I)^-1 v ) (sptrsv.cu)
int launch_precond(dcora_problem *p, const double *X, const double *V, double *OUT,
                   int mask_flag, unsigned long long *trace = nullptr, int trace_agent = 0,
                   const double *dot_with = nullptr);
struct AgentTape;
int upload_factor(dcora::DeviceMatrices *M, const std::vector<HostFactor> &F,
                  std::vector<AgentTape> &tapes, int nwarps, const HostLayout &H);
int trsv_warps_for(double rows_per_agent);  // warps one agent's solve is dealt to

// workspace vector indices (p->vec(i))
enum Ws : int {
  WS_X1 = 0, WS_X2, WS_EG, WS_GF, WS_S, WS_ETA, WS_DELTA, WS_HD, WS_R, WS_Z,
  WS_T0, WS_T1, WS_HETA, WS_TRSV_PANEL, WS_TRSV_TMP, WS_TRSV_SOL, WS_IO0, WS_IO1, WS_IO2, WS_COUNT
};

}  // namespace dcora

#define DCORA_DISPATCH_R(r, ...)                           \
  switch (r) {                                             \
    case 1: { constexpr int R = 1; __VA_ARGS__ } break;           \
    case 2: { constexpr int R = 2; __VA_ARGS__ } break;           \
    case 3: { constexpr int R = 3; __VA_ARGS__ } break;           \
    case 4: { constexpr int R = 4; __VA_ARGS__ } break;           \
    case 5: { constexpr int R = 5; __VA_ARGS__ } break;           \
    case 6: { constexpr int R = 6; __VA_ARGS__ } break;           \
    case 7: { constexpr int R = 7; __VA_ARGS__ } break;           \
    case 8: { constexpr int R = 8; __VA_ARGS__ } break;           \
    case 9: { constexpr int R = 9; __VA_ARGS__ } break;           \
    case 10: { constexpr int R = 10; __VA_ARGS__ } break;         \
    case 11: { constexpr int R = 11; __VA_ARGS__ } break;         \
    case 12: { constexpr int R = 12; __VA_ARGS__ } break;         \
    case 13: { constexpr int R = 13; __VA_ARGS__ } break;         \
    case 14: { constexpr int R = 14; __VA_ARGS__ } break;         \
    case 15: { constexpr int R = 15; __VA_ARGS__ } break;         \
    case 16: { constexpr int R = 16; __VA_ARGS__ } break;         \
    default:                                               \
      set_error("rank outside 1..16");                     \
      return DCORA_ERR_ARG;                                \
  }

// K1: pose rows per CTA (the active-set tiles are built to match)
inline int spmm_rows_per_cta(int d) { return 4 * (32 / (d + 1)); }

#define DCORA_DISPATCH_RD(r, d, ...)                       \
  if ((d) == 3) { constexpr int DH = 4; DCORA_DISPATCH_R(r, __VA_ARGS__) } \
  else { constexpr int DH = 3; DCORA_DISPATCH_R(r, __VA_ARGS__) }

#define DCORA_LAUNCH_CHECK(p)                              \
  do {                                                     \
    (p)->launches++;                                       \
    DCORA_CUDA(cudaGetLastError());                        \
  } while (0)

