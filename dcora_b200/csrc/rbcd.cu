// Multi-agent RBCD plumbing on the device: the cross-agent coupling product
// that rebuilds every agent's linear term G from its neighbours' public
// states (K8), the public-state gather (pack), per-agent cost terms, and the
// Nesterov combine+project step (K9).
//
// Replaces Graph::constructLinearCostTerm{PGO,RASLAM} (src/Graph.cpp:685-822,
// 1190-1772), Agent::getSharedStateDicts / updateNeighborStates
// (src/Agent.cpp:113-152,844-906) and Agent::updateY / updateV (:1202-1214).
#include <cstring>

#include "solver.cuh"

namespace dcora {

DeviceCoupling::~DeviceCoupling() {
  cudaSetDevice(device);
  cudaFree(perm_h); cudaFree(browptr); cudaFree(bcol); cudaFree(bval);
  cudaFree(rem_row); cudaFree(rem_rowptr); cudaFree(rem_col); cudaFree(rem_agent);
  cudaFree(rem_assign); cudaFree(rem_val);
}

namespace {

template <typename T>
cudaError_t upload(T **dst, const std::vector<T> &src) {
  cudaError_t e = cudaMalloc(dst, std::max<size_t>(1, src.size()) * sizeof(T));
  if (e != cudaSuccess) return e;
  if (src.empty()) return cudaSuccess;
  return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

bool is_device_ptr(const void *ptr) {
  if (!ptr) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// out[j*R + q] = X[idx[j]*R + q]
__global__ void __launch_bounds__(256)
gather_cols_kernel(int64_t count, int R, const int *__restrict__ idx,
                   const double *__restrict__ X, double *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= count * R) return;
  const int64_t j = t / R;
  const int q = (int)(t - j * R);
  out[t] = X[(int64_t)idx[j] * R + q];
}

}  // namespace
}  // namespace dcora

using namespace dcora;

extern "C" {

int dcora_problem_set_coupling(dcora_problem *p, const dcora_halo_segment *segs,
                               int32_t nseg, const int32_t *rowptr,
                               const int32_t *col, const double *val) {
  return dcora_problem_set_coupling_keyed(p, segs, nseg, rowptr, col, val, nullptr);
}

int dcora_problem_set_coupling_keyed(dcora_problem *p, const dcora_halo_segment *segs,
                                     int32_t nseg, const int32_t *rowptr,
                                     const int32_t *col, const double *val,
                                     const int64_t *row_key) {
  if (!p || !rowptr || nseg < 0 || (nseg > 0 && !segs)) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  DeviceMatrices &M = *p->M;
  int64_t KH = 0;
  for (int s = 0; s < nseg; ++s)
    KH += (int64_t)(M.d + 1) * segs[s].n + segs[s].l + segs[s].b;
  const int64_t KE = M.K + KH;
  std::vector<int64_t> rp(KE + 1);
  for (int64_t i = 0; i <= KE; ++i) rp[i] = rowptr[i];
  HostCoupling H;
  DCORA_TRY(build_coupling_layout(M.d, M.order, M.coff, M.n, M.l, M.poff, segs, nseg,
                                  rp.data(), col, val, row_key, &H));
  auto C = std::make_shared<DeviceCoupling>();
  C->device = M.device;
  C->KH = H.KH;
  C->identity_perm = H.identity_perm;
  C->nnzb = (int64_t)H.bcol.size();
  C->nrem_rows = (int64_t)H.rem_row.size();
  C->nnz_in = rp[KE];
  std::vector<int> rem_agent(H.rem_row.size());
  {
    int a = 0;
    for (size_t s = 0; s < H.rem_row.size(); ++s) {
      while (H.rem_row[s] >= M.coff[a + 1]) ++a;
      rem_agent[s] = a;
    }
  }
  DCORA_CUDA(upload(&C->perm_h, H.perm_h));
  DCORA_CUDA(upload(&C->browptr, H.browptr));
  DCORA_CUDA(upload(&C->bcol, H.bcol));
  DCORA_CUDA(upload(&C->bval, H.bval));
  DCORA_CUDA(upload(&C->rem_row, H.rem_row));
  DCORA_CUDA(upload(&C->rem_rowptr, H.rem_rowptr));
  DCORA_CUDA(upload(&C->rem_col, H.rem_col));
  DCORA_CUDA(upload(&C->rem_agent, rem_agent));
  C->rem_agent_h = rem_agent;
  DCORA_CUDA(upload(&C->rem_assign, H.rem_assign));
  DCORA_CUDA(upload(&C->rem_val, H.rem_val));
  p->C = C;
  // launch lists cached per agent subset hold row lists of the previous coupling
  p->act_cache.clear();
  p->act.reset();
  cudaFree(p->xe); p->xe = nullptr; p->tracked = nullptr;
  cudaFree(p->hstage); p->hstage = nullptr;
  return DCORA_OK;
}

int64_t dcora_problem_halo_dim(const dcora_problem *p) { return (p && p->C) ? p->C->KH : 0; }

int dcora_problem_update_G(dcora_problem *p, const double *X, const double *Hh,
                           const uint8_t *active) {
  if (!p || !X) return DCORA_ERR_ARG;
  if (!p->C) {
    set_error("dcora_problem_update_G: no coupling registered (dcora_problem_set_coupling)");
    return DCORA_ERR_ARG;
  }
  DCORA_CUDA(cudaSetDevice(p->M->device));
  DeviceMatrices &M = *p->M;
  DeviceCoupling &C = *p->C;
  if (C.KH > 0 && !Hh) {
    set_error("dcora_problem_update_G: halo matrix missing");
    return DCORA_ERR_ARG;
  }
  const int r = p->r;
  if (!p->xe) {
    DCORA_CUDA(cudaMalloc(&p->xe, sizeof(double) * (size_t)std::max<int64_t>(1, (M.K + C.KH) * r)));
    p->tracked = nullptr;
  }
  {
    double *xin = nullptr;  // tracked: no conversion when X is the tracked buffer
    DCORA_TRY(acquire_iterate(p, X, p->xe, &xin));
  }
  if (C.KH > 0) {
    double *dst = p->xe + (size_t)M.K * r;
    const size_t bytes = sizeof(double) * (size_t)C.KH * r;
    const bool dev = is_device_ptr(Hh);
    if (C.identity_perm) {
      DCORA_CUDA(cudaMemcpyAsync(dst, Hh, bytes,
                                 dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                 p->stream));
    } else {
      const double *s = Hh;
      if (!dev) {
        if (!p->hstage) DCORA_CUDA(cudaMalloc(&p->hstage, bytes));
        DCORA_CUDA(cudaMemcpyAsync(p->hstage, Hh, bytes, cudaMemcpyHostToDevice, p->stream));
        s = p->hstage;
      }
      DCORA_TRY(launch_permute_cols(p, C.perm_h, C.KH, s, dst, r, true));
    }
  }
  DCORA_TRY(set_user_mask(p, active));
  SpmmMat A;
  A.coupling = true;
  A.np = M.np; A.browptr = C.browptr; A.bcol = C.bcol; A.bval = C.bval;
  A.prowcol = M.prowcol; A.pose_agent = M.pose_agent;
  A.nrem_rows = C.nrem_rows; A.rem_row = C.rem_row; A.rem_rowptr = C.rem_rowptr;
  A.rem_col = C.rem_col; A.rem_val = C.rem_val; A.rem_assign = C.rem_assign;
  A.rem_agent = C.rem_agent;
  DCORA_TRY(launch_spmm_mat(p, A, p->xe, nullptr, p->G, F_USER));
  p->has_G = true;
  DCORA_CUDA(cudaStreamSynchronize(p->stream));
  return DCORA_OK;
}

int dcora_problem_pack_columns(dcora_problem *p, const double *X, const int32_t *idx,
                               int64_t count, double *out) {
  if (!p || !X || (count > 0 && (!idx || !out)) || count < 0) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  if (count == 0) return DCORA_OK;
  const int r = p->r;
  // cache the index list on the device
  if (p->pack_idx_host.size() != (size_t)count ||
      std::memcmp(p->pack_idx_host.data(), idx, sizeof(int32_t) * count) != 0) {
    for (int64_t j = 0; j < count; ++j)
      if (idx[j] < 0 || idx[j] >= p->M->K) {
        set_error("dcora_problem_pack_columns: column index out of range");
        return DCORA_ERR_ARG;
      }
    p->pack_idx_host.assign(idx, idx + count);
    cudaFree(p->pack_idx);
    p->pack_idx = nullptr;
    DCORA_CUDA(cudaMalloc(&p->pack_idx, sizeof(int) * count));
    DCORA_CUDA(cudaMemcpyAsync(p->pack_idx, idx, sizeof(int) * count, cudaMemcpyHostToDevice,
                               p->stream));
  }
  const double *x = X;
  if (!is_device_ptr(X)) {
    double *t = p->vec(WS_IO0);
    if (!t) return DCORA_ERR_ALLOC;
    DCORA_CUDA(cudaMemcpyAsync(t, X, sizeof(double) * p->vec_len, cudaMemcpyHostToDevice,
                               p->stream));
    x = t;
  }
  const bool dev_out = is_device_ptr(out);
  double *o = out;
  if (!dev_out) {
    o = p->vec(WS_IO1);
    if (!o) return DCORA_ERR_ALLOC;
    if ((size_t)count * r > p->vec_len) {
      set_error("dcora_problem_pack_columns: count exceeds K");
      return DCORA_ERR_ARG;
    }
  }
  const int64_t tot = count * r;
  gather_cols_kernel<<<(int)((tot + 255) / 256), 256, 0, p->stream>>>(count, r, p->pack_idx, x, o);
  DCORA_LAUNCH_CHECK(p);
  if (!dev_out)
    DCORA_CUDA(cudaMemcpyAsync(out, o, sizeof(double) * tot, cudaMemcpyDeviceToHost, p->stream));
  DCORA_CUDA(cudaStreamSynchronize(p->stream));
  return DCORA_OK;
}

int dcora_problem_load_iterate(dcora_problem *p, const double *X_host, double *X_device) {
  if (!p || !X_host || !X_device) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  if (X_device == p->tracked) p->tracked = nullptr;
  DCORA_CUDA(cudaMemcpyAsync(X_device, X_host, sizeof(double) * p->vec_len,
                             cudaMemcpyHostToDevice, p->stream));
  DCORA_CUDA(cudaStreamSynchronize(p->stream));
  return DCORA_OK;
}

int dcora_problem_store_iterate(dcora_problem *p, const double *X_device, double *X_host) {
  if (!p || !X_host || !X_device) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  DCORA_CUDA(cudaMemcpyAsync(X_host, X_device, sizeof(double) * p->vec_len,
                             cudaMemcpyDeviceToHost, p->stream));
  DCORA_CUDA(cudaStreamSynchronize(p->stream));
  return DCORA_OK;
}

int dcora_problem_cost_terms(dcora_problem *p, const double *X, double *quad_out,
                             double *lin_out) {
  if (!p || !X) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  double *x = nullptr, *t = p->vec(WS_IO1);
  if (!t) return DCORA_ERR_ALLOC;
  DCORA_TRY(acquire_iterate(p, X, p->vec(WS_IO0), &x));
  DCORA_TRY(launch_spmm(p, x, nullptr, t, -1));
  const double *xs[2] = {t, x};
  const double *ys[2] = {x, p->G};
  DCORA_TRY(launch_dots(p, 2, xs, ys, -1));
  const int dst[2] = {S_TMP0, S_TMP1};
  DCORA_TRY(launch_reduce_scalars(p, 2, dst));
  DCORA_TRY(read_scalars(p));
  for (int a = 0; a < p->M->A; ++a) {
    if (quad_out) quad_out[a] = 0.5 * p->h_scal[a * NS + S_TMP0];
    if (lin_out) lin_out[a] = p->has_G ? p->h_scal[a * NS + S_TMP1] : 0.0;
  }
  return DCORA_OK;
}

int dcora_problem_combine_project(dcora_problem *p, double a, const double *Am,
                                  double b, const double *Bm, const uint8_t *active,
                                  double *OUT) {
  if (!p || !Am || !Bm || !OUT) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  double *x = p->vec(WS_IO0), *y = p->vec(WS_IO1), *o = p->vec(WS_IO2);
  if (!x || !y || !o) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, Am, x, p->r));
  DCORA_TRY(import_matrix(p, Bm, y, p->r));
  DCORA_TRY(set_user_mask(p, active));
  DCORA_TRY(launch_copy(p, o, x, -1));
  DCORA_TRY(launch_axpby(p, y, -1, a, x, -1, b, y, F_USER));
  DCORA_TRY(launch_project_manifold(p, y, o, F_USER));
  return export_matrix(p, o, OUT, p->r);
}

int dcora_problem_combine(dcora_problem *p, double a, const double *Am, double b,
                          const double *Bm, double *OUT) {
  if (!p || !Am || !Bm || !OUT) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  double *x = p->vec(WS_IO0), *y = p->vec(WS_IO1);
  if (!x || !y) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, Am, x, p->r));
  DCORA_TRY(import_matrix(p, Bm, y, p->r));
  DCORA_TRY(launch_axpby(p, y, -1, a, x, -1, b, y, -1));
  return export_matrix(p, y, OUT, p->r);
}

}  // extern "C"
