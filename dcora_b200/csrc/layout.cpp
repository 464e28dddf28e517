// Host-side conversion of the agents' data matrices (CSR, reference column
// ordering) to the device layout used by the sm_100a kernels.
//
// Internal column order of agent a (columns [coff_a, coff_a + k_a)):
//     [ Y_1 p_1 | Y_2 p_2 | ... | Y_n p_n | r_1 .. r_l | L_1 .. L_b ]
// i.e. pose blocks of (d+1) contiguous columns, then unit spheres, then
// landmarks.  For DCORA_ORDER_SE this is the reference ordering itself
// (include/DCORA/manifold/LiftedVariable.h:27-30); for DCORA_ORDER_RA
// ([Y..|r..|p..|L..], LiftedVariable.h:148-152) it is a permutation that is
// applied on device when matrices cross the C-ABI.
//
// Q is split into
//   * BSR over pose blocks, dense (d+1)x(d+1) blocks, one 128-byte line per
//     3-D block (the connection-Laplacian part, Graph.cpp:579-683 / Q11,Q13,
//     pose part of Q33, Graph.cpp:1137-1141);
//   * a scalar CSR remainder holding every entry that touches a unit-sphere
//     or landmark row/column (Q22, Q23, landmark part of Q33).
// Both store Q^T (output-column major) so that Y = X*Q is reproduced with the
// reference's own pairing of operands even if Q is only symmetric to rounding.
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <cstring>

#include "common.cuh"

namespace dcora {

int setup_threads() {
  if (const char *e = std::getenv("DCORA_SETUP_THREADS")) {
    const int v = std::atoi(e);
    if (v > 0) return v;
  }
  int hw = (int)std::thread::hardware_concurrency();
  if (hw <= 0) hw = 1;
  int local = 1;
  if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) local = std::max(1, std::atoi(e));
  return std::max(1, hw / local);
}


static inline int ref_to_internal(int c, int d, int n, int l, int order) {
  if (order == DCORA_ORDER_SE) return c;
  const int dn = d * n;
  if (c < dn) return (c / d) * (d + 1) + (c % d);         // rotation column
  if (c < dn + l) return (d + 1) * n + (c - dn);          // unit sphere
  if (c < dn + l + n) return (c - dn - l) * (d + 1) + d;  // translation
  return (d + 1) * n + l + (c - dn - l - n);              // landmark
}

namespace {
struct AgentLayout {
  std::vector<int> m;                // reference col -> internal col (local)
  std::vector<int> browptr, bcol;    // local block cols (pose index j)
  std::vector<double> bval;
  std::vector<int> rem_row, rem_rowptr, rem_col;  // local internal indices
  std::vector<uint8_t> rem_assign;
  std::vector<double> rem_val;
  int rc = DCORA_OK;
};

void convert_agent(const dcora_agent_desc &g, int d, int order, AgentLayout *out) {
  AgentLayout &L = *out;
  const int dh = d + 1;
  const int k = dh * g.n + g.l + g.b;
  const int npc = dh * g.n;
  L.m.resize(k);
  for (int c = 0; c < k; ++c) L.m[c] = ref_to_internal(c, d, g.n, g.l, order);
  // pass A: block-column sets per output block row; remainder counts
  std::vector<std::vector<int>> cols(g.n);
  std::vector<int> rem_count(k, 0);
  for (int R = 0; R < k; ++R) {
    const int ri = L.m[R];
    for (int p = g.rowptr[R]; p < g.rowptr[R + 1]; ++p) {
      const int C = g.col[p];
      if (C < 0 || C >= k) { L.rc = DCORA_ERR_ARG; return; }
      const int ci = L.m[C];
      if (ri < npc && ci < npc) {
        std::vector<int> &v = cols[ci / dh];
        const int j = ri / dh;
        if (std::find(v.begin(), v.end(), j) == v.end()) v.push_back(j);
      } else {
        rem_count[ci]++;
      }
    }
  }
  L.browptr.assign(g.n + 1, 0);
  for (int i = 0; i < g.n; ++i) {
    std::sort(cols[i].begin(), cols[i].end());
    L.browptr[i + 1] = L.browptr[i] + (int)cols[i].size();
  }
  L.bcol.resize(L.browptr[g.n]);
  for (int i = 0; i < g.n; ++i)
    std::copy(cols[i].begin(), cols[i].end(), L.bcol.begin() + L.browptr[i]);
  std::vector<std::vector<int>>().swap(cols);
  L.bval.assign((size_t)L.bcol.size() * dh * dh, 0.0);
  // remainder rows: every sphere/landmark row (assign) + pose rows with entries
  std::vector<int> slot(k, -1);
  L.rem_rowptr.assign(1, 0);
  for (int ci = 0; ci < k; ++ci) {
    const bool extra = ci >= npc;
    if (extra || rem_count[ci] > 0) {
      slot[ci] = (int)L.rem_row.size();
      L.rem_row.push_back(ci);
      L.rem_assign.push_back(extra ? 1 : 0);
      L.rem_rowptr.push_back(L.rem_rowptr.back() + rem_count[ci]);
    }
  }
  L.rem_col.assign(L.rem_rowptr.back(), 0);
  L.rem_val.assign(L.rem_rowptr.back(), 0.0);
  std::vector<int> fill(L.rem_rowptr.begin(), L.rem_rowptr.end() - 1);
  // pass B: values.  Entry (R, C, v) of Q contributes X[:,R]*v to Y[:,C].
  for (int R = 0; R < k; ++R) {
    const int ri = L.m[R];
    for (int p = g.rowptr[R]; p < g.rowptr[R + 1]; ++p) {
      const int ci = L.m[g.col[p]];
      const double v = g.val[p];
      if (ri < npc && ci < npc) {
        const int i = ci / dh, aa = ci % dh, j = ri / dh, bb = ri % dh;
        const int *b0 = L.bcol.data() + L.browptr[i];
        const int *b1 = L.bcol.data() + L.browptr[i + 1];
        const int64_t q = std::lower_bound(b0, b1, j) - L.bcol.data();
        L.bval[(size_t)q * dh * dh + aa * dh + bb] += v;
      } else {
        const int w = fill[slot[ci]]++;
        L.rem_col[w] = ri;
        L.rem_val[w] = v;
      }
    }
  }
}
}  // namespace

int build_layout(const dcora_agent_desc *agents, int A, int r, int d, int order,
                 HostLayout *out) {
  if (A <= 0 || !agents || (d != 2 && d != 3) || r < 1 || r > MAX_RANK) {
    set_error("build_layout: bad arguments (need d in {2,3}, 1<=r<=16)");
    return DCORA_ERR_ARG;
  }
  HostLayout &H = *out;
  H.r = r; H.d = d; H.order = order; H.A = A;
  const int dh = d + 1;
  H.coff.assign(A + 1, 0);
  H.poff.assign(A + 1, 0);
  H.n.resize(A); H.l.resize(A); H.b.resize(A);
  H.nnz_in = 0;
  for (int a = 0; a < A; ++a) {
    const dcora_agent_desc &g = agents[a];
    if (g.n < 0 || g.l < 0 || g.b < 0 || !g.rowptr || (!g.col && g.rowptr[0] != 0)) {
      set_error("build_layout: bad agent descriptor");
      return DCORA_ERR_ARG;
    }
    if (order == DCORA_ORDER_SE && (g.l || g.b)) {
      set_error("build_layout: SE ordering cannot carry unit spheres / landmarks");
      return DCORA_ERR_ARG;
    }
    H.n[a] = g.n; H.l[a] = g.l; H.b[a] = g.b;
    const int64_t k = (int64_t)dh * g.n + g.l + g.b;
    H.coff[a + 1] = H.coff[a] + k;
    H.poff[a + 1] = H.poff[a] + g.n;
    H.nnz_in += g.rowptr[k];
  }
  H.K = H.coff[A];
  H.np = H.poff[A];
  if (H.K * (int64_t)MAX_RANK >= ((int64_t)1 << 31) * 8) {
    set_error("build_layout: problem dimension exceeds indexing range");
    return DCORA_ERR_ARG;
  }
  if (H.K >= ((int64_t)1 << 31)) {
    set_error("build_layout: K exceeds int32");
    return DCORA_ERR_ARG;
  }
  std::vector<AgentLayout> AL(A);
#pragma omp parallel for schedule(dynamic, 1) num_threads(setup_threads())
  for (int a = 0; a < A; ++a) convert_agent(agents[a], d, order, &AL[a]);
  for (int a = 0; a < A; ++a)
    if (AL[a].rc != DCORA_OK) {
      set_error("build_layout: column index out of range");
      return AL[a].rc;
    }

  // ---- merge
  H.identity_perm = (order == DCORA_ORDER_SE);
  H.perm.resize(H.K);
  H.prowcol.resize(H.np);
  H.browptr.assign(H.np + 1, 0);
  size_t nnzb = 0, nrow = 0, nrem = 0;
  for (int a = 0; a < A; ++a) {
    nnzb += AL[a].bcol.size();
    nrow += AL[a].rem_row.size();
    nrem += AL[a].rem_col.size();
  }
  H.bcol.resize(nnzb);
  H.bval.resize(nnzb * dh * dh);
  H.rem_row.resize(nrow); H.rem_assign.resize(nrow);
  H.rem_rowptr.assign(nrow + 1, 0);
  H.rem_col.resize(nrem); H.rem_val.resize(nrem);
  size_t ob = 0, orow = 0, orem = 0;
  for (int a = 0; a < A; ++a) {
    const AgentLayout &L = AL[a];
    const int c0 = (int)H.coff[a];
    const int k = (int)(H.coff[a + 1] - H.coff[a]);
    for (int c = 0; c < k; ++c) H.perm[c0 + L.m[c]] = c0 + c;
    for (int i = 0; i < H.n[a]; ++i) {
      H.prowcol[H.poff[a] + i] = c0 + i * dh;
      H.browptr[H.poff[a] + i + 1] = (int)ob + L.browptr[i + 1];
    }
    for (size_t q = 0; q < L.bcol.size(); ++q) H.bcol[ob + q] = c0 + L.bcol[q] * dh;
    std::copy(L.bval.begin(), L.bval.end(), H.bval.begin() + ob * dh * dh);
    for (size_t q = 0; q < L.rem_row.size(); ++q) {
      H.rem_row[orow + q] = c0 + L.rem_row[q];
      H.rem_assign[orow + q] = L.rem_assign[q];
      H.rem_rowptr[orow + q + 1] = (int)orem + L.rem_rowptr[q + 1];
    }
    for (size_t q = 0; q < L.rem_col.size(); ++q) {
      H.rem_col[orem + q] = c0 + L.rem_col[q];
      H.rem_val[orem + q] = L.rem_val[q];
    }
    ob += L.bcol.size(); orow += L.rem_row.size(); orem += L.rem_col.size();
  }

  // ---- chunk table
  H.chunks.clear();
  H.agent_chunk0.assign(A + 1, 0);
  for (int a = 0; a < A; ++a) {
    H.agent_chunk0[a] = (int)H.chunks.size();
    const int c0 = (int)H.coff[a];
    for (int i = 0; i < H.n[a]; i += POSE_CHUNK) {
      const int np = std::min(POSE_CHUNK, H.n[a] - i);
      H.chunks.push_back({c0 + i * dh, np * dh, KIND_POSE, a});
    }
    int base = c0 + dh * H.n[a];
    for (int i = 0; i < H.l[a]; i += EXTRA_CHUNK)
      H.chunks.push_back({base + i, std::min(EXTRA_CHUNK, H.l[a] - i), KIND_SPHERE, a});
    base += H.l[a];
    for (int i = 0; i < H.b[a]; i += EXTRA_CHUNK)
      H.chunks.push_back({base + i, std::min(EXTRA_CHUNK, H.b[a] - i), KIND_EUCLID, a});
  }
  H.agent_chunk0[A] = (int)H.chunks.size();
  return DCORA_OK;
}


// ---------------------------------------------------------------- coupling
// Cross-agent coupling matrix C ((K + KH) x K, CSR): G = [X | H] * C restates
// Graph::constructLinearCostTerm{PGO,RASLAM} (Graph.cpp:685-822,1190-1772) as
// one sparse product -- every shared edge contributes Xc * (cross block of the
// global Q).  Rows are "extended reference" columns: the K local columns
// (agents concatenated, each in the reference ordering) followed by the halo
// segments (neighbour public states received from other processes), each
// segment a pseudo-agent (n, l, b) in the same ordering.
int build_coupling_layout(int d, int order, const std::vector<int64_t> &coff,
                          const std::vector<int> &n, const std::vector<int> &l,
                          const std::vector<int> &poff, const dcora_halo_segment *segs,
                          int nseg, const int64_t *rowptr, const int32_t *col,
                          const double *val, const int64_t *row_key, HostCoupling *out) {
  HostCoupling &C = *out;
  const int dh = d + 1;
  const int A = (int)n.size();
  const int64_t K = coff[A];
  std::vector<int64_t> hoff(nseg + 1, 0);
  for (int s = 0; s < nseg; ++s) {
    if (segs[s].n < 0 || segs[s].l < 0 || segs[s].b < 0) {
      set_error("build_coupling_layout: bad halo segment");
      return DCORA_ERR_ARG;
    }
    if (order == DCORA_ORDER_SE && (segs[s].l || segs[s].b)) {
      set_error("build_coupling_layout: SE ordering cannot carry spheres / landmarks");
      return DCORA_ERR_ARG;
    }
    hoff[s + 1] = hoff[s] + (int64_t)dh * segs[s].n + segs[s].l + segs[s].b;
  }
  C.KH = hoff[nseg];
  const int64_t KE = K + C.KH;
  if (KE >= ((int64_t)1 << 31)) {
    set_error("build_coupling_layout: K + KH exceeds int32");
    return DCORA_ERR_ARG;
  }
  // extended reference column -> extended internal column; pose-block start
  std::vector<int> r2i(KE), blk(KE, -1);
  for (int a = 0; a < A; ++a) {
    const int c0 = (int)coff[a], k = (int)(coff[a + 1] - coff[a]);
    for (int c = 0; c < k; ++c) r2i[c0 + c] = c0 + ref_to_internal(c, d, n[a], l[a], order);
    for (int i = 0; i < n[a]; ++i)
      for (int q = 0; q < dh; ++q) blk[c0 + i * dh + q] = c0 + i * dh;
  }
  C.identity_perm = (order == DCORA_ORDER_SE);
  C.perm_h.resize(C.KH);
  for (int s = 0; s < nseg; ++s) {
    const int c0 = (int)(K + hoff[s]), k = (int)(hoff[s + 1] - hoff[s]);
    for (int c = 0; c < k; ++c) {
      const int ci = ref_to_internal(c, d, segs[s].n, segs[s].l, order);
      r2i[c0 + c] = c0 + ci;
      C.perm_h[hoff[s] + ci] = (int)hoff[s] + c;
    }
    for (int i = 0; i < segs[s].n; ++i)
      for (int q = 0; q < dh; ++q) blk[c0 + i * dh + q] = c0 + i * dh;
  }
  // local internal column -> output pose block index
  const int np = poff[A];
  std::vector<int> oblk(K, -1);
  for (int a = 0; a < A; ++a)
    for (int i = 0; i < n[a]; ++i)
      for (int q = 0; q < dh; ++q) oblk[coff[a] + i * dh + q] = poff[a] + i;
  C.prowcol.resize(np);
  for (int a = 0; a < A; ++a)
    for (int i = 0; i < n[a]; ++i) C.prowcol[poff[a] + i] = (int)coff[a] + i * dh;

  std::vector<std::vector<int>> cols(np);
  std::vector<int> rem_count(K, 0);
  for (int64_t R = 0; R < KE; ++R) {
    const int ri = r2i[R];
    for (int64_t p = rowptr[R]; p < rowptr[R + 1]; ++p) {
      if (col[p] < 0 || col[p] >= K) {
        set_error("build_coupling_layout: column index out of range");
        return DCORA_ERR_ARG;
      }
      const int ci = r2i[col[p]];
      if (blk[ri] >= 0 && oblk[ci] >= 0) {
        std::vector<int> &v = cols[oblk[ci]];
        if (std::find(v.begin(), v.end(), blk[ri]) == v.end()) v.push_back(blk[ri]);
      } else {
        rem_count[ci]++;
      }
    }
  }
  // accumulation order of the sources of an output column: ascending key.  With the
  // caller's keys (the GLOBAL column of every extended row) G is summed in the same order
  // however the agents are dealt to ranks: iterates are bit-identical across world sizes.
  std::vector<int64_t> keyI(KE);
  for (int64_t R = 0; R < KE; ++R) keyI[r2i[R]] = row_key ? row_key[R] : r2i[R];
  auto by_key = [&](int a, int b) { return keyI[a] < keyI[b] || (keyI[a] == keyI[b] && a < b); };
  C.browptr.assign(np + 1, 0);
  for (int i = 0; i < np; ++i) {
    std::sort(cols[i].begin(), cols[i].end(), by_key);
    C.browptr[i + 1] = C.browptr[i] + (int)cols[i].size();
  }
  C.bcol.resize(C.browptr[np]);
  for (int i = 0; i < np; ++i)
    std::copy(cols[i].begin(), cols[i].end(), C.bcol.begin() + C.browptr[i]);
  std::vector<std::vector<int>>().swap(cols);
  C.bval.assign((size_t)C.bcol.size() * dh * dh, 0.0);
  std::vector<int> slot(K, -1);
  C.rem_rowptr.assign(1, 0);
  for (int ci = 0; ci < K; ++ci) {
    const bool extra = oblk[ci] < 0;
    if (extra || rem_count[ci] > 0) {
      slot[ci] = (int)C.rem_row.size();
      C.rem_row.push_back(ci);
      C.rem_assign.push_back(extra ? 1 : 0);
      C.rem_rowptr.push_back(C.rem_rowptr.back() + rem_count[ci]);
    }
  }
  C.rem_col.assign(C.rem_rowptr.back(), 0);
  C.rem_val.assign(C.rem_rowptr.back(), 0.0);
  std::vector<int> fill(C.rem_rowptr.begin(), C.rem_rowptr.end() - 1);
  for (int64_t R = 0; R < KE; ++R) {
    const int ri = r2i[R];
    for (int64_t p = rowptr[R]; p < rowptr[R + 1]; ++p) {
      const int ci = r2i[col[p]];
      if (blk[ri] >= 0 && oblk[ci] >= 0) {
        const int i = oblk[ci];
        const int *b0 = C.bcol.data() + C.browptr[i];
        const int *b1 = C.bcol.data() + C.browptr[i + 1];
        const int64_t q = std::find(b0, b1, blk[ri]) - C.bcol.data();
        // component of the output column / source column inside its block
        const int a_out = ci - C.prowcol[i], b_src = ri - blk[ri];
        C.bval[(size_t)q * dh * dh + a_out * dh + b_src] += val[p];
      } else {
        const int w = fill[slot[ci]]++;
        C.rem_col[w] = ri;
        C.rem_val[w] = val[p];
      }
    }
  }
  if (row_key) {
    std::vector<std::pair<int, double>> tmp;
    for (size_t s = 0; s + 1 < C.rem_rowptr.size(); ++s) {
      const int b = C.rem_rowptr[s], e = C.rem_rowptr[s + 1];
      tmp.clear();
      for (int w = b; w < e; ++w) tmp.emplace_back(C.rem_col[w], C.rem_val[w]);
      std::stable_sort(tmp.begin(), tmp.end(),
                       [&](const std::pair<int, double> &x, const std::pair<int, double> &y) {
                         return by_key(x.first, y.first);
                       });
      for (int w = b; w < e; ++w) { C.rem_col[w] = tmp[w - b].first; C.rem_val[w] = tmp[w - b].second; }
    }
  }
  return DCORA_OK;
}

}  // namespace dcora
