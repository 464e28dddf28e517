// Host-side construction of the BLOCK triangular-solve schedule (K5).
//
// The reference applies its preconditioner with CHOLMOD's
// solve (src/QuadraticProblem.cpp:79, src/Graph.cpp:1901-1917).  A scalar,
// level-scheduled SpTRSV has a critical path of thousands of dependent rows
// per agent, which a GPU cannot hide.  Here the factor L (chol.cpp, nested-
// dissection ordering) is regrouped by the ND blocks B_1..B_m (leaf
// subdomains and separators, consecutive factor rows):
//
//   forward   y_B = D_B ( b_B - L[B, <B] y )          D_B = inv(L[B, B])
//   backward  x_B = D_B^T ( y_B - L[>B, B]^T x )
//
// so a block costs one sparse "phase A" (off-diagonal rows, a small SpMM
// against the already solved part) and one dense "phase B" (a triangular
// matrix-vector product with the explicitly inverted diagonal block); blocks
// of one level of the block elimination tree are independent and the tree is
// only O(log n) levels deep.  Off-diagonal entries that couple two pose
// supervariables are stored as dense (d+1)x(d+1) blocks (one index per 128 B,
// like Q itself), everything else (unit-sphere / landmark rows and columns)
// as scalar CSR.
#include <algorithm>
#include <cmath>
#include <numeric>

#include "trsv.h"

namespace dcora {

namespace {

struct Csr {
  std::vector<int> ptr, col;
  std::vector<double> val;
};

// inverse of the lower-triangular diagonal block rows [s, e) of L (CSC)
void invert_diag_block(const HostFactor &F, int s, int e, std::vector<double> *Dout) {
  const int w = e - s;
  std::vector<double> Lb((size_t)w * w, 0.0);  // row-major dense lower
  for (int j = s; j < e; ++j)
    for (int p = F.Lp[j]; p < F.Lp[j + 1]; ++p) {
      const int i = F.Li[p];
      if (i >= e) break;
      Lb[(size_t)(i - s) * w + (j - s)] = F.Lx[p];
    }
  std::vector<double> &D = *Dout;
  D.assign((size_t)w * w, 0.0);
  for (int c = 0; c < w; ++c) {
    D[(size_t)c * w + c] = 1.0 / Lb[(size_t)c * w + c];
    for (int i = c + 1; i < w; ++i) {
      double sacc = 0.0;
      const double *li = &Lb[(size_t)i * w];
      for (int j = c; j < i; ++j) sacc += li[j] * D[(size_t)j * w + c];
      D[(size_t)i * w + c] = -sacc / li[i];
    }
  }
}

}  // namespace

int build_agent_trsv(const HostFactor &F, int dh, AgentTrsv *out) {
  const int k = F.k;
  // ---- final block partition: ND blocks, split to at most TRSV_MAX_BLOCK rows
  std::vector<int> bp(1, 0);
  for (size_t b = 0; b + 1 < F.blk_ptr.size(); ++b) {
    const int s = F.blk_ptr[b], e = F.blk_ptr[b + 1];
    int cur = s;
    for (int r = s; r < e;) {
      const int step = F.pose_start[r] >= 0 ? dh : 1;
      if (r + step - cur > TRSV_MAX_BLOCK && r > cur) { bp.push_back(r); cur = r; }
      r += step;
    }
    bp.push_back(e);
  }
  const int nb = (int)bp.size() - 1;
  std::vector<int> blk_of(k);
  for (int b = 0; b < nb; ++b)
    for (int r = bp[b]; r < bp[b + 1]; ++r) blk_of[r] = b;
  for (int r = 0; r < k; ++r)
    if (F.pose_start[r] >= 0 && blk_of[F.pose_start[r]] != blk_of[r]) return DCORA_ERR_ARG;

  // ---- off-diagonal part of L by rows (forward) and of L^T by rows (backward)
  Csr M[2];
  M[0].ptr.assign(k + 1, 0);
  M[1].ptr.assign(k + 1, 0);
  for (int j = 0; j < k; ++j)
    for (int p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p) {
      const int i = F.Li[p];
      if (blk_of[i] == blk_of[j]) continue;
      M[0].ptr[i + 1]++;
      M[1].ptr[j + 1]++;
    }
  for (int dir = 0; dir < 2; ++dir) {
    for (int i = 0; i < k; ++i) M[dir].ptr[i + 1] += M[dir].ptr[i];
    M[dir].col.resize(M[dir].ptr[k]);
    M[dir].val.resize(M[dir].ptr[k]);
  }
  {
    std::vector<int> f0(M[0].ptr.begin(), M[0].ptr.end() - 1);
    std::vector<int> f1(M[1].ptr.begin(), M[1].ptr.end() - 1);
    for (int j = 0; j < k; ++j)
      for (int p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p) {
        const int i = F.Li[p];
        if (blk_of[i] == blk_of[j]) continue;
        M[0].col[f0[i]] = j; M[0].val[f0[i]++] = F.Lx[p];   // L(i, j), row i
        M[1].col[f1[j]] = i; M[1].val[f1[j]++] = F.Lx[p];   // L^T(j, i), row j
      }
  }

  // ---- inverted diagonal blocks (shared by both directions)
  std::vector<std::vector<double>> Dinv(nb);
  for (int b = 0; b < nb; ++b) invert_diag_block(F, bp[b], bp[b + 1], &Dinv[b]);

  out->k = k;
  out->nblocks = nb;
  for (int dir = 0; dir < 2; ++dir) {
    AgentTrsvDir &T = out->dir[dir];
    const Csr &A = M[dir];
    // levels of the block DAG
    std::vector<int> lvl(nb, 0);
    int nlev = 0;
    for (int t = 0; t < nb; ++t) {
      const int b = dir == 0 ? t : nb - 1 - t;
      int lv = 0;
      for (int r = bp[b]; r < bp[b + 1]; ++r)
        for (int p = A.ptr[r]; p < A.ptr[r + 1]; ++p) lv = std::max(lv, lvl[blk_of[A.col[p]]] + 1);
      lvl[b] = lv;
      nlev = std::max(nlev, lv + 1);
    }
    std::vector<std::vector<int>> by_level(nlev);
    for (int b = 0; b < nb; ++b) by_level[lvl[b]].push_back(b);

    T.srowptr.assign(k + 1, 0);
    T.ta_bptr.assign(1, 0);
    std::vector<int> scnt(k, 0);
    std::vector<int> mark(k, -1);
    for (int lv = 0; lv < nlev; ++lv) {
      TrsvLevel L;
      L.ta0 = (int)T.ta_row.size();
      L.sb0 = (int)T.sb_row0.size();
      L.lr0 = (int)T.lr_blk.size();
      int64_t nblk_level = 0, npose_tasks = 0;
      for (int b : by_level[lv]) {
        const int s = bp[b], e = bp[b + 1], w = e - s;
        bool has = false;
        for (int r = s; r < e && !has; ++r) has = A.ptr[r + 1] > A.ptr[r];
        const int direct = (dir == 0 && !has) ? 1 : 0;
        if (!direct) {
          for (int r = s; r < e;) {
            if (F.pose_start[r] == r) {
              // pose task: union of the source pose blocks over the dh rows
              std::vector<int> cols;
              for (int a = 0; a < dh; ++a)
                for (int p = A.ptr[r + a]; p < A.ptr[r + a + 1]; ++p) {
                  const int ps = F.pose_start[A.col[p]];
                  if (ps >= 0) {
                    if (mark[ps] != r) { mark[ps] = r; cols.push_back(ps); }
                  } else {
                    scnt[r + a]++;
                  }
                }
              std::sort(cols.begin(), cols.end());
              const size_t base = T.bcol.size();
              T.bcol.insert(T.bcol.end(), cols.begin(), cols.end());
              T.bval.resize(T.bcol.size() * (size_t)dh * dh, 0.0);
              for (int a = 0; a < dh; ++a)
                for (int p = A.ptr[r + a]; p < A.ptr[r + a + 1]; ++p) {
                  const int c = A.col[p], ps = F.pose_start[c];
                  if (ps < 0) continue;
                  const size_t q = base + (std::lower_bound(cols.begin(), cols.end(), ps) - cols.begin());
                  T.bval[q * dh * dh + (size_t)a * dh + (c - ps)] = A.val[p];
                }
              T.ta_row.push_back(r);
              T.ta_bptr.push_back((int)T.bcol.size());
              nblk_level += (int64_t)cols.size();
              ++npose_tasks;
              r += dh;
            } else {
              scnt[r] += A.ptr[r + 1] - A.ptr[r];
              T.ta_row.push_back(~r);
              T.ta_bptr.push_back((int)T.bcol.size());
              r += 1;
            }
          }
        }
        // phase B
        const std::vector<double> &D = Dinv[b];
        const int64_t off = (int64_t)T.dense.size();
        if (w <= TRSV_SMALL_BLOCK) {
          // packed by columns: forward lower (column j: rows j..w-1 of D),
          // backward upper (column j: rows 0..j of D^T, i.e. D[j][0..j])
          for (int j = 0; j < w; ++j) {
            if (dir == 0) for (int i = j; i < w; ++i) T.dense.push_back(D[(size_t)i * w + j]);
            else for (int i = 0; i <= j; ++i) T.dense.push_back(D[(size_t)j * w + i]);
          }
          T.sb_row0.push_back(s); T.sb_w.push_back(w); T.sb_off.push_back(off);
          T.sb_direct.push_back(direct);
        } else {
          // packed by rows: forward row i = D[i][0..i]; backward row i of D^T =
          // D[i..w-1][i]
          for (int i = 0; i < w; ++i) {
            if (dir == 0) for (int j = 0; j <= i; ++j) T.dense.push_back(D[(size_t)i * w + j]);
            else for (int j = i; j < w; ++j) T.dense.push_back(D[(size_t)j * w + i]);
          }
          const int lb = (int)T.lb_row0.size();
          T.lb_row0.push_back(s); T.lb_w.push_back(w); T.lb_off.push_back(off);
          T.lb_direct.push_back(direct);
          for (int i = 0; i < w; ++i) { T.lr_blk.push_back(lb); T.lr_i.push_back(i); }
        }
      }
      L.ta1 = (int)T.ta_row.size();
      L.sb1 = (int)T.sb_row0.size();
      L.lr1 = (int)T.lr_blk.size();
      const double mean = npose_tasks ? (double)nblk_level / (double)npose_tasks : 0.0;
      L.gpt = mean <= 3.0 ? 1 : mean <= 6.0 ? 2 : mean <= 12.0 ? 4 : 8;
      L.pad = 0;
      T.levels.push_back(L);
    }
    // scalar CSR of the off-diagonal entries whose source is not a pose column
    // (or whose target row is a scalar row)
    for (int r = 0; r < k; ++r) T.srowptr[r + 1] = T.srowptr[r] + scnt[r];
    T.scol.resize(T.srowptr[k]);
    T.sval.resize(T.srowptr[k]);
    std::vector<int> fill(T.srowptr.begin(), T.srowptr.end() - 1);
    for (int r = 0; r < k; ++r) {
      const bool pose_row = F.pose_start[r] >= 0;
      for (int p = A.ptr[r]; p < A.ptr[r + 1]; ++p) {
        const int c = A.col[p];
        if (pose_row && F.pose_start[c] >= 0) continue;
        T.scol[fill[r]] = c;
        T.sval[fill[r]++] = A.val[p];
      }
    }
    for (int r = 0; r < k; ++r)
      if (fill[r] != T.srowptr[r + 1]) return DCORA_ERR_ARG;
  }
  return DCORA_OK;
}

// CPU emulation of the device schedule (same grouping and operand pairing as
// sptrsv.cu, sequential): x = (L L^T)^-1 b in FACTOR ordering.  Used only by
// the host self check of the set-up code.
void agent_trsv_apply_host(const AgentTrsv &S, int dh, const double *b, double *x) {
  const int k = S.k;
  std::vector<double> T(k), Y(k);
  for (int dir = 0; dir < 2; ++dir) {
    const AgentTrsvDir &D = S.dir[dir];
    for (const TrsvLevel &L : D.levels) {
      for (int t = L.ta0; t < L.ta1; ++t) {
        const int rr = D.ta_row[t];
        const bool pose = rr >= 0;
        const int r0 = pose ? rr : ~rr;
        const int nr = pose ? dh : 1;
        for (int a = 0; a < nr; ++a) {
          double acc = 0.0;
          if (pose)
            for (int p = D.ta_bptr[t]; p < D.ta_bptr[t + 1]; ++p)
              for (int c = 0; c < dh; ++c)
                acc += D.bval[(size_t)p * dh * dh + a * dh + c] * Y[D.bcol[p] + c];
          for (int p = D.srowptr[r0 + a]; p < D.srowptr[r0 + a + 1]; ++p)
            acc += D.sval[p] * Y[D.scol[p]];
          T[r0 + a] = (dir == 0 ? b[r0 + a] : Y[r0 + a]) - acc;
        }
      }
      std::vector<double> tmp;
      for (int s = L.sb0; s < L.sb1; ++s) {
        const int r0 = D.sb_row0[s], w = D.sb_w[s];
        const double *M = D.dense.data() + D.sb_off[s];
        const double *src = D.sb_direct[s] ? b : T.data();
        tmp.assign(w, 0.0);
        size_t o = 0;
        for (int j = 0; j < w; ++j) {
          if (dir == 0) for (int i = j; i < w; ++i) tmp[i] += M[o++] * src[r0 + j];
          else for (int i = 0; i <= j; ++i) tmp[i] += M[o++] * src[r0 + j];
        }
        for (int i = 0; i < w; ++i) Y[r0 + i] = tmp[i];
      }
      // large blocks: all rows read T, then write Y
      std::vector<double> outv(L.lr1 - L.lr0);
      for (int q = L.lr0; q < L.lr1; ++q) {
        const int lb = D.lr_blk[q], i = D.lr_i[q];
        const int r0 = D.lb_row0[lb], w = D.lb_w[lb];
        const double *src = D.lb_direct[lb] ? b : T.data();
        double acc = 0.0;
        if (dir == 0) {
          const double *row = D.dense.data() + D.lb_off[lb] + (int64_t)i * (i + 1) / 2;
          for (int j = 0; j <= i; ++j) acc += row[j] * src[r0 + j];
        } else {
          const double *row = D.dense.data() + D.lb_off[lb] + (int64_t)i * w - (int64_t)i * (i - 1) / 2;
          for (int j = i; j < w; ++j) acc += row[j - i] * src[r0 + j];
        }
        outv[q - L.lr0] = acc;
      }
      for (int q = L.lr0; q < L.lr1; ++q) Y[D.lb_row0[D.lr_blk[q]] + D.lr_i[q]] = outv[q - L.lr0];
    }
  }
  for (int i = 0; i < k; ++i) x[i] = Y[i];
}

}  // namespace dcora
