// Host-side construction of the BLOCK triangular-solve tapes (K5).
//
// The reference applies its preconditioner with CHOLMOD's solve
// (src/QuadraticProblem.cpp:79, src/Graph.cpp:1901-1917).  A scalar,
// level-scheduled SpTRSV has a critical path of thousands of dependent rows
// per agent, which a GPU cannot hide.  Here the factor L (chol.cpp, nested-
// dissection ordering) is regrouped by the ND blocks B_1..B_m (leaf
// subdomains and separators, consecutive factor rows):
//
//   forward   y_B = D_B ( b_B - L[B, <B] y )          D_B = inv(L[B, B])
//   backward  x_B = D_B^T ( y_B - L[>B, B]^T x )
//
// so a block costs one sparse "phase A" (off-diagonal rows against the already
// solved part, stored as dense (d+1)x(d+1) blocks) and one dense "phase B" (a
// triangular matrix-vector product with the explicitly inverted diagonal
// block); blocks of one level of the block elimination tree are independent
// and the tree is only O(log n) levels deep.  The work of every phase is dealt
// to the warps of the agent's cluster here, on the host, and written out as
// one tape per warp (format: trsv.h).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <queue>

#include "trsv.h"

namespace dcora {

namespace {

struct Csr {
  std::vector<int> ptr, col;
  std::vector<double> val;
};

inline uint64_t pack2(int lo, int hi) {
  return (uint64_t)(uint32_t)lo | ((uint64_t)(uint32_t)hi << 32);
}
inline uint64_t dbits(double v) {
  uint64_t u;
  std::memcpy(&u, &v, 8);
  return u;
}
inline double bitsd(uint64_t u) {
  double v;
  std::memcpy(&v, &u, 8);
  return v;
}
inline int lo32(uint64_t u) { return (int)(uint32_t)(u & 0xffffffffu); }
inline int hi32(uint64_t u) { return (int)(uint32_t)(u >> 32); }

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

struct ATask {
  int r0, nrows;
  std::vector<int> src;        // source row starts, sorted
  std::vector<double> val;     // dh*dh per source, [a][b]
  int gpt = 1;
};

struct Round {
  int gpt, niter;
  std::vector<int> tasks;      // indices into the level's task list (<= 8/gpt)
};

// longest-processing-time assignment: items sorted by decreasing cost
struct Lpt {
  typedef std::pair<int64_t, int> LI;
  std::priority_queue<LI, std::vector<LI>, std::greater<LI>> pq;
  explicit Lpt(int nw) { for (int w = 0; w < nw; ++w) pq.push(LI(0, w)); }
  int take(int64_t cost) {
    LI t = pq.top();
    pq.pop();
    pq.push(LI(t.first + cost, t.second));
    return t.second;
  }
};

}  // namespace

int build_agent_tape(const HostFactor &F, int dh, int nwarps, int nwarps_total, int row_offset,
                     AgentTape *out) {
  const int k = F.k;
  const int bs = dh * dh;
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
  out->nwarps = nwarps_total;
  out->dh = dh;
  out->warp_tape.assign(nwarps_total, std::vector<uint64_t>());
  out->stored_values = 0;
  const int R0 = row_offset;
  // first row of the (d+1)-row window a source column is read through: its pose
  // block, or -- for a unit-sphere / landmark row -- a window that stays inside
  // the agent's own rows (the other rows of the window meet zeros)
  auto src_start = [&](int c) {
    if (F.pose_start[c] >= 0) return F.pose_start[c];
    return std::max(0, std::min(c, k - dh));
  };

  for (int dir = 0; dir < 2; ++dir) {
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
    out->nlev[dir] = nlev;
    std::vector<std::vector<int>> by_level(nlev);
    for (int b = 0; b < nb; ++b) by_level[lvl[b]].push_back(b);

    for (int lv = 0; lv < nlev; ++lv) {
      // ------------------------------------------------ phase A tasks
      std::vector<ATask> tasks;
      std::vector<int> direct_of(by_level[lv].size(), 0);
      int64_t total_blocks = 0;
      for (size_t bi = 0; bi < by_level[lv].size(); ++bi) {
        const int b = by_level[lv][bi];
        const int s = bp[b], e = bp[b + 1];
        bool has = false;
        for (int r = s; r < e && !has; ++r) has = A.ptr[r + 1] > A.ptr[r];
        // only small forward blocks without predecessors read their right-hand
        // side directly (through perm); everything else gets a phase-A task
        const int direct = (dir == 0 && !has && e - s <= TRSV_SMALL_BLOCK) ? 1 : 0;
        direct_of[bi] = direct;
        if (direct) continue;
        for (int r = s; r < e;) {
          ATask tk;
          tk.r0 = r;
          tk.nrows = F.pose_start[r] == r ? dh : 1;
          // sources: pose blocks (start row ps) or single scalar rows (b = 0)
          for (int a = 0; a < tk.nrows; ++a)
            for (int p = A.ptr[r + a]; p < A.ptr[r + a + 1]; ++p) {
              tk.src.push_back(src_start(A.col[p]));
            }
          std::sort(tk.src.begin(), tk.src.end());
          tk.src.erase(std::unique(tk.src.begin(), tk.src.end()), tk.src.end());
          tk.val.assign(tk.src.size() * (size_t)bs, 0.0);
          for (int a = 0; a < tk.nrows; ++a)
            for (int p = A.ptr[r + a]; p < A.ptr[r + a + 1]; ++p) {
              const int c = A.col[p];
              const int ps = src_start(c);
              const size_t q = std::lower_bound(tk.src.begin(), tk.src.end(), ps) - tk.src.begin();
              tk.val[q * bs + (size_t)a * dh + (c - ps)] = A.val[p];
            }
          total_blocks += (int64_t)tk.src.size();
          r += tk.nrows;
          tasks.push_back(std::move(tk));
        }
      }
      const bool a_present = !tasks.empty();
      // groups per task: split tasks that are long compared with the ideal
      // per-group load of the level
      const int64_t target = std::max<int64_t>(2, (total_blocks + (int64_t)nwarps * 8 - 1) / ((int64_t)nwarps * 8));
      for (ATask &tk : tasks) {
        int g = 1;
        while (g < 8 && (int64_t)tk.src.size() > target * g) g *= 2;
        tk.gpt = g;
      }
      std::vector<int> ord(tasks.size());
      std::iota(ord.begin(), ord.end(), 0);
      std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) {
        if (tasks[x].gpt != tasks[y].gpt) return tasks[x].gpt > tasks[y].gpt;
        return tasks[x].src.size() > tasks[y].src.size();
      });
      std::vector<Round> rounds;
      for (size_t i = 0; i < ord.size();) {
        Round rd;
        rd.gpt = tasks[ord[i]].gpt;
        rd.niter = 0;
        const int tpw = 8 / rd.gpt;
        while (i < ord.size() && (int)rd.tasks.size() < tpw && tasks[ord[i]].gpt == rd.gpt) {
          const int nbk = (int)tasks[ord[i]].src.size();
          rd.niter = std::max(rd.niter, (nbk + rd.gpt - 1) / rd.gpt);
          rd.tasks.push_back(ord[i]);
          ++i;
        }
        rounds.push_back(std::move(rd));
      }
      std::stable_sort(rounds.begin(), rounds.end(),
                       [](const Round &x, const Round &y) { return x.niter > y.niter; });
      std::vector<std::vector<int>> warp_rounds(nwarps);
      {
        Lpt lpt(nwarps);
        for (size_t i = 0; i < rounds.size(); ++i)
          warp_rounds[lpt.take(rounds[i].niter + 2)].push_back((int)i);
      }
      // ------------------------------------------------ phase B items
      struct Small { int b, direct; };
      struct Large { int b, i, len; };
      std::vector<Small> smalls;
      std::vector<Large> larges;
      for (size_t bi = 0; bi < by_level[lv].size(); ++bi) {
        const int b = by_level[lv][bi];
        const int w = bp[b + 1] - bp[b];
        if (w <= TRSV_SMALL_BLOCK) smalls.push_back({b, direct_of[bi]});
        else for (int i = 0; i < w; ++i) larges.push_back({b, i, dir == 0 ? i + 1 : w - i});
      }
      std::stable_sort(smalls.begin(), smalls.end(), [&](const Small &x, const Small &y) {
        return bp[x.b + 1] - bp[x.b] > bp[y.b + 1] - bp[y.b];
      });
      std::stable_sort(larges.begin(), larges.end(),
                       [](const Large &x, const Large &y) { return x.len > y.len; });
      std::vector<std::vector<int>> warp_small(nwarps), warp_large(nwarps);
      {
        Lpt lpt(nwarps);
        for (size_t i = 0; i < smalls.size(); ++i)
          warp_small[lpt.take(bp[smalls[i].b + 1] - bp[smalls[i].b] + 16)].push_back((int)i);
        for (size_t i = 0; i < larges.size(); ++i)
          warp_large[lpt.take(larges[i].len / 16 + 8)].push_back((int)i);
      }
      // ------------------------------------------------ emit
      for (int wi = 0; wi < nwarps; ++wi) {
        std::vector<uint64_t> &tp = out->warp_tape[wi];
        tp.push_back(pack2((int)warp_rounds[wi].size(), a_present ? 1 : 0));
        for (int ri : warp_rounds[wi]) {
          const Round &rd = rounds[ri];
          tp.push_back(pack2(rd.niter, rd.gpt));
          for (int g = 0; g < 8; ++g) {
            const int ti = g / rd.gpt;
            if (ti < (int)rd.tasks.size()) {
              const ATask &tk = tasks[rd.tasks[ti]];
              tp.push_back(pack2(R0 + tk.r0, tk.nrows));
              tp.push_back(pack2(R0 + F.perm[tk.r0], 0));
            } else {
              tp.push_back(pack2(-1, 0));
              tp.push_back(pack2(0, 0));
            }
          }
          while (tp.size() % 4) tp.push_back(0);  // iterations start 32-byte aligned
          for (int it = 0; it < rd.niter; ++it) {
            int c[8];
            const double *vp[8];
            for (int g = 0; g < 8; ++g) {
              const int ti = g / rd.gpt, gg = g % rd.gpt;
              c[g] = R0;  // padding: any valid row, zero block
              vp[g] = nullptr;
              if (ti < (int)rd.tasks.size()) {
                const ATask &tk = tasks[rd.tasks[ti]];
                const int q = it * rd.gpt + gg;
                if (q < (int)tk.src.size()) { c[g] = R0 + tk.src[q]; vp[g] = &tk.val[(size_t)q * bs]; }
              }
            }
            for (int g = 0; g < 8; g += 2) tp.push_back(pack2(c[g], c[g + 1]));
            // lane (group g, column bb) owns DH consecutive words: its column of the
            // block, rows 0..DH-1 (one address computation and, for DH = 4, two
            // 128-bit shared loads per iteration on the device)
            for (int g = 0; g < 8; ++g)
              for (int bb = 0; bb < dh; ++bb)
                for (int a = 0; a < dh; ++a)
                  tp.push_back(dbits(vp[g] ? vp[g][a * dh + bb] : 0.0));
            out->stored_values += 8 * bs;
          }
        }
        tp.push_back(pack2((int)warp_small[wi].size(), (int)warp_large[wi].size()));
        for (int si : warp_small[wi]) {
          const int b = smalls[si].b, s = bp[b], w = bp[b + 1] - s;
          const std::vector<double> &D = Dinv[b];
          tp.push_back(pack2(R0 + s, w));
          tp.push_back(pack2(smalls[si].direct, 0));
          for (int i = 0; i < w; i += 2)
            tp.push_back(pack2(R0 + F.perm[s + i], i + 1 < w ? R0 + F.perm[s + i + 1] : 0));
          // packed by columns: forward lower (column j: rows j..w-1 of D),
          // backward upper (column j: rows 0..j of D^T, i.e. D[j][0..j])
          for (int j = 0; j < w; ++j) {
            if (dir == 0) for (int i = j; i < w; ++i) tp.push_back(dbits(D[(size_t)i * w + j]));
            else for (int i = 0; i <= j; ++i) tp.push_back(dbits(D[(size_t)j * w + i]));
          }
          out->stored_values += (int64_t)w * (w + 1) / 2;
        }
        for (int li : warp_large[wi]) {
          const Large &lg = larges[li];
          const int s = bp[lg.b], w = bp[lg.b + 1] - s, i = lg.i;
          const std::vector<double> &D = Dinv[lg.b];
          const int j0 = dir == 0 ? 0 : i;
          tp.push_back(pack2(R0 + s + j0, lg.len));
          tp.push_back(pack2(R0 + s + i, R0 + F.perm[s + i]));
          // forward row i = D[i][0..i]; backward row i of D^T = D[i..w-1][i]
          if (dir == 0) for (int j = 0; j <= i; ++j) tp.push_back(dbits(D[(size_t)i * w + j]));
          else for (int j = i; j < w; ++j) tp.push_back(dbits(D[(size_t)j * w + i]));
          out->stored_values += lg.len;
        }
      }
      // warps of a larger launch cluster than this agent's own (batches mix
      // agent sizes; the work split of an agent depends on the agent alone, so
      // its result does not depend on what it is batched with): barriers only
      for (int wi = nwarps; wi < nwarps_total; ++wi) {
        out->warp_tape[wi].push_back(pack2(0, a_present ? 1 : 0));
        out->warp_tape[wi].push_back(pack2(0, 0));
      }
    }
  }
  for (std::vector<uint64_t> &tp : out->warp_tape)
    while (tp.size() % TAPE_CHUNK_WORDS) tp.push_back(0);
  return DCORA_OK;
}

// CPU replay of the tapes, warp by warp within a phase (phases are separated
// by barriers on the device).
void agent_tape_apply_host(const AgentTape &S, const double *b, double *x, double *T,
                           double *Y) {
  const int dh = S.dh, bs = dh * dh, nw = S.nwarps;
  std::vector<size_t> pos(nw, 0);
  struct Write { int row; double v; };
  for (int dir = 0; dir < 2; ++dir) {
    for (int lv = 0; lv < S.nlev[dir]; ++lv) {
      // phase A
      for (int wi = 0; wi < nw; ++wi) {
        const std::vector<uint64_t> &tp = S.warp_tape[wi];
        size_t &p = pos[wi];
        const int nrounds = lo32(tp[p]);
        ++p;
        for (int rd = 0; rd < nrounds; ++rd) {
          const int niter = lo32(tp[p]), gpt = hi32(tp[p]);
          ++p;
          int r0[8], nrows[8], pb[8];
          for (int g = 0; g < 8; ++g) {
            r0[g] = lo32(tp[p]); nrows[g] = hi32(tp[p]); pb[g] = lo32(tp[p + 1]);
            p += 2;
          }
          std::vector<double> acc(8 * 4, 0.0);
          p = (p + 3) & ~(size_t)3;
          for (int it = 0; it < niter; ++it) {
            int c[8];
            for (int g = 0; g < 8; g += 2) { c[g] = lo32(tp[p]); c[g + 1] = hi32(tp[p]); ++p; }
            for (int g = 0; g < 8; ++g)
              for (int a = 0; a < dh; ++a)
                for (int bb = 0; bb < dh; ++bb)
                  acc[g * 4 + a] += bitsd(tp[p + (size_t)(g * dh + bb) * dh + a]) * Y[c[g] + bb];
            p += 8 * bs;
          }
          for (int g = 0; g < 8; g += gpt) {
            if (r0[g] < 0) continue;
            for (int a = 0; a < nrows[g]; ++a) {
              double s = 0.0;
              for (int gg = 0; gg < gpt; ++gg) s += acc[(g + gg) * 4 + a];
              const double rhs = dir == 0 ? b[pb[g] + a] : Y[r0[g] + a];
              T[r0[g] + a] = rhs - s;
            }
          }
        }
      }
      // phase B (writes applied after every warp has read: the device barrier)
      std::vector<Write> writes;
      for (int wi = 0; wi < nw; ++wi) {
        const std::vector<uint64_t> &tp = S.warp_tape[wi];
        size_t &p = pos[wi];
        const int nsmall = lo32(tp[p]), nlarge = hi32(tp[p]);
        ++p;
        for (int s = 0; s < nsmall; ++s) {
          const int r0 = lo32(tp[p]), w = hi32(tp[p]);
          const int direct = lo32(tp[p + 1]);
          p += 2;
          std::vector<int> pm(w);
          for (int i = 0; i < w; i += 2) {
            pm[i] = lo32(tp[p]);
            if (i + 1 < w) pm[i + 1] = hi32(tp[p]);
            ++p;
          }
          std::vector<double> tv(w), acc(w, 0.0);
          for (int i = 0; i < w; ++i) tv[i] = direct ? b[pm[i]] : T[r0 + i];
          for (int j = 0; j < w; ++j) {
            if (dir == 0) for (int i = j; i < w; ++i) acc[i] += bitsd(tp[p++]) * tv[j];
            else for (int i = 0; i <= j; ++i) acc[i] += bitsd(tp[p++]) * tv[j];
          }
          for (int i = 0; i < w; ++i) {
            writes.push_back({r0 + i, acc[i]});
            if (dir == 1) x[pm[i]] = acc[i];
          }
        }
        for (int l = 0; l < nlarge; ++l) {
          const int src0 = lo32(tp[p]), len = hi32(tp[p]);
          const int dest = lo32(tp[p + 1]), dperm = hi32(tp[p + 1]);
          p += 2;
          double acc = 0.0;
          for (int j = 0; j < len; ++j) acc += bitsd(tp[p++]) * T[src0 + j];
          writes.push_back({dest, acc});
          if (dir == 1) x[dperm] = acc;
        }
      }
      for (const Write &wv : writes) Y[wv.row] = wv.v;
    }
  }
}

}  // namespace dcora
