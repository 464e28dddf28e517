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
// Both the sparse part ("A": off-diagonal rows against the already solved
// levels, (d+1)x(d+1) tiles) and the dense part ("B": the explicitly inverted
// diagonal block, tiled the same way) are expressed as TILE TASKS (trsv.h) and
// executed by one device loop.  Blocks of one level of the block elimination
// tree are independent and the tree is O(log n) levels deep; a cluster barrier
// separates the slots A(0) B(0) A(1) B(1) ...
//
// Scheduling.  An "A" task accumulates into its destination, so its tiles need
// not run in the slot right before the destination's "B" task: a tile may run
// in any slot after its SOURCE block is solved.  Long destination rows (the
// top separators: hundreds of tiles per row, a handful of rows) are therefore
// cut into pieces that run early, in the slots where their sources become
// available; what remains on the critical path of an upper level is only the
// contribution of the level just below it.  Inside a slot the tasks are packed
// into rounds of 8 lane groups and dealt to the warps by longest-processing-
// time; the cost model (AgentTape::crit_iters) is what DCORA_TAPE_STATS prints.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstdlib>
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

struct Task {
  int type = 0;                // TapeType
  int r0 = 0, nrows = 0;       // destination rows (agent-local factor rows)
  std::vector<int> src;        // source row starts (agent-local)
  std::vector<double> val;     // dh*dh per source, [a][b]
  int gpt = 1;
  int ntiles() const { return (int)src.size(); }
};

struct Round {
  int type, gpt, niter;
  std::vector<int> tasks;      // indices into the slot's task list (<= 8/gpt)
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

int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s ? atoi(s) : dflt;
}

constexpr int ROUND_COST = 3;  // fixed cost of a round in units of one iteration

}  // namespace

int build_agent_tape(const HostFactor &F, int dh, int nwarps, int nwarps_total, int row_offset,
                     AgentTape *out) {
  const int k = F.k;
  const int bs = dh * dh;
  // long destination rows are cut into pieces of >= piece_min tiles that run as early as
  // their sources allow; rows of <= split_min tiles stay whole
  const int split_min = env_int("DCORA_TAPE_SPLIT", 24);
  const int piece_min = std::max(1, env_int("DCORA_TAPE_PIECE", 12));
  const bool timing = std::getenv("DCORA_CHOL_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[tape] %-26s %.3f s\n", what, std::chrono::duration<double>(now - t_prev).count());
    t_prev = now;
  };
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

  lap("partition + row lists");
  // ---- inverted diagonal blocks (shared by both directions)
  std::vector<std::vector<double>> Dinv(nb);
  for (int b = 0; b < nb; ++b) invert_diag_block(F, bp[b], bp[b + 1], &Dinv[b]);

  out->k = k;
  out->nblocks = nb;
  out->nwarps = nwarps_total;
  out->dh = dh;
  out->warp_tape.assign(nwarps_total, std::vector<uint64_t>());
  {
    // every tape starts with a round of TAPE_AHEAD zero records
    const int recw0 = tape_record_words(dh);
    for (std::vector<uint64_t> &tp : out->warp_tape) {
      tp.assign((size_t)TAPE_AHEAD * recw0, 0);
      for (int it = 0; it < TAPE_AHEAD; ++it) {
        uint64_t *rec = &tp[(size_t)it * recw0];
        for (int w = 0; w < 4; ++w) rec[12 + w] = pack2(row_offset, row_offset);
        if (it == 0) {
          rec[4] = pack2(TAPE_AHEAD, 1 | (TAPE_FWD << 8));
          for (int w = 0; w < 4; ++w) rec[5 + w] = pack2(-1, -1);
        }
      }
    }
  }
  out->stored_values = 0;
  out->crit_iters = out->crit_rounds = out->ideal_iters = 0;
  out->nslots = 0;
  const int R0 = row_offset;
  // first row of the (d+1)-row window a source column is read through: its pose
  // block, or -- for a unit-sphere / landmark row -- a window that stays inside
  // the agent's own rows (the other rows of the window meet zeros)
  auto src_start = [&](int c) {
    if (F.pose_start[c] >= 0) return F.pose_start[c];
    return std::max(0, std::min(c, k - dh));
  };
  const bool stats = getenv("DCORA_TAPE_STATS") != nullptr;

  lap("inverted diagonal blocks");
  // ---- levels of the block DAG, forward (L) and backward (L^T)
  std::vector<int> lvl[2];
  int nlev[2] = {0, 0};
  for (int dir = 0; dir < 2; ++dir) {
    const Csr &A = M[dir];
    lvl[dir].assign(nb, 0);
    for (int t = 0; t < nb; ++t) {
      const int b = dir == 0 ? t : nb - 1 - t;
      int lv = 0;
      for (int r = bp[b]; r < bp[b + 1]; ++r)
        for (int p = A.ptr[r]; p < A.ptr[r + 1]; ++p)
          lv = std::max(lv, lvl[dir][blk_of[A.col[p]]] + 1);
      lvl[dir][b] = lv;
      nlev[dir] = std::max(nlev[dir], lv + 1);
    }
    out->nlev[dir] = nlev[dir];
  }
  const int Lf = nlev[0], Lb = nlev[1];
  // global slots: forward level l (1..Lf-1) -> l - 1; the "mid" slot -> Lf - 1;
  // backward level m (1..Lb-1) -> Lf - 1 + m
  const int nslots_all = Lf + Lb - 1;
  std::vector<std::vector<Task>> slots((size_t)std::max(1, nslots_all));
  std::vector<int64_t> slot_load(slots.size(), 0);

  lap("levels");
  // ---- folded off-diagonal rows  Lp[i, B] = L[i, B] D_B  (dense from the block's first
  // column to the row's last non-zero column in B)
  struct RowPart { int blk, len; std::vector<double> v; };
  std::vector<std::vector<RowPart>> lp(k);
  std::vector<std::vector<std::pair<int, int>>> rows_of_block(nb);  // (row, index in lp[row])
  for (int i = 0; i < k; ++i) {
    const Csr &A = M[0];
    for (int p = A.ptr[i]; p < A.ptr[i + 1];) {
      const int b = blk_of[A.col[p]];
      int q = p;
      while (q < A.ptr[i + 1] && blk_of[A.col[q]] == b) ++q;
      const int s = bp[b], w = bp[b + 1] - s;
      const std::vector<double> &D = Dinv[b];
      RowPart rp;
      rp.blk = b;
      rp.len = A.col[q - 1] - s + 1;
      rp.v.assign(rp.len, 0.0);
      for (int t = p; t < q; ++t) {
        const int c = A.col[t] - s;
        const double lv = A.val[t];
        const double *dr = &D[(size_t)c * w];
        for (int j = 0; j <= c; ++j) rp.v[j] += lv * dr[j];
      }
      rows_of_block[b].push_back(std::make_pair(i, (int)lp[i].size()));
      lp[i].push_back(std::move(rp));
      p = q;
    }
  }

  // A destination's tiles, each with the first slot it may run in, become one task in the
  // destination's own (last) slot -- or, for long rows, several pieces that run as early as
  // their sources allow.
  auto emit_pieces = [&](int type, int r0, int nrows, const std::vector<int> &src,
                         const std::vector<double> &val, const std::vector<int> &avail,
                         int last) {
    std::vector<int> ord(src.size());
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return avail[x] < avail[y]; });
    auto emit = [&](size_t i0, size_t i1, int slot) {
      Task tk;
      tk.type = type; tk.r0 = r0; tk.nrows = nrows;
      for (size_t i = i0; i < i1; ++i) {
        tk.src.push_back(src[ord[i]]);
        tk.val.insert(tk.val.end(), val.begin() + (size_t)ord[i] * bs,
                      val.begin() + (size_t)(ord[i] + 1) * bs);
      }
      slot_load[slot] += tk.ntiles();
      slots[slot].push_back(std::move(tk));
    };
    if ((int)src.size() <= split_min) { emit(0, src.size(), last); return; }
    size_t i0 = 0;
    for (size_t i = 0; i < ord.size();) {
      size_t j = i;
      while (j < ord.size() && avail[ord[j]] == avail[ord[i]]) ++j;
      const int slot = avail[ord[i]];
      if (slot < last && (int)(j - i0) >= piece_min) { emit(i0, j, slot); i0 = j; }
      i = j;
    }
    if (i0 < ord.size()) emit(i0, ord.size(), last);
  };

  lap("folded rows Lp");
  // ---- forward tasks:  T[C rows] -= Lp[C rows, B] T[B]
  for (int r = 0; r < k;) {
    const int nrows = F.pose_start[r] == r ? dh : 1;
    const int lc = lvl[0][blk_of[r]];
    // blocks touched by the group's rows, with the longest row part
    std::vector<std::pair<int, int>> bl;  // (block, len)
    for (int a = 0; a < nrows; ++a)
      for (const RowPart &rp : lp[r + a]) bl.push_back(std::make_pair(rp.blk, rp.len));
    std::sort(bl.begin(), bl.end());
    std::vector<int> src, avail;
    std::vector<double> val;
    for (size_t i = 0; i < bl.size(); ++i) {
      if (i + 1 < bl.size() && bl[i + 1].first == bl[i].first) continue;  // keep the longest
      const int b = bl[i].first, len = bl[i].second, s = bp[b];
      for (int j = 0; j * dh < len; ++j) {
        const int c = s + j * dh;
        const int ps = std::max(0, std::min(c, k - dh));
        std::vector<double> tv((size_t)bs, 0.0);
        for (int a = 0; a < nrows; ++a)
          for (const RowPart &rp : lp[r + a]) {
            if (rp.blk != b) continue;
            for (int jj = j * dh; jj < std::min(rp.len, (j + 1) * dh); ++jj)
              tv[(size_t)a * dh + (s + jj - ps)] = rp.v[jj];
          }
        src.push_back(ps);
        val.insert(val.end(), tv.begin(), tv.end());
        avail.push_back(lvl[0][b]);
      }
    }
    if (!src.empty()) emit_pieces(TAPE_FWD, r, nrows, src, val, avail, lc - 1);
    r += nrows;
  }

  // ---- backward tasks:  X[B tile] -= Lp[C rows, B tile]^T X[C rows]
  for (int b = 0; b < nb; ++b) {
    const int s = bp[b], w = bp[b + 1] - s;
    const int nt = (w + dh - 1) / dh;
    const std::vector<std::pair<int, int>> &rows = rows_of_block[b];  // ascending rows
    for (int t = 0; t < nt; ++t) {
      const int nrows = std::min(dh, w - t * dh);
      std::vector<int> src, avail;
      std::vector<double> val;
      int cur_ps = -1;
      for (const std::pair<int, int> &ri : rows) {
        const int i = ri.first;
        const RowPart &rp = lp[i][ri.second];
        if (rp.len <= t * dh) continue;
        const int ps = src_start(i);
        if (ps != cur_ps) {
          src.push_back(ps);
          val.insert(val.end(), (size_t)bs, 0.0);
          avail.push_back(0);
          cur_ps = ps;
        }
        double *tv = &val[val.size() - bs];
        for (int a = 0; a < nrows; ++a) {
          const int jj = t * dh + a;
          if (jj < rp.len) tv[(size_t)a * dh + (i - ps)] = rp.v[jj];
        }
        avail.back() = std::max(avail.back(), Lf + lvl[1][blk_of[i]]);
      }
      if (!src.empty())
        emit_pieces(TAPE_BWD, s + t * dh, nrows, src, val, avail, Lf - 1 + lvl[1][b]);
    }
  }

  // ---- "mid" tasks:  X[B] = G_B T[B],  G_B = D_B^T D_B = (L_BB L_BB^T)^-1.  T[B] is final
  // after the block's forward slot and X[B] is first needed by the backward slots: the
  // task goes to the least loaded slot in between.
  {
    std::vector<int> border(nb);
    std::iota(border.begin(), border.end(), 0);
    std::stable_sort(border.begin(), border.end(), [&](int x, int y) {
      if (lvl[0][x] != lvl[0][y]) return lvl[0][x] > lvl[0][y];  // tightest windows first
      return bp[x + 1] - bp[x] > bp[y + 1] - bp[y];
    });
    for (int b : border) {
      const int s = bp[b], w = bp[b + 1] - s;
      const int nt = (w + dh - 1) / dh;
      const std::vector<double> &D = Dinv[b];
      std::vector<double> G((size_t)w * w, 0.0);
      for (int i = 0; i < w; ++i)
        for (int j = 0; j <= i; ++j) {
          double acc = 0.0;
          for (int kk = i; kk < w; ++kk) acc += D[(size_t)kk * w + i] * D[(size_t)kk * w + j];
          G[(size_t)i * w + j] = acc;
          G[(size_t)j * w + i] = acc;
        }
      int slot = lvl[0][b];
      for (int c = lvl[0][b]; c <= Lf - 1; ++c)
        if (slot_load[c] < slot_load[slot]) slot = c;
      for (int i = 0; i < nt; ++i) {
        Task tk;
        tk.type = TAPE_MID;
        tk.r0 = s + i * dh;
        tk.nrows = std::min(dh, w - i * dh);
        for (int j = 0; j < nt; ++j) {
          const int c = s + j * dh;
          const int ps = std::max(0, std::min(c, k - dh));
          std::vector<double> tv((size_t)bs, 0.0);
          for (int a = 0; a < tk.nrows; ++a)
            for (int jj = j * dh; jj < std::min(w, (j + 1) * dh); ++jj)
              tv[(size_t)a * dh + (s + jj - ps)] = G[(size_t)(i * dh + a) * w + jj];
          tk.src.push_back(ps);
          tk.val.insert(tk.val.end(), tv.begin(), tv.end());
        }
        slot_load[slot] += tk.ntiles();
        slots[slot].push_back(std::move(tk));
      }
    }
  }

  lap("tasks (fwd, bwd, mid)");
  // ---------------------------------------------------- pack every slot and emit
  for (size_t sl = 0; sl < slots.size(); ++sl) {
    std::vector<Task> &tasks = slots[sl];
    out->nslots++;   // (an empty slot still costs its barrier: there is none in practice)
    int64_t total = 0;
    for (const Task &tk : tasks) total += tk.ntiles();
    // groups per task: split tasks that are long compared with the ideal
    // per-group load of the slot
    const int64_t target =
        std::max<int64_t>(2, (total + (int64_t)nwarps * 8 - 1) / ((int64_t)nwarps * 8));
    for (Task &tk : tasks) {
      int g = 1;
      while (g < 8 && (int64_t)tk.ntiles() > target * g) g *= 2;
      tk.gpt = g;
    }
    std::vector<int> ord(tasks.size());
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) {
      if (tasks[x].type != tasks[y].type) return tasks[x].type < tasks[y].type;
      if (tasks[x].gpt != tasks[y].gpt) return tasks[x].gpt > tasks[y].gpt;
      return tasks[x].ntiles() > tasks[y].ntiles();
    });
    std::vector<Round> rounds;
    for (size_t i = 0; i < ord.size();) {
      Round rd;
      rd.type = tasks[ord[i]].type;
      rd.gpt = tasks[ord[i]].gpt;
      rd.niter = 0;
      const int tpw = 8 / rd.gpt;
      while (i < ord.size() && (int)rd.tasks.size() < tpw && tasks[ord[i]].gpt == rd.gpt &&
             tasks[ord[i]].type == rd.type) {
        const int nbk = tasks[ord[i]].ntiles();
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
        warp_rounds[lpt.take(rounds[i].niter + ROUND_COST)].push_back((int)i);
    }
    {
      int maxit = 0, maxrounds = 0, maxtask = 0;
      for (int wi = 0; wi < nwarps; ++wi) {
        int it = 0;
        for (int ri : warp_rounds[wi]) it += rounds[ri].niter;
        maxit = std::max(maxit, it);
        maxrounds = std::max(maxrounds, (int)warp_rounds[wi].size());
      }
      for (const Task &tk : tasks) maxtask = std::max(maxtask, tk.ntiles());
      const int64_t ideal = (total + (int64_t)nwarps * 8 - 1) / ((int64_t)nwarps * 8);
      out->crit_iters += maxit;
      out->crit_rounds += maxrounds;
      out->ideal_iters += ideal;
      if (stats)
        fprintf(stderr, "[tape] slot %2zu (%s): tasks %5zu tiles %7lld maxtask %4d rounds %4zu | "
                        "busiest warp: %3d iterations %2d rounds (ideal %lld)\n",
                sl, (int)sl < Lf - 1 ? "fwd" : (int)sl == Lf - 1 ? "mid" : "bwd", tasks.size(),
                (long long)total, maxtask, rounds.size(), maxit, maxrounds, (long long)ideal);
    }
    const int recw = tape_record_words(dh);
    auto emit_round = [&](std::vector<uint64_t> &tp, const Round *rd, bool last) {
      // rd == nullptr: a zero record (a warp without work in this slot still meets the barrier)
      const int niter = rd ? rd->niter : 1, gpt = rd ? rd->gpt : 1, type = rd ? rd->type : TAPE_FWD;
      for (int it = 0; it < niter; ++it) {
        const size_t w0 = tp.size();
        tp.resize(w0 + recw, 0);
        uint64_t *rec = &tp[w0];
        int c[8];
        const double *vp[8];
        for (int g = 0; g < 8; ++g) {
          const int ti = g / gpt, gg = g % gpt;
          c[g] = R0;  // padding: any valid row, zero tile
          vp[g] = nullptr;
          if (rd && ti < (int)rd->tasks.size()) {
            const Task &tk = tasks[rd->tasks[ti]];
            const int q = it * gpt + gg;
            if (q < tk.ntiles()) { c[g] = R0 + tk.src[q]; vp[g] = &tk.val[(size_t)q * bs]; }
          }
        }
        for (int g = 0; g < 8; g += 2) rec[12 + g / 2] = pack2(c[g], c[g + 1]);
        if (it == 0) {
          rec[4] = pack2(niter, gpt | (type << 8) | ((last ? 1 : 0) << 16));
          int dpk[8];
          for (int g = 0; g < 8; ++g) {
            const int ti = g / gpt;
            dpk[g] = -1;
            if (rd && ti < (int)rd->tasks.size()) {
              const Task &tk = tasks[rd->tasks[ti]];
              dpk[g] = ((R0 + tk.r0) << 3) | tk.nrows;
            }
          }
          for (int g = 0; g < 8; g += 2) rec[5 + g / 2] = pack2(dpk[g], dpk[g + 1]);
        }
        // lane (group g, column bb) owns column bb of its tile, rows 0..DH-1
        uint64_t *v = rec + TAPE_REC_HEADER;
        for (int g = 0; g < 8; ++g)
          for (int bb = 0; bb < dh; ++bb)
            for (int a = 0; a < dh; ++a)
              v[tape_value_word(dh, g, bb, a)] = dbits(vp[g] ? vp[g][a * dh + bb] : 0.0);
        if (rd) out->stored_values += 8 * bs;
      }
    };
    for (int wi = 0; wi < nwarps; ++wi) {
      std::vector<uint64_t> &tp = out->warp_tape[wi];
      const std::vector<int> &wr = warp_rounds[wi];
      if (wr.empty()) emit_round(tp, nullptr, true);
      for (size_t q = 0; q < wr.size(); ++q) emit_round(tp, &rounds[wr[q]], q + 1 == wr.size());
    }
    // warps of a larger launch cluster than this agent's own (batches mix
    // agent sizes; the work split of an agent depends on the agent alone, so
    // its result does not depend on what it is batched with): barriers only
    for (int wi = nwarps; wi < nwarps_total; ++wi) emit_round(out->warp_tape[wi], nullptr, true);
  }
  if (stats)
    fprintf(stderr, "[tape] k %d blocks %d levels %d/%d slots %d: busiest-warp iterations %lld "
                    "(ideal %lld), rounds %lld, stored values %lld (2 nnz(L) = %lld)\n",
            k, nb, Lf, Lb, out->nslots, (long long)out->crit_iters, (long long)out->ideal_iters,
            (long long)out->crit_rounds, (long long)out->stored_values,
            2 * (long long)F.Lx.size());
  // words 0..3 of record j <- own source rows (words 12..15) of record j + TAPE_AHEAD;
  // word 9 of every round's first record <- records left in its barrier slot
  {
    const int recw = tape_record_words(dh);
    for (std::vector<uint64_t> &tp : out->warp_tape) {
      const size_t n = tp.size() / recw;
      for (size_t j = 0; j < n; ++j)
        for (int w = 0; w < 4; ++w)
          tp[j * recw + w] = j + TAPE_AHEAD < n ? tp[(j + TAPE_AHEAD) * recw + 12 + w]
                                                : pack2(row_offset, row_offset);
      std::vector<size_t> firsts;   // first records of the rounds of the current slot
      for (size_t j = 0; j < n;) {
        const int niter = lo32(tp[j * recw + 4]);
        const bool last = ((hi32(tp[j * recw + 4]) >> 16) & 1) != 0;
        firsts.push_back(j);
        j += (size_t)niter;
        if (last) {
          for (size_t f : firsts) tp[f * recw + 9] = pack2((int)(j - f), 0);
          firsts.clear();
        }
      }
    }
  }
  lap("packing + emission");
  return DCORA_OK;
}

// CPU replay of the tapes, slot by slot (slots are separated by barriers on the
// device; inside a slot no task reads what another one writes).
void agent_tape_apply_host(const AgentTape &S, const HostFactor &F, int row_offset,
                           const double *b, double *x, double *T, double *X) {
  const int dh = S.dh, nw = S.nwarps, R0 = row_offset;
  for (int f = 0; f < S.k; ++f) T[R0 + f] = b[F.perm[f]];
  const int recw = tape_record_words(dh);
  std::vector<size_t> pos(nw, (size_t)TAPE_AHEAD * recw);   // past the zero round
  for (int sl = 0; sl < S.nslots; ++sl) {
    for (int wi = 0; wi < nw; ++wi) {
      const std::vector<uint64_t> &tp = S.warp_tape[wi];
      size_t &p = pos[wi];
      for (bool last = false; !last;) {
        const uint64_t *h = &tp[p];
        const int niter = lo32(h[4]), cw = hi32(h[4]);
        const int gpt = cw & 0xff, type = (cw >> 8) & 0xff;
        last = ((cw >> 16) & 1) != 0;
        int r0[8], nrows[8];
        for (int g = 0; g < 8; ++g) {
          const int d = (g & 1) ? hi32(h[5 + g / 2]) : lo32(h[5 + g / 2]);
          r0[g] = d >> 3;
          nrows[g] = d < 0 ? 0 : (d & 7);
        }
        const double *src = type == TAPE_BWD ? X : T;
        double *dst = type == TAPE_FWD ? T : X;
        const bool accumulate = type != TAPE_MID;
        std::vector<double> acc(8 * 4, 0.0);
        for (int it = 0; it < niter; ++it) {
          const uint64_t *rec = &tp[p];
          int c[8];
          for (int g = 0; g < 8; g += 2) { c[g] = lo32(rec[12 + g / 2]); c[g + 1] = hi32(rec[12 + g / 2]); }
          const uint64_t *v = rec + TAPE_REC_HEADER;
          for (int g = 0; g < 8; ++g)
            for (int a = 0; a < dh; ++a)
              for (int bb = 0; bb < dh; ++bb)
                acc[g * 4 + a] += bitsd(v[tape_value_word(dh, g, bb, a)]) * src[c[g] + bb];
          p += recw;
        }
        for (int g = 0; g < 8; g += gpt) {
          if (r0[g] < 0) continue;
          for (int a = 0; a < nrows[g]; ++a) {
            double s = 0.0;
            for (int gg = 0; gg < gpt; ++gg) s += acc[(g + gg) * 4 + a];
            dst[r0[g] + a] = accumulate ? dst[r0[g] + a] - s : s;
          }
        }
      }
    }
  }
  for (int f = 0; f < S.k; ++f) x[F.perm[f]] = X[R0 + f];
}

}  // namespace dcora
