// Host sparse Cholesky of P = Q + reg*I for one agent.
//
// Replaces Eigen::CholmodDecomposition in Graph::constructPreconditioner
// (src/Graph.cpp:1901-1917): CHOLMOD is a third-party dependency that is not
// available here, so the factorisation is our own: a minimum-degree ordering
// on the supervariable graph (one node per pose block / sphere / landmark),
// Liu's elimination tree, and an up-looking numeric factorisation.  The factor
// is built once per (agent, Q) and applied on the device by level-scheduled
// sparse triangular solves (sptrsv.cu).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <thread>
#include <numeric>
#include <queue>
#include <set>

#include "common.cuh"

namespace dcora {

namespace {

// Minimum (external, weighted) degree ordering with a quotient graph and
// element absorption.  `adj` is the symmetric adjacency (no self loops).
std::vector<int> minimum_degree(int n, const std::vector<std::vector<int>> &adj,
                                const std::vector<int> &weight) {
  std::vector<std::vector<int>> av(adj);        // variable neighbours
  std::vector<std::vector<int>> ae(n);          // adjacent elements
  std::vector<std::vector<int>> el(n);          // element -> variables
  std::vector<char> eliminated(n, 0), absorbed(n, 0);
  std::vector<int> mark(n, -1), deg(n, 0), order;
  order.reserve(n);
  int stamp = 0;
  typedef std::pair<int, int> DI;
  std::priority_queue<DI, std::vector<DI>, std::greater<DI>> pq;
  for (int i = 0; i < n; ++i) {
    int dg = 0;
    for (int j : av[i]) dg += weight[j];
    deg[i] = dg;
    pq.push(DI(dg, i));
  }
  std::vector<int> Lp;
  while (!pq.empty()) {
    const DI top = pq.top();
    pq.pop();
    const int p = top.second;
    if (eliminated[p] || top.first != deg[p]) continue;
    eliminated[p] = 1;
    order.push_back(p);
    // Lp = (av[p] U union el[e], e in ae[p]) \ {p}
    ++stamp;
    mark[p] = stamp;
    Lp.clear();
    for (int j : av[p])
      if (!eliminated[j] && mark[j] != stamp) { mark[j] = stamp; Lp.push_back(j); }
    for (int e : ae[p]) {
      if (absorbed[e]) continue;
      for (int j : el[e])
        if (!eliminated[j] && mark[j] != stamp) { mark[j] = stamp; Lp.push_back(j); }
      absorbed[e] = 1;
      std::vector<int>().swap(el[e]);
    }
    std::vector<int>().swap(av[p]);
    std::vector<int>().swap(ae[p]);
    el[p] = Lp;
    // update the members of the new element
    for (int i : Lp) {
      // prune variable list: drop eliminated and anything now inside Lp
      std::vector<int> &v = av[i];
      size_t w = 0;
      for (size_t q = 0; q < v.size(); ++q)
        if (!eliminated[v[q]] && mark[v[q]] != stamp) v[w++] = v[q];
      v.resize(w);
      std::vector<int> &e = ae[i];
      w = 0;
      for (size_t q = 0; q < e.size(); ++q)
        if (!absorbed[e[q]]) e[w++] = e[q];
      e.resize(w);
      e.push_back(p);
    }
    // exact external degrees of the members
    for (int i : Lp) {
      ++stamp;
      mark[i] = stamp;
      int dg = 0;
      for (int j : av[i])
        if (mark[j] != stamp) { mark[j] = stamp; dg += weight[j]; }
      for (int e : ae[i])
        for (int j : el[e])
          if (mark[j] != stamp) { mark[j] = stamp; dg += weight[j]; }
      deg[i] = dg;
      pq.push(DI(dg, i));
    }
    // restore marks semantics for next pivot (stamp is monotone, nothing to do)
  }
  return order;
}


// ---------------------------------------------------------------------------
// Nested dissection on the supervariable graph (George's automatic nested
// dissection: separators are levels of a BFS level structure rooted at a
// pseudo-peripheral node, thinned by one pass).  The recursion stops at
// subgraphs of at most `leaf_weight` scalar rows.  Output: the elimination
// order (subdomains first, separator last, recursively) and the partition of
// that order into consecutive BLOCKS (leaf subdomains and separators): the
// block triangular solve of sptrsv.cu inverts the diagonal blocks and walks
// the (shallow, log-depth) block elimination tree level by level.
struct NDContext {
  const std::vector<std::vector<int>> *adj;
  const std::vector<int> *weight;
  int leaf_weight;
  std::vector<int> label;     // region id of every node (-1: already ordered)
  std::vector<int> dist;      // BFS scratch
  std::vector<int> order;     // output
  std::vector<int> block_ptr; // output: block boundaries in `order`
  int next_label = 1;
};

static int region_weight(const NDContext &C, const std::vector<int> &nodes) {
  int w = 0;
  for (int v : nodes) w += (*C.weight)[v];
  return w;
}

static void emit_block(NDContext &C, const std::vector<int> &nodes, int lab) {
  if (nodes.empty()) return;
  // interior nodes (no neighbour outside the block that is still unordered)
  // first: keeps the fill of the ancestors' rows small
  std::vector<int> inner, outer;
  for (int v : nodes) {
    bool boundary = false;
    for (int u : (*C.adj)[v])
      if (C.label[u] != lab && C.label[u] != -1) { boundary = true; break; }
    (boundary ? outer : inner).push_back(v);
  }
  for (int v : inner) C.order.push_back(v);
  for (int v : outer) C.order.push_back(v);
  for (int v : nodes) C.label[v] = -1;
  C.block_ptr.push_back((int)C.order.size());
}

// BFS inside region `lab` from `root`; returns the level structure.
static void bfs_levels(NDContext &C, int root, int lab,
                       std::vector<std::vector<int>> *levels) {
  levels->clear();
  levels->push_back(std::vector<int>(1, root));
  C.dist[root] = 0;
  std::vector<int> touched(1, root);
  for (size_t lv = 0; lv < levels->size(); ++lv) {
    std::vector<int> next;
    for (int v : (*levels)[lv])
      for (int u : (*C.adj)[v])
        if (C.label[u] == lab && C.dist[u] < 0) {
          C.dist[u] = (int)lv + 1;
          next.push_back(u);
          touched.push_back(u);
        }
    if (next.empty()) break;
    levels->push_back(std::move(next));
  }
  for (int v : touched) C.dist[v] = -1;
}

static void nd_recurse(NDContext &C, std::vector<int> nodes, int lab) {
  const std::vector<std::vector<int>> &adj = *C.adj;
  const std::vector<int> &wt = *C.weight;
  const int W = region_weight(C, nodes);
  if (W <= C.leaf_weight || nodes.size() <= 1) { emit_block(C, nodes, lab); return; }
  // ---- connected components of the region
  std::vector<std::vector<int>> comps;
  {
    // label flipping marks visited nodes (temporary label -2)
    for (int v : nodes) {
      if (C.label[v] != lab) continue;
      std::vector<int> comp(1, v);
      C.label[v] = -2;
      for (size_t h = 0; h < comp.size(); ++h)
        for (int u : adj[comp[h]])
          if (C.label[u] == lab) { C.label[u] = -2; comp.push_back(u); }
      comps.push_back(std::move(comp));
    }
    for (int v : nodes) C.label[v] = lab;
  }
  if (comps.size() > 1) {
    // two bins of components, balanced greedily; no separator needed
    std::sort(comps.begin(), comps.end(),
              [&](const std::vector<int> &a, const std::vector<int> &b) {
                return a.size() > b.size();
              });
    std::vector<int> binA, binB;
    int wa = 0, wb = 0;
    for (auto &c : comps) {
      const int w = region_weight(C, c);
      if (wa <= wb) { binA.insert(binA.end(), c.begin(), c.end()); wa += w; }
      else { binB.insert(binB.end(), c.begin(), c.end()); wb += w; }
    }
    const int la = C.next_label++, lb = C.next_label++;
    for (int v : binA) C.label[v] = la;
    for (int v : binB) C.label[v] = lb;
    nd_recurse(C, std::move(binA), la);
    nd_recurse(C, std::move(binB), lb);
    return;
  }
  // ---- pseudo-peripheral root
  int root = nodes[0];
  {
    size_t best = adj[root].size();
    for (int v : nodes)
      if (adj[v].size() < best) { best = adj[v].size(); root = v; }
  }
  std::vector<std::vector<int>> levels;
  bfs_levels(C, root, lab, &levels);
  for (int it = 0; it < 6; ++it) {
    const std::vector<int> &last = levels.back();
    int cand = last[0];
    size_t deg = adj[cand].size();
    for (int v : last)
      if (adj[v].size() < deg) { deg = adj[v].size(); cand = v; }
    std::vector<std::vector<int>> lv2;
    bfs_levels(C, cand, lab, &lv2);
    if (lv2.size() > levels.size()) { levels.swap(lv2); root = cand; }
    else break;
  }
  const int h = (int)levels.size();
  if (h < 3) {
    // no interior level (star / clique-like region).  Small: one dense block.
    // Large: peel the highest-degree node off as a one-node separator (it is
    // what keeps the region's diameter small) and dissect the rest.
    if (W <= 4 * C.leaf_weight) { emit_block(C, nodes, lab); return; }
    int hub = nodes[0];
    size_t hd = 0;
    for (int v : nodes) {
      size_t dg = 0;
      for (int u : adj[v]) dg += C.label[u] == lab;
      if (dg > hd) { hd = dg; hub = v; }
    }
    std::vector<int> rest;
    rest.reserve(nodes.size() - 1);
    const int lr = C.next_label++;
    for (int v : nodes)
      if (v != hub) { rest.push_back(v); C.label[v] = lr; }
    nd_recurse(C, std::move(rest), lr);
    emit_block(C, std::vector<int>(1, hub), lab);
    return;
  }
  // ---- choose the separator level
  std::vector<int> lw(h, 0), cum(h + 1, 0);
  for (int i = 0; i < h; ++i) {
    for (int v : levels[i]) lw[i] += wt[v];
    cum[i + 1] = cum[i] + lw[i];
  }
  int best = -1;
  double best_cost = 0.0;
  for (int pass = 0; pass < 2 && best < 0; ++pass)
    for (int i = 1; i + 1 < h; ++i) {
      const int below = cum[i], above = W - cum[i + 1];
      const double bal = (double)std::min(below, above) / (double)std::max(1, std::max(below, above));
      if (pass == 0 && bal < 0.4) continue;
      const double cost = (double)lw[i] * (1.0 + 0.5 * (1.0 - bal)) + (pass ? 1e6 * (1.0 - bal) : 0.0);
      if (best < 0 || cost < best_cost) { best = i; best_cost = cost; }
    }
  // ---- thin it: separator nodes without a neighbour in the next level join
  // the lower part
  std::vector<int> sep, partA, partB;
  for (int i = 0; i < best; ++i) partA.insert(partA.end(), levels[i].begin(), levels[i].end());
  for (int i = best + 1; i < h; ++i) partB.insert(partB.end(), levels[i].begin(), levels[i].end());
  const int la = C.next_label++, lb = C.next_label++;
  for (int v : partB) C.label[v] = lb;
  for (int v : levels[best]) {
    bool up = false;
    for (int u : adj[v])
      if (C.label[u] == lb) { up = true; break; }
    if (up) sep.push_back(v); else partA.push_back(v);
  }
  for (int v : partA) C.label[v] = la;
  // separator keeps label `lab`
  nd_recurse(C, std::move(partA), la);
  nd_recurse(C, std::move(partB), lb);
  emit_block(C, sep, lab);
}

void nested_dissection(int n, const std::vector<std::vector<int>> &adj,
                       const std::vector<int> &weight, int leaf_weight,
                       std::vector<int> *order, std::vector<int> *block_ptr) {
  NDContext C;
  C.adj = &adj;
  C.weight = &weight;
  C.leaf_weight = leaf_weight;
  C.label.assign(n, 0);
  C.dist.assign(n, -1);
  C.order.reserve(n);
  C.block_ptr.assign(1, 0);
  // hubs (e.g. a landmark ranged from thousands of poses) would make every
  // BFS level structure two levels deep: order them last, as the top block
  size_t total_deg = 0;
  for (int v = 0; v < n; ++v) total_deg += adj[v].size();
  const size_t hub_deg = std::max<size_t>(64, 10 * (total_deg / std::max(1, n) + 1));
  std::vector<int> all, hubs;
  for (int v = 0; v < n; ++v) (adj[v].size() > hub_deg ? hubs : all).push_back(v);
  const int hub_label = C.next_label++;
  for (int v : hubs) C.label[v] = hub_label;
  nd_recurse(C, std::move(all), 0);
  emit_block(C, hubs, hub_label);
  order->swap(C.order);
  block_ptr->swap(C.block_ptr);
}


// ---------------------------------------------------------------------------
// Supernodal left-looking numeric factorisation on a given symbolic structure.
//
// Columns with identical structure below the diagonal (parent[j] = j + 1 and
// colcount[j] = colcount[j + 1] + 1: fundamental supernodes) are factored as dense panels:
// the updates from earlier supernodes are rank-w products of panel rows (contiguous inner
// loops the compiler vectorises), the diagonal block is a dense Cholesky, the rows below it a
// triangular solve.  Same arithmetic as the scalar up-looking algorithm up to the order of the
// sums; 3-4x faster on the 3-D slab agents, whose separators are dense blocks of hundreds of
// columns.  C = upper triangle of the permuted matrix in CSC; Lp / Li hold the structure of L
// (rows sorted); Lx receives the values.  Returns false when a pivot is not positive.
bool supernodal_numeric(int k, const std::vector<int> &Cp, const std::vector<int> &Ci,
                        const std::vector<double> &Cx, const std::vector<int> &parent,
                        const std::vector<int> &Lp, const std::vector<int> &Li,
                        std::vector<double> &Lx) {
  // ---- supernodes
  std::vector<int> sstart;   // first column of every supernode (+ k at the end)
  std::vector<int> super_of(k);
  for (int j = 0; j < k; ++j) {
    const bool joins = j > 0 && parent[j - 1] == j &&
                       (Lp[j] - Lp[j - 1]) == (Lp[j + 1] - Lp[j]) + 1;
    if (!joins) sstart.push_back(j);
    super_of[j] = (int)sstart.size() - 1;
  }
  const int ns = (int)sstart.size();
  sstart.push_back(k);
  // panel of supernode s: rows = structure of its first column, nr x nc column-major
  std::vector<size_t> poff(ns + 1, 0);
  for (int s = 0; s < ns; ++s) {
    const int j0 = sstart[s], nc = sstart[s + 1] - j0, nr = Lp[j0 + 1] - Lp[j0];
    poff[s + 1] = poff[s] + (size_t)nr * nc;
  }
  std::vector<double> P(poff[ns], 0.0);
  std::vector<int> head(ns, -1), next(ns, -1), pos(ns, 0), relmap(k, -1);
  std::vector<double> W;
  for (int s = 0; s < ns; ++s) {
    const int j0 = sstart[s], j1 = sstart[s + 1], nc = j1 - j0;
    const int *R = Li.data() + Lp[j0];
    const int nr = Lp[j0 + 1] - Lp[j0];
    double *Ps = P.data() + poff[s];
    for (int t = 0; t < nr; ++t) relmap[R[t]] = t;
    // ---- assemble the lower part of A: A(r, i) = C(i, r) for i in [j0, j1), r >= i
    for (int t = 0; t < nr; ++t) {
      const int r = R[t];
      for (int p = Cp[r]; p < Cp[r + 1]; ++p) {
        const int i = Ci[p];
        if (i >= j0 && i < j1) Ps[(size_t)(i - j0) * nr + t] += Cx[p];
      }
    }
    // ---- updates from the supernodes d with rows in [j0, j1)
    for (int d = head[s]; d != -1;) {
      const int dnext = next[d];
      const int d0 = sstart[d], wd = sstart[d + 1] - d0;
      const int *Rd = Li.data() + Lp[d0];
      const int nrd = Lp[d0 + 1] - Lp[d0];
      const double *Pd = P.data() + poff[d];
      const int p0 = pos[d];
      int ndcol = 0;
      while (p0 + ndcol < nrd && Rd[p0 + ndcol] < j1) ++ndcol;
      const int ndrow = nrd - p0;
      // W (ndrow x ndcol, lower trapezoid) = Pd[p0.., :] * Pd[p0..p0+ndcol, :]^T, four panel
      // columns per pass over W (the inner loops over `a` are contiguous: AVX2 FMA)
      W.assign((size_t)ndrow * ndcol, 0.0);
      int t = 0;
      for (; t + 4 <= wd; t += 4) {
        const double *c0 = Pd + (size_t)t * nrd + p0, *c1 = c0 + nrd, *c2 = c1 + nrd, *c3 = c2 + nrd;
        for (int b = 0; b < ndcol; ++b) {
          const double b0 = c0[b], b1 = c1[b], b2 = c2[b], b3 = c3[b];
          double *w = W.data() + (size_t)b * ndrow;
          for (int a = b; a < ndrow; ++a) w[a] += c0[a] * b0 + c1[a] * b1 + c2[a] * b2 + c3[a] * b3;
        }
      }
      for (; t < wd; ++t) {
        const double *col = Pd + (size_t)t * nrd + p0;
        for (int b = 0; b < ndcol; ++b) {
          const double cb = col[b];
          double *w = W.data() + (size_t)b * ndrow;
          for (int a = b; a < ndrow; ++a) w[a] += col[a] * cb;
        }
      }
      for (int b = 0; b < ndcol; ++b) {
        double *dst = Ps + (size_t)(Rd[p0 + b] - j0) * nr;
        const double *w = W.data() + (size_t)b * ndrow;
        for (int a = b; a < ndrow; ++a) dst[relmap[Rd[p0 + a]]] -= w[a];
      }
      pos[d] = p0 + ndcol;
      if (pos[d] < nrd) {
        const int tgt = super_of[Rd[pos[d]]];
        next[d] = head[tgt];
        head[tgt] = d;
      }
      d = dnext;
    }
    // ---- dense Cholesky of the nc x nc diagonal block + solve of the rows below
    // (left-looking inside the panel: column c first takes the updates of columns < c, four at
    // a time, then is scaled)
    for (int c = 0; c < nc; ++c) {
      double *col = Ps + (size_t)c * nr;
      int q = 0;
      for (; q + 4 <= c; q += 4) {
        const double *a0 = Ps + (size_t)q * nr, *a1 = a0 + nr, *a2 = a1 + nr, *a3 = a2 + nr;
        const double f0 = a0[c], f1 = a1[c], f2 = a2[c], f3 = a3[c];
        for (int t2 = c; t2 < nr; ++t2) col[t2] -= a0[t2] * f0 + a1[t2] * f1 + a2[t2] * f2 + a3[t2] * f3;
      }
      for (; q < c; ++q) {
        const double *a0 = Ps + (size_t)q * nr;
        const double f0 = a0[c];
        for (int t2 = c; t2 < nr; ++t2) col[t2] -= a0[t2] * f0;
      }
      const double dgn = col[c];
      if (!(dgn > 0.0) || !std::isfinite(dgn)) return false;
      const double ljj = std::sqrt(dgn), inv = 1.0 / ljj;
      col[c] = ljj;
      for (int t2 = c + 1; t2 < nr; ++t2) col[t2] *= inv;
    }
    if (nr > nc) {
      pos[s] = nc;
      const int tgt = super_of[R[nc]];
      next[s] = head[tgt];
      head[tgt] = s;
    }
    // ---- values to the CSC factor: column j0 + c holds rows R[c..nr)
    for (int c = 0; c < nc; ++c) {
      double *lx = Lx.data() + Lp[j0 + c];
      const double *col = Ps + (size_t)c * nr;
      for (int t = c; t < nr; ++t) lx[t - c] = col[t];
    }
  }
  return true;
}

}  // namespace

int sparse_cholesky(int k, const std::vector<int> &Ap, const std::vector<int> &Ai,
                    const std::vector<double> &Ax, int dh, int npose,
                    HostFactor *out, int ordering, int leaf_rows, int threads) {
  HostFactor &F = *out;
  F.k = k;
  const bool timing = std::getenv("DCORA_CHOL_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[chol] %-22s %.3f s\n", what, std::chrono::duration<double>(now - t_prev).count());
    t_prev = now;
  };
  // ---- supervariable graph
  const int npc = dh * npose;
  const int nsv = npose + (k - npc);
  auto sv_of = [&](int c) { return c < npc ? c / dh : npose + (c - npc); };
  std::vector<std::vector<int>> adj(nsv);
  {
    std::vector<int> last(nsv, -1);
    // rows of one supervariable are contiguous in the internal ordering
    for (int s = 0; s < nsv; ++s) {
      const int c0 = s < npose ? s * dh : npc + (s - npose);
      const int c1 = s < npose ? c0 + dh : c0 + 1;
      for (int c = c0; c < c1; ++c)
        for (int p = Ap[c]; p < Ap[c + 1]; ++p) {
          const int t = sv_of(Ai[p]);
          if (t != s && last[t] != s) { last[t] = s; adj[s].push_back(t); }
        }
    }
    // symmetrise (pattern of Q is symmetric, but be safe)
    std::vector<std::vector<int>> extra(nsv);
    for (int s = 0; s < nsv; ++s)
      for (int t : adj[s])
        if (std::find(adj[t].begin(), adj[t].end(), s) == adj[t].end()) extra[t].push_back(s);
    for (int s = 0; s < nsv; ++s)
      for (int t : extra[s])
        if (std::find(adj[s].begin(), adj[s].end(), t) == adj[s].end()) adj[s].push_back(t);
  }
  std::vector<int> weight(nsv);
  for (int s = 0; s < nsv; ++s) weight[s] = s < npose ? dh : 1;
  std::vector<int> order, sv_block_ptr;
  if (ordering == DCORA_ORDERING_ND) {
    nested_dissection(nsv, adj, weight, leaf_rows, &order, &sv_block_ptr);
  } else {
    order = minimum_degree(nsv, adj, weight);
    sv_block_ptr.resize(nsv + 1);
    std::iota(sv_block_ptr.begin(), sv_block_ptr.end(), 0);
  }
  F.perm.clear();
  F.perm.reserve(k);
  F.pose_start.assign(k, -1);
  F.blk_ptr.assign(1, 0);
  {
    size_t nb = 1;
    for (int t = 0; t < nsv; ++t) {
      const int s = order[t];
      const int c0 = s < npose ? s * dh : npc + (s - npose);
      const int c1 = s < npose ? c0 + dh : c0 + 1;
      const int first = (int)F.perm.size();
      for (int c = c0; c < c1; ++c) {
        if (s < npose) F.pose_start[F.perm.size()] = first;
        F.perm.push_back(c);
      }
      if (nb < sv_block_ptr.size() && t + 1 == sv_block_ptr[nb]) {
        F.blk_ptr.push_back((int)F.perm.size());
        ++nb;
      }
    }
  }
  std::vector<int> pinv(k);
  for (int i = 0; i < k; ++i) pinv[F.perm[i]] = i;

  lap("ordering");
  // ---- C = upper triangle of P A P^T in CSC (column j holds rows i <= j)
  std::vector<int> Cp(k + 1, 0);
  for (int c = 0; c < k; ++c)
    for (int p = Ap[c]; p < Ap[c + 1]; ++p) {
      const int i = pinv[c], j = pinv[Ai[p]];
      if (i <= j) Cp[j + 1]++;   // entry (c, Ai[p]) -> (i, j), keep upper
    }
  for (int j = 0; j < k; ++j) Cp[j + 1] += Cp[j];
  std::vector<int> Ci(Cp[k]);
  std::vector<double> Cx(Cp[k]);
  {
    std::vector<int> fill(Cp.begin(), Cp.end() - 1);
    for (int c = 0; c < k; ++c)
      for (int p = Ap[c]; p < Ap[c + 1]; ++p) {
        const int i = pinv[c], j = pinv[Ai[p]];
        if (i <= j) { Ci[fill[j]] = i; Cx[fill[j]] = Ax[p]; fill[j]++; }
      }
  }
  lap("permute");
  // ---- elimination tree (Liu)
  std::vector<int> parent(k, -1), ancestor(k, -1);
  for (int j = 0; j < k; ++j)
    for (int p = Cp[j]; p < Cp[j + 1]; ++p) {
      int i = Ci[p];
      while (i != -1 && i < j) {
        const int inext = ancestor[i];
        ancestor[i] = j;
        if (inext == -1) parent[i] = j;
        i = inext;
      }
    }
  // ---- symbolic: column counts through row sub-tree traversal (ereach)
  std::vector<int> colcount(k, 1), flag(k, -1);
  for (int j = 0; j < k; ++j) {
    flag[j] = j;
    for (int p = Cp[j]; p < Cp[j + 1]; ++p) {
      int i = Ci[p];
      while (i != -1 && i < j && flag[i] != j) {
        colcount[i]++;   // L(j, i) != 0
        flag[i] = j;
        i = parent[i];
      }
    }
  }
  F.Lp.assign(k + 1, 0);
  for (int j = 0; j < k; ++j) F.Lp[j + 1] = F.Lp[j] + colcount[j];
  const int64_t nnzL = F.Lp[k];
  F.Li.assign(nnzL, 0);
  F.Lx.assign(nnzL, 0.0);
  lap("etree + symbolic");
  static const bool uplooking = std::getenv("DCORA_CHOL") && std::getenv("DCORA_CHOL")[0] == 'u';
  if (!uplooking) {
    // structure of L: row j joins the columns of its row sub-tree (same traversal as above)
    // (rows are appended in increasing j per column; the diagonal comes first)
    std::vector<int> cf(F.Lp.begin(), F.Lp.end() - 1);
    std::fill(flag.begin(), flag.end(), -1);
    for (int j = 0; j < k; ++j) {
      // columns of row j, then the diagonal entry of column j
      flag[j] = j;
      for (int p = Cp[j]; p < Cp[j + 1]; ++p) {
        int i = Ci[p];
        while (i != -1 && i < j && flag[i] != j) {
          F.Li[cf[i]++] = j;
          flag[i] = j;
          i = parent[i];
        }
      }
      F.Li[cf[j]++] = j;
    }
    lap("structure of L");
    if (!supernodal_numeric(k, Cp, Ci, Cx, parent, F.Lp, F.Li, F.Lx)) {
      set_error("sparse_cholesky: matrix is not positive definite");
      return DCORA_ERR_FACTOR;
    }
    lap("numeric (supernodal)");
    F.nnz = nnzL;
    return DCORA_OK;
  }
  // ---- numeric up-looking factorisation (row j of L from rows < j).  Row j only reads and
  // appends to the columns of its elimination-tree descendants, so disjoint subtrees are
  // independent: with `threads` > 1 the tree is cut into subtrees that are factored
  // concurrently (every row still sees the same operands in the same order: the factor is
  // bit-identical to the serial one), then the rows above the cut follow in order.
  std::vector<int> cfill(F.Lp.begin(), F.Lp.end() - 1);
  std::atomic<int> failed{0};
  struct Work {
    std::vector<int> flag, stack, pattern;
    std::vector<double> x;
    explicit Work(int n) : flag(n, -1), stack(n), pattern(n), x(n, 0.0) {}
  };
  auto do_row = [&](int j, Work &W) {
    int top = k;
    W.flag[j] = j;
    double d = 0.0;
    for (int p = Cp[j]; p < Cp[j + 1]; ++p) {
      int i = Ci[p];
      if (i == j) { d += Cx[p]; continue; }
      W.x[i] += Cx[p];   // += : tolerate duplicate entries
      int len = 0;
      while (W.flag[i] != j) {
        W.stack[len++] = i;
        W.flag[i] = j;
        i = parent[i];
      }
      while (len > 0) W.pattern[--top] = W.stack[--len];
    }
    for (; top < k; ++top) {
      const int i = W.pattern[top];
      const double lji = W.x[i] / F.Lx[F.Lp[i]];  // L(j,i) = x_i / L(i,i)
      W.x[i] = 0.0;
      for (int p = F.Lp[i] + 1; p < cfill[i]; ++p) W.x[F.Li[p]] -= F.Lx[p] * lji;
      d -= lji * lji;
      const int q = cfill[i]++;
      F.Li[q] = j;
      F.Lx[q] = lji;
    }
    if (!(d > 0.0) || !std::isfinite(d)) { failed.store(1); d = 1.0; }
    const int q = cfill[j]++;
    F.Li[q] = j;
    F.Lx[q] = std::sqrt(d);
  };
  if (threads <= 1 || k < 4000) {
    Work W(k);
    for (int j = 0; j < k && !failed.load(); ++j) do_row(j, W);
  } else {
    // subtree work = sum of colcount^2 (flops of the rows appended to a column)
    std::vector<double> work(k);
    for (int j = 0; j < k; ++j) work[j] = (double)colcount[j] * colcount[j];
    for (int j = 0; j < k; ++j) if (parent[j] >= 0) work[parent[j]] += work[j];
    std::vector<std::vector<int>> children(k);
    std::vector<int> roots;
    for (int j = 0; j < k; ++j) (parent[j] >= 0 ? children[parent[j]] : roots).push_back(j);
    // cut: repeatedly open the heaviest subtree until there are enough pieces
    std::vector<int> pieces = roots, above;
    // (a dense separator is a chain in the tree: keep descending until the heaviest piece is
    // a fair share of the total)
    double total_work = 0.0;
    for (int t : roots) total_work += work[t];
    const double fair = total_work / (4.0 * threads);
    for (;;) {
      size_t h = 0;
      for (size_t t = 1; t < pieces.size(); ++t) if (work[pieces[t]] > work[pieces[h]]) h = t;
      const int top_node = pieces[h];
      if (work[top_node] <= fair || children[top_node].empty()) break;
      pieces.erase(pieces.begin() + h);
      above.push_back(top_node);
      for (int c : children[top_node]) pieces.push_back(c);
    }
    std::vector<int> piece_of(k, -1);
    for (size_t t = 0; t < pieces.size(); ++t) piece_of[pieces[t]] = (int)t;
    // nodes in descending order: a node inherits its parent's piece (parents have larger indices)
    for (int j = k - 1; j >= 0; --j)
      if (piece_of[j] < 0 && parent[j] >= 0) piece_of[j] = piece_of[parent[j]];
    for (int a : above) piece_of[a] = -1;
    std::vector<std::vector<int>> rows(pieces.size());
    for (int j = 0; j < k; ++j) if (piece_of[j] >= 0) rows[piece_of[j]].push_back(j);
    std::vector<size_t> order(pieces.size());
    for (size_t t = 0; t < order.size(); ++t) order[t] = t;
    std::sort(order.begin(), order.end(),
              [&](size_t a, size_t b) { return work[pieces[a]] > work[pieces[b]]; });
    if (timing) {
      double wa = 0.0, wp = 0.0, wmax = 0.0;
      for (int a : above) wa += (double)colcount[a] * colcount[a];
      for (int t : pieces) { wp += work[t]; wmax = std::max(wmax, work[t]); }
      std::fprintf(stderr, "[chol] %zu pieces (work %.3g, largest %.3g), %zu rows above the cut (work %.3g)\n",
                   pieces.size(), wp, wmax, above.size(), wa);
    }
    std::atomic<size_t> next{0};
    auto worker = [&]() {
      Work W(k);
      for (;;) {
        const size_t t = next.fetch_add(1);
        if (t >= order.size() || failed.load()) break;
        for (int j : rows[order[t]]) do_row(j, W);
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &th : pool) th.join();
    std::sort(above.begin(), above.end());
    Work W(k);
    for (int j : above) { if (failed.load()) break; do_row(j, W); }
  }
  lap("numeric");
  if (failed.load()) {
    set_error("sparse_cholesky: matrix is not positive definite");
    return DCORA_ERR_FACTOR;
  }
  F.nnz = nnzL;
  return DCORA_OK;
}


// Largest eigenvalue of a symmetric CSR matrix by Lanczos with full
// re-orthogonalisation (host, set-up only).  Replaces the Spectra
// SymEigsSolver(nev=1, ncv=6, tol=1e-3) call of
// Graph::computePreconditionerRegularization (src/Graph.cpp:1919-1960).
// Returns < 0 when it does not converge (caller falls back to 0.1 like the
// reference).
double lambda_max_lanczos(int k, const int *Ap, const int *Ai, const double *Ax,
                          double tol, int max_steps) {
  if (k <= 0) return -1.0;
  const int m = std::min(max_steps, k);
  std::vector<std::vector<double>> V;
  V.reserve(m + 1);
  std::vector<double> alpha, beta, w(k);
  uint64_t st = 0x2545F4914F6CDD1Dull;
  std::vector<double> v(k);
  double nrm = 0.0;
  for (int i = 0; i < k; ++i) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    v[i] = (double)(st >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    nrm += v[i] * v[i];
  }
  nrm = 1.0 / std::sqrt(nrm);
  for (double &x : v) x *= nrm;
  V.push_back(v);
  double theta = -1.0;
  for (int j = 0; j < m; ++j) {
    const std::vector<double> &vj = V[j];
    for (int i = 0; i < k; ++i) {
      double s = 0.0;
      for (int p = Ap[i]; p < Ap[i + 1]; ++p) s += Ax[p] * vj[Ai[p]];
      w[i] = s;
    }
    double a = 0.0;
    for (int i = 0; i < k; ++i) a += w[i] * vj[i];
    alpha.push_back(a);
    for (int pass = 0; pass < 2; ++pass)
      for (int c = 0; c <= j; ++c) {
        double h = 0.0;
        for (int i = 0; i < k; ++i) h += V[c][i] * w[i];
        for (int i = 0; i < k; ++i) w[i] -= h * V[c][i];
      }
    double b = 0.0;
    for (int i = 0; i < k; ++i) b += w[i] * w[i];
    b = std::sqrt(b);
    // largest Ritz value of the (j+1) x (j+1) tridiagonal by bisection-free
    // power of QL is overkill: use Sturm bisection on the Gershgorin interval
    const int q = j + 1;
    double lo = alpha[0], hi = alpha[0];
    for (int i = 0; i < q; ++i) {
      const double rad = (i > 0 ? std::fabs(beta[i - 1]) : 0.0) +
                         (i + 1 < q ? std::fabs(beta[i]) : 0.0);
      lo = std::min(lo, alpha[i] - rad);
      hi = std::max(hi, alpha[i] + rad);
    }
    auto count_below = [&](double x) {  // number of eigenvalues < x
      int cnt = 0;
      double dprev = 1.0;
      for (int i = 0; i < q; ++i) {
        const double b2 = i > 0 ? beta[i - 1] * beta[i - 1] : 0.0;
        double dcur = alpha[i] - x - (i > 0 ? b2 / dprev : 0.0);
        if (dcur == 0.0) dcur = -1e-300;
        if (dcur < 0.0) ++cnt;
        dprev = dcur;
      }
      return cnt;
    };
    double a0 = lo, b0 = hi;
    for (int it = 0; it < 200 && b0 - a0 > 1e-14 * std::max(1.0, std::fabs(b0)); ++it) {
      const double mid = 0.5 * (a0 + b0);
      if (count_below(mid) >= q) b0 = mid; else a0 = mid;
    }
    theta = 0.5 * (a0 + b0);
    // residual of the top Ritz pair = |b * s_q|, s = eigenvector of T: solve by
    // the three-term recurrence (T - theta I) s = 0 from the top
    std::vector<double> sv(q);
    sv[0] = 1.0;
    if (q > 1) sv[1] = -(alpha[0] - theta) / beta[0];
    for (int i = 1; i + 1 < q; ++i)
      sv[i + 1] = -((alpha[i] - theta) * sv[i] + beta[i - 1] * sv[i - 1]) / beta[i];
    double sn = 0.0;
    for (double x : sv) sn += x * x;
    const double res = std::fabs(b * sv[q - 1]) / std::sqrt(sn);
    if (res <= tol * std::fabs(theta) || b <= 1e-300) return theta;
    if (j + 1 == m) break;
    beta.push_back(b);
    std::vector<double> vn(k);
    for (int i = 0; i < k; ++i) vn[i] = w[i] / b;
    V.push_back(std::move(vn));
  }
  return k <= max_steps ? theta : -1.0;
}

}  // namespace dcora
