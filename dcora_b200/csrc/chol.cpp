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
#include <cmath>
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

}  // namespace

int sparse_cholesky(int k, const std::vector<int> &Ap, const std::vector<int> &Ai,
                    const std::vector<double> &Ax, int dh, int npose,
                    HostFactor *out) {
  HostFactor &F = *out;
  F.k = k;
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
  const std::vector<int> order = minimum_degree(nsv, adj, weight);
  F.perm.clear();
  F.perm.reserve(k);
  for (int s : order) {
    const int c0 = s < npose ? s * dh : npc + (s - npose);
    const int c1 = s < npose ? c0 + dh : c0 + 1;
    for (int c = c0; c < c1; ++c) F.perm.push_back(c);
  }
  std::vector<int> pinv(k);
  for (int i = 0; i < k; ++i) pinv[F.perm[i]] = i;

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
  // ---- numeric up-looking factorisation (row k of L from rows < k)
  std::vector<int> cfill(F.Lp.begin(), F.Lp.end() - 1), stack(k), pattern(k);
  std::vector<double> x(k, 0.0);
  std::fill(flag.begin(), flag.end(), -1);
  for (int j = 0; j < k; ++j) {
    // pattern of row j = reach of column j's upper entries in the etree,
    // produced in topological order (cs_ereach)
    int top = k;
    flag[j] = j;
    double d = 0.0;
    for (int p = Cp[j]; p < Cp[j + 1]; ++p) {
      int i = Ci[p];
      if (i == j) { d += Cx[p]; continue; }
      x[i] += Cx[p];   // += : tolerate duplicate entries
      int len = 0;
      while (flag[i] != j) {
        stack[len++] = i;
        flag[i] = j;
        i = parent[i];
      }
      while (len > 0) pattern[--top] = stack[--len];
    }
    for (; top < k; ++top) {
      const int i = pattern[top];
      const double lji = x[i] / F.Lx[F.Lp[i]];  // L(j,i) = x_i / L(i,i)
      x[i] = 0.0;
      for (int p = F.Lp[i] + 1; p < cfill[i]; ++p) x[F.Li[p]] -= F.Lx[p] * lji;
      d -= lji * lji;
      const int q = cfill[i]++;
      F.Li[q] = j;
      F.Lx[q] = lji;
    }
    if (!(d > 0.0) || !std::isfinite(d)) {
      set_error("sparse_cholesky: matrix is not positive definite");
      return DCORA_ERR_FACTOR;
    }
    const int q = cfill[j]++;
    F.Li[q] = j;
    F.Lx[q] = std::sqrt(d);
  }
  return DCORA_OK;
}

}  // namespace dcora
