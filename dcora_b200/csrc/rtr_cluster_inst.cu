// One relaxation rank of the per-agent RTR cluster kernel (rtr_cluster.cuh); built once per
// rank with -DDCORA_CLUSTER_R=<r> so that the ranks compile in parallel.
#include "rtr_cluster.cuh"

#ifndef DCORA_CLUSTER_R
#error "compile with -DDCORA_CLUSTER_R=<rank>"
#endif
#define DCORA_CAT2(a, b) a##b
#define DCORA_CAT(a, b) DCORA_CAT2(a, b)

namespace dcora {

int DCORA_CAT(launch_cluster_r, DCORA_CLUSTER_R)(dcora_problem *p, const ClusterArgs &args,
                                                 int n_agents, int nw) {
  constexpr int R = DCORA_CLUSTER_R;
  if (p->M->d == 3) {
    if constexpr (R >= 3) return launch_cluster_shape<R, 4>(p, args, n_agents, nw);
  } else {
    if constexpr (R >= 2) return launch_cluster_shape<R, 3>(p, args, n_agents, nw);
  }
  set_error("rank below the dimension");
  return DCORA_ERR_ARG;
}

}  // namespace dcora
