#!/usr/bin/env python
"""lambda_min of the distributed certificate on a small synthetic range-aided grid after a few
RBCD iterations; run once plainly (1 GPU) and once under torchrun (N GPUs): the numbers must agree.
    python tools/dist_certify_check.py
    python -m torch.distributed.run --nproc-per-node 2 ... tools/dist_certify_check.py"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
    import torch
    import torch.distributed as dist
    import dcora_b200 as dc
    from dcora_b200 import rbcd
    from dcora_b200.certify_dist import certify_distributed
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    a = argparse.Namespace(grid=10, agents=16, rank=5, range_fraction=0.1, seed=0)
    log = lambda m: None
    meas, gt, idx, gp = bench.build_problem(a, log)
    S = rbcd.RBCD(gp, a.rank, rank=rank, world=world, device=local,
                  params=dc.ROptParameters(time_bound_s=0.0))
    S.set_X(bench.initial_X(a, meas, idx, gp))
    for _ in range(5):
        S.iterate()
    f, g = S.monitor()
    res = certify_distributed(S, 1e-3, tol=1e-10, max_iters=int(os.environ.get('CERT_ITERS', '60')),
                              precondition=os.environ.get('CERT_PRECOND', '1') == '1')
    f2, g2 = S.monitor()
    if rank == 0:
        sys.stderr.write(json.dumps({"world": world, "f": f, "gradnorm": g, "lambda_min": res.lambda_min,
                                     "is_psd": res.is_psd, "iterations": res.iterations,
                                     "products": res.products, "converged": res.converged,
                                     "state_restored": (f, g) == (f2, g2)}) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
