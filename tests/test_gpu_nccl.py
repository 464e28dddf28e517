"""Two ranks over NCCL against one rank: the agents of a range-aided
grid are dealt to 2 processes (one GPU each), public states travel by all-gather, and the
iterates after several RBCD / RBCD++ iterations must be BIT-IDENTICAL to the 1-rank run's:
the coupling product sums every column of G in ascending global-column order
(dcora_problem_set_coupling_keyed), everything else is per agent.  Skipped with < 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(S, robots, r):
    from dcora_b200 import graph as pg, synthetic
    meas, gt, idx = synthetic.grid3d(S, 0.1, seed=6)
    gp = pg.GlobalProblem(meas, pg.Partition.contiguous(meas.n, robots))
    X0 = synthetic.lifted_initial_guess(meas, idx, gp, S, r)
    return gp, X0


def _run(gp, X0, r, rank, world, group, iters, acc):
    import dcora_b200 as dc
    from dcora_b200 import rbcd
    G = rbcd.RBCD(gp, r, rank=rank, world=world, device=rank, group=group,
                  params=dc.ROptParameters(time_bound_s=0.0), acceleration=acc)
    G.set_X(X0)
    hist = []
    for _ in range(iters):
        G.iterate()
        hist.append(G.monitor())
    return G.plan.c0, G.get_X(), hist


def _worker(rank, world, port, S, robots, r, iters, acc, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world, device_id=torch.device("cuda", rank))
        gp, X0 = _problem(S, robots, r)
        c0, X, hist = _run(gp, X0, r, rank, world, None, iters, acc)
        q.put((rank, c0, X, hist))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, -1, traceback.format_exc(), repr(e)))


@pytest.mark.parametrize("acc", [False, True], ids=["rbcd", "rbcd++"])
def test_two_ranks_are_bit_identical_to_one_rank(acc):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    S, robots, r, iters = 8, 16, 5, 6
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, S, robots, r, iters, acc, q))
             for rk in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    gp, X0 = _problem(S, robots, r)
    _, X1, hist1 = _run(gp, X0, r, 0, 1, None, iters, acc)
    got = np.zeros_like(X1)
    for rank, c0, X, hist in sorted(res, key=lambda t: t[0]):
        assert c0 >= 0, f"rank {rank}: {X}"
        got[:, c0:c0 + X.shape[1]] = X
        for (f2, g2), (f1, g1) in zip(hist, hist1):
            # f = sum_a quad_a + 0.5 lin_a cancels over 4 orders of magnitude and the 2-rank
            # sum goes through an all-reduce: the monitor values agree to rounding of the terms
            assert abs(f2 - f1) <= 1e-9 * abs(f1) and abs(g2 - g1) <= 1e-9 * g1
    assert np.array_equal(got, X1)
    assert hist1[-1][0] < hist1[0][0]


def _stair_worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world, device_id=torch.device("cuda", rank))
        q.put((rank,) + _staircase(rank, world))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, None, traceback.format_exc(), repr(e)))


def _staircase(rank, world):
    import dcora_b200 as dc
    from dcora_b200 import rbcd
    from dcora_b200.staircase import rbcd_staircase
    from oracle import manifold as mf
    from tests.common import central_problem
    from tests.test_gpu_rbcd import _setup
    gp, blocks, cross = _setup("smallGrid3D", 5, 3)
    G = rbcd.RBCD(gp, 3, rank=rank, world=world, device=rank,
                  params=dc.ROptParameters(time_bound_s=0.0, gradnorm_tol=1e-9,
                                           RTR_iterations=10, RTR_tCG_iterations=100))
    X0 = mf.random_point(central_problem("smallGrid3D", 3)[0].L, np.random.default_rng(7))
    res = rbcd_staircase(G, X0, r_max=7, eta=1e-3, gradnorm_tol=1e-4, max_iters=800,
                         check_every=10, certify_tol=1e-9)
    return (res.certified, res.rank, res.cost, res.lambda_min,
            [(lv["r"], lv["is_psd"]) for lv in res.levels])


def test_two_rank_staircase_escapes_and_certifies_like_one_rank():
    # RBCD -> distributed certificate -> distributed escapeSaddle -> next rank with the agents
    # dealt to 2 processes: same levels, same certified optimum as on one process
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_stair_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    one = _staircase(0, 1)
    assert one[0] and abs(2 * one[2] - 1025.3980521) < 1e-4
    for r in res:
        assert r[1] is not None, f"rank {r[0]}: {r[2]}"
        certified, rank_, cost, lam, levels = r[1:]
        assert certified and rank_ == one[1] and levels == one[4]
        assert abs(cost - one[2]) <= 1e-8 * abs(one[2])
        assert lam > -1e-3
