"""world_size-2 test of the N>1 host path on CPU (gloo): every rank builds its own exchange
plan (dcora_b200.rbcd.make_plans), packs its public columns, all-gathers them and rebuilds the
linear term G of its agents from [X_local | H] -- the data flow of RBCD.exchange/update_G with
the device gather and coupling product replaced by their numpy definitions.  Checked against
the centralised cross-coupling product (which tests/test_host_graph.py pins to the oracle)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, robots, r, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        from dcora_b200 import graph as pg, rbcd
        from tests.common import product_measurements
        from tests.test_host_graph import _multi
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world)
        meas, ra = _multi(name, robots)
        M, part = product_measurements(meas, robots)
        gp = pg.GlobalProblem(M, part, ra=ra)
        blocks = rbcd.rank_agents(gp.A, world)
        _, cross = gp.split(agents=range(*blocks[rank]))
        plan = rbcd.make_plans(gp, world, cross)[rank]
        X = np.random.default_rng(0).standard_normal((r, gp.K))      # same on every rank
        Xl = X[:, plan.c0:plan.c0 + plan.K]
        send = torch.from_numpy(np.ascontiguousarray(Xl[:, plan.pack_idx].T))   # (seg_cols, r)
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send)
        H = torch.cat(recv, 0).numpy().T
        XE = np.concatenate([Xl, H], axis=1)
        G = (plan.coupling.T @ XE.T).T
        # termination / monitoring all-reduce (RBCD.monitor)
        v = torch.tensor([float(np.sum(G * Xl)), float(plan.K)], dtype=torch.float64)
        dist.all_reduce(v)
        _, cross_all = gp.split(agents=[])
        Gref = (cross_all.T @ X.T).T
        err = float(np.abs(G - Gref[:, plan.c0:plan.c0 + plan.K]).max())
        ok_sum = abs(float(v[0]) - float(np.sum(Gref * X))) <= 1e-9 * max(1.0, abs(float(np.sum(Gref * X))))
        q.put((rank, err, float(np.abs(Gref).max()), int(v[1]) == gp.K and ok_sum))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e), 0.0, False))


@pytest.mark.parametrize("name,robots", [("smallGrid3D", 5), ("range_aided_slam_test_3d", 2)])
def test_two_rank_exchange_rebuilds_G(name, robots):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, name, robots, 4, q)) for rk in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, scale, ok in res:
        assert not isinstance(err, str), f"rank {rank}: {err}"
        assert err <= 1e-12 * max(1.0, scale) and ok
