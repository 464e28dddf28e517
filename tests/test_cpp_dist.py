"""C++ host of the multi-process path (include/dcora_b200/DCORA_dist.hpp, examples/rbcd_ra.cpp)
against its Python twin: PyFG reader, partitioned data matrices, exchange plans (CPU), and the
RBCD / RBCD++ iterates of 1 and 2 processes driven from C++ with NCCL (GPU)."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests.common import product_measurements, write_pyfg


@pytest.fixture(scope="module")
def bins():
    from dcora_b200 import build
    build.build()
    exes = {os.path.basename(e): e for e in build.build_examples()}
    return exes


def _dataset(tmp_path, S=5, robots=6, seed=3, range_fraction=0.15):
    """A noisy range-aided grid written as PyFG and read back by the ORACLE's reader; the
    product's GlobalProblem is built from what the oracle read."""
    from dcora_b200 import graph as pg, synthetic
    from oracle import dataio
    meas, gt, idx = synthetic.grid3d(S, range_fraction, seed=seed)
    part = pg.Partition.contiguous(meas.n, robots)
    path = str(tmp_path / f"grid{S}_{robots}.pyfg")
    write_pyfg(path, meas, part, gt=gt)
    ds = dataio.read_pyfg(path)
    local, _, _ = dataio.robot_measurements(ds)
    # every measurement once, in file order, local (robot, frame) ids
    M, part2 = product_measurements(ds.meas, robots)
    gp = pg.GlobalProblem(M, part2)
    return path, gp, (meas, gt, idx)


@pytest.mark.parametrize("range_fraction", [0.15, 0.0], ids=["range-aided", "pose-graph"])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_cpp_pyfg_reader_partition_and_plans_match_python(bins, tmp_path, world, range_fraction):
    from dcora_b200 import rbcd
    path, gp, _ = _dataset(tmp_path, range_fraction=range_fraction)
    assert gp.ra == (range_fraction > 0)
    out = subprocess.run([bins["check_pyfg_partition"], path, str(world)], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout.splitlines()[-1])
    Q = gp.Q().tocoo()
    fro = float(np.sqrt(np.sum(Q.data ** 2)))
    chk = float(np.sum(Q.data * np.cos(0.001 * Q.row + 0.002 * Q.col)))
    assert (got["d"], got["n"], got["b"], got["l"], got["K"]) == \
        (gp.d, gp.meas.n, gp.meas.b, len(gp.meas.rg_i), gp.K)
    assert abs(got["fro"] - fro) <= 1e-12 * fro
    assert abs(got["checksum"] - chk) <= 1e-10 * fro
    plans = rbcd.make_plans(gp, world)
    assert len(got["plans"]) == world
    for p, g in zip(plans, got["plans"]):
        assert (g["a0"], g["a1"], g["c0"], g["K"]) == (p.a0, p.a1, p.c0, p.K)
        assert g["colors"] == [int(c) for c in p.colors] and g["num_colors"] == p.num_colors
        C = p.coupling.tocoo()
        assert g["coupling_rows"] == C.shape[0]
        cf = float(np.sqrt(np.sum(C.data ** 2)))
        cc = float(np.sum(C.data * np.cos(0.001 * C.row + 0.002 * C.col)))
        assert abs(g["coupling_fro"] - cf) <= 1e-12 * max(cf, 1.0)
        assert abs(g["coupling_checksum"] - cc) <= 1e-10 * max(cf, 1.0)
        if world > 1:
            assert tuple(g["seg"]) == tuple(p.seg)
            assert g["pack_sum"] == int(np.sum(p.pack_idx))
            assert g["key_sum"] == int(np.sum(p.row_key))


@pytest.mark.gpu
@pytest.mark.parametrize("ranks,acc", [(1, False), (1, True), (2, False), (2, True)])
def test_cpp_rbcd_host_matches_python_driver(bins, tmp_path, ranks, acc):
    import torch
    if torch.cuda.device_count() < ranks:
        pytest.skip(f"needs {ranks} GPUs")
    import dcora_b200 as dc
    from dcora_b200 import rbcd, synthetic
    path, gp, (meas, gt, idx) = _dataset(tmp_path, S=6, robots=8, seed=5)
    r, iters = 5, 6
    X0 = synthetic.lifted_initial_guess(meas, idx, gp, 6, r)
    x0 = tmp_path / "x0.bin"
    np.asfortranarray(X0).T.copy().tofile(x0)          # column-major r x K
    xo = tmp_path / "x.bin"
    cmd = [bins["rbcd_ra"], path, "--ranks", ranks, "--rank", r, "--iters", iters, "--tol", 0,
           "--x0", x0, "--out", xo] + (["--acc"] if acc else [])
    out = subprocess.run(list(map(str, cmd)), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    hist = [l for l in lines if "iter" in l]
    assert len(hist) == iters and lines[-1]["ranks"] == ranks
    Xc = np.fromfile(xo).reshape(gp.K, r).T
    G = rbcd.RBCD(gp, r, params=dc.ROptParameters(time_bound_s=0.0), acceleration=acc)
    G.set_X(X0)
    for it in range(iters):
        G.iterate()
        f, gn = G.monitor()
        assert abs(hist[it]["cost"] - f) <= 1e-8 * abs(f)
        assert abs(hist[it]["gradnorm"] - gn) <= 1e-6 * gn
    Xp = G.get_X()
    assert np.linalg.norm(Xc - Xp) <= 1e-7 * np.linalg.norm(Xp)


@pytest.mark.parametrize("name", ["range_aided_slam_test_2d", "range_aided_slam_test_3d"])
@pytest.mark.parametrize("world", [1, 2])
def test_cpp_reader_on_the_reference_fixtures_with_landmarks(bins, tmp_path, name, world):
    # the reference's 2-robot range-aided fixtures (poses, landmarks LA0 / LB0, ranges to poses
    # and to landmarks; 2-D: VERTEX_SE2 / EDGE_SE2 lines) re-written as PyFG from the golden
    # arrays and read by the C++ reader: same partitioned data matrices and plans as Python
    from dcora_b200 import graph as pg, rbcd
    from oracle import dataio
    from tests.common import load_golden
    meas, info = load_golden(name, prefix="local_")
    M, part = product_measurements(meas, 2)
    path = str(tmp_path / (name + ".pyfg"))
    write_pyfg(path, M, part)
    ds = dataio.read_pyfg(path)
    M2, part2 = product_measurements(ds.meas, 2)
    gp = pg.GlobalProblem(M2, part2, ra=True)
    out = subprocess.run([bins["check_pyfg_partition"], path, str(world)], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout.splitlines()[-1])
    Q = gp.Q().tocoo()
    fro = float(np.sqrt(np.sum(Q.data ** 2)))
    chk = float(np.sum(Q.data * np.cos(0.001 * Q.row + 0.002 * Q.col)))
    assert (got["d"], got["n"], got["b"], got["l"], got["K"], got["robots"]) == \
        (gp.d, gp.meas.n, gp.meas.b, len(gp.meas.rg_i), gp.K, 2)
    assert got["b"] == 2
    assert abs(got["fro"] - fro) <= 1e-12 * fro and abs(got["checksum"] - chk) <= 1e-10 * fro
    for p, g in zip(rbcd.make_plans(gp, world), got["plans"]):
        C = p.coupling.tocoo()
        cf = float(np.sqrt(np.sum(C.data ** 2)))
        assert (g["a0"], g["a1"], g["K"]) == (p.a0, p.a1, p.K)
        assert abs(g["coupling_fro"] - cf) <= 1e-12 * max(cf, 1.0)
        if world > 1:
            assert tuple(g["seg"]) == tuple(p.seg) and g["key_sum"] == int(np.sum(p.row_key))
