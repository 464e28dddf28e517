"""Generate the committed golden fixtures from the reference's data files.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py [--solve]
It parses the reference datasets with the oracle reader (oracle/dataio.py,
which restates src/DCORA_utils.cpp:179-375,437-1352) and stores the parsed
measurement arrays + ground truth as compressed .npz, so that the GPU box --
which has no /root/reference -- can run the parity tests.  With --solve it also
runs the oracle solver to a tight tolerance and records the optimal costs /
lambda_min in expected.json (slow: minutes).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import dataio  # noqa: E402

DATA = "/root/reference/data/"
G2O = ["pose_graph_optimization_test_2d", "pose_graph_optimization_test_3d",
       "tinyGrid3D", "smallGrid3D", "sphere2500", "torus3D", "city10000", "kitti_00"]
PYFG = ["range_aided_slam_test_2d", "range_aided_slam_test_3d", "tiers"]


def pack_meas(m):
    out = {"d": np.int64(m.d)}
    for name, obj in (("pp", m.pp), ("pl", m.pl), ("rg", m.rg)):
        for key, val in vars(obj).items():
            if isinstance(val, np.ndarray):
                out[f"{name}_{key}"] = val
    return out


def raw_quaternions():
    """quats.npz: the EDGE_SE3:QUAT quaternions (qx qy qz qw) exactly as written in the
    data files.  The reference reader does not normalise them
    (Eigen::Quaterniond(..).toRotationMatrix(), DCORA_utils.cpp:331), so the measured
    rotations are slightly non-orthogonal; tests/common.write_g2o needs the raw values
    to regenerate a g2o file that parses to bit-identical measurements."""
    out = {}
    for name in G2O:
        q = []
        with open(DATA + name + ".g2o") as fh:
            for line in fh:
                tok = line.split()
                if tok and tok[0] == "EDGE_SE3:QUAT":
                    q.append([float(v) for v in tok[6:10]])
        if q:
            out[name] = np.array(q)
    np.savez_compressed(os.path.join(HERE, "quats.npz"), **out)
    print("quats:", {k: v.shape for k, v in out.items()})


def main():
    if "--quats" in sys.argv:
        return raw_quaternions()
    for name in G2O:
        meas, n, gt = dataio.read_g2o(DATA + name + ".g2o")
        d = meas.d
        X = np.zeros((d, (d + 1) * n))
        has_gt = len(gt) == n
        for i, (R, t) in gt.items():
            X[:, i * (d + 1):i * (d + 1) + d] = R
            X[:, i * (d + 1) + d] = t
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind="g2o", n=n, l=0, b=0,
                            has_gt=has_gt, X_gt=X, **pack_meas(meas))
        print(name, "d", d, "n", n, "m", len(meas.pp))
    for name in PYFG:
        ds = dataio.read_pyfg(DATA + name + ".pyfg")
        meas, n, l, b, X = dataio.global_measurements(ds)
        extra = {}
        if name != "tiers":
            # per-robot views for the multi-agent RA tests (local ids)
            extra = {"local_" + k: v for k, v in pack_meas(ds.meas).items()}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind="pyfg", n=n, l=l, b=b,
                            has_gt=True, X_gt=X, **pack_meas(meas), **extra)
        print(name, "d", ds.dim, "n", n, "l", l, "b", b)
    raw_quaternions()


if __name__ == "__main__":
    main()
