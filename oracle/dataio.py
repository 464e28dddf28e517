"""Oracle: g2o / pyfg readers and measurement containers (TEST INFRASTRUCTURE).

Restates src/DCORA_utils.cpp:179-375 (read_g2o_file), :437-1167
(read_pyfg_file) and :1169-1352 (getLocalToGlobalStateMapping /
getGlobalMeasurements) of the reference.  Measurements are held as
structure-of-arrays so that the same containers scale to millions of edges.
"""
from __future__ import annotations

import numpy as np

POSE, LANDMARK = 0, 1  # StateType of an endpoint
MAP_ID = ord("M") - ord("A")  # include/DCORA/DCORA_types.h:43-44


class PosePose:
    """RelativePosePoseMeasurement arrays (include/DCORA/Measurements.h)."""

    def __init__(self, d, r1, p1, r2, p2, R, t, kappa, tau, weight=None):
        m = len(p1)
        self.d = d
        self.r1 = np.asarray(r1, np.int64).reshape(m)
        self.p1 = np.asarray(p1, np.int64).reshape(m)
        self.r2 = np.asarray(r2, np.int64).reshape(m)
        self.p2 = np.asarray(p2, np.int64).reshape(m)
        self.R = np.asarray(R, np.float64).reshape(m, d, d)
        self.t = np.asarray(t, np.float64).reshape(m, d)
        self.kappa = np.asarray(kappa, np.float64).reshape(m)
        self.tau = np.asarray(tau, np.float64).reshape(m)
        self.weight = (np.ones(m) if weight is None
                       else np.asarray(weight, np.float64).reshape(m))

    def __len__(self):
        return len(self.p1)

    def take(self, idx):
        return PosePose(self.d, self.r1[idx], self.p1[idx], self.r2[idx],
                        self.p2[idx], self.R[idx], self.t[idx],
                        self.kappa[idx], self.tau[idx], self.weight[idx])

    @staticmethod
    def empty(d):
        z = np.zeros(0, np.int64)
        return PosePose(d, z, z, z, z, np.zeros((0, d, d)), np.zeros((0, d)),
                        np.zeros(0), np.zeros(0))


class PoseLandmark:
    """RelativePoseLandmarkMeasurement arrays."""

    def __init__(self, d, r1, p1, r2, p2, t, tau, weight=None):
        m = len(p1)
        self.d = d
        self.r1 = np.asarray(r1, np.int64).reshape(m)
        self.p1 = np.asarray(p1, np.int64).reshape(m)
        self.r2 = np.asarray(r2, np.int64).reshape(m)
        self.p2 = np.asarray(p2, np.int64).reshape(m)
        self.t = np.asarray(t, np.float64).reshape(m, d)
        self.tau = np.asarray(tau, np.float64).reshape(m)
        self.weight = (np.ones(m) if weight is None
                       else np.asarray(weight, np.float64).reshape(m))

    def __len__(self):
        return len(self.p1)

    def take(self, idx):
        return PoseLandmark(self.d, self.r1[idx], self.p1[idx], self.r2[idx],
                            self.p2[idx], self.t[idx], self.tau[idx],
                            self.weight[idx])

    @staticmethod
    def empty(d):
        z = np.zeros(0, np.int64)
        return PoseLandmark(d, z, z, z, z, np.zeros((0, d)), np.zeros(0))


class Range:
    """RangeMeasurement arrays; ``l`` is the unit-sphere index owned by r1."""

    def __init__(self, r1, p1, r2, p2, type1, type2, l, rng, precision,
                 weight=None):
        m = len(p1)
        self.r1 = np.asarray(r1, np.int64).reshape(m)
        self.p1 = np.asarray(p1, np.int64).reshape(m)
        self.r2 = np.asarray(r2, np.int64).reshape(m)
        self.p2 = np.asarray(p2, np.int64).reshape(m)
        self.type1 = np.asarray(type1, np.int64).reshape(m)
        self.type2 = np.asarray(type2, np.int64).reshape(m)
        self.l = np.asarray(l, np.int64).reshape(m)
        self.range = np.asarray(rng, np.float64).reshape(m)
        self.precision = np.asarray(precision, np.float64).reshape(m)
        self.weight = (np.ones(m) if weight is None
                       else np.asarray(weight, np.float64).reshape(m))

    def __len__(self):
        return len(self.p1)

    def take(self, idx):
        return Range(self.r1[idx], self.p1[idx], self.r2[idx], self.p2[idx],
                     self.type1[idx], self.type2[idx], self.l[idx],
                     self.range[idx], self.precision[idx], self.weight[idx])

    @staticmethod
    def empty():
        z = np.zeros(0, np.int64)
        return Range(z, z, z, z, z, z, z, np.zeros(0), np.zeros(0))


class Measurements:
    def __init__(self, d, pp=None, pl=None, rg=None):
        self.d = d
        self.pp = pp if pp is not None else PosePose.empty(d)
        self.pl = pl if pl is not None else PoseLandmark.empty(d)
        self.rg = rg if rg is not None else Range.empty()


def quat_to_rot(qx, qy, qz, qw):
    """Eigen::Quaterniond(w,x,y,z).toRotationMatrix() -- NO normalisation,
    exactly as Eigen does it (used at DCORA_utils.cpp:253,320,490-493)."""
    tx, ty, tz = 2 * qx, 2 * qy, 2 * qz
    twx, twy, twz = tx * qw, ty * qw, tz * qw
    txx, txy, txz = tx * qx, ty * qx, tz * qx
    tyy, tyz, tzz = ty * qy, tz * qy, tz * qz
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def rot2d(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s], [s, c]])


def read_g2o(filename):
    """read_g2o_file (DCORA_utils.cpp:179-375).

    Returns (Measurements with all robot ids 0, num_poses, ground_truth) where
    ground_truth maps pose index -> (R, t).
    """
    dim = 0
    p1, p2, Rs, ts, kap, tau = [], [], [], [], [], []
    gt = {}
    with open(filename) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "VERTEX_SE2":
                dim = dim or 2
                i = int(tok[1])
                x, y, th = map(float, tok[2:5])
                assert i not in gt, "duplicate pose ID"
                gt[i] = (rot2d(th), np.array([x, y]))
            elif tok[0] == "VERTEX_SE3:QUAT":
                dim = dim or 3
                i = int(tok[1])
                v = list(map(float, tok[2:9]))
                assert i not in gt, "duplicate pose ID"
                gt[i] = (quat_to_rot(v[3], v[4], v[5], v[6]), np.array(v[0:3]))
            elif tok[0] == "EDGE_SE2":
                # NB the reference only sets `dim` on VERTEX lines (:224-256)
                # and would LOG(FATAL) at :366 on a file without them (e.g.
                # data/kitti_00.g2o); inferring it from the edge type is a
                # deliberate, documented deviation.
                dim = dim or 2
                i, j = int(tok[1]), int(tok[2])
                dx, dy, dth, I11, I12, I13, I22, I23, I33 = map(float, tok[3:12])
                tran = np.array([[I11, I12], [I12, I22]])
                p1.append(i); p2.append(j)
                Rs.append(rot2d(dth)); ts.append([dx, dy])
                tau.append(2.0 / np.trace(np.linalg.inv(tran)))   # :288-292
                kap.append(I33)
            elif tok[0] == "EDGE_SE3:QUAT":
                dim = dim or 3
                i, j = int(tok[1]), int(tok[2])
                v = list(map(float, tok[3:31]))
                dx, dy, dz, qx, qy, qz, qw = v[0:7]
                (I11, I12, I13, I14, I15, I16, I22, I23, I24, I25, I26, I33,
                 I34, I35, I36, I44, I45, I46, I55, I56, I66) = v[7:28]
                tran = np.array([[I11, I12, I13], [I12, I22, I23], [I13, I23, I33]])
                rot = np.array([[I44, I45, I46], [I45, I55, I56], [I46, I56, I66]])
                p1.append(i); p2.append(j)
                Rs.append(quat_to_rot(qx, qy, qz, qw)); ts.append([dx, dy, dz])
                tau.append(3.0 / np.trace(np.linalg.inv(tran)))        # :337-340
                kap.append(3.0 / (2.0 * np.trace(np.linalg.inv(rot))))  # :342-346
            else:
                raise ValueError("unrecognized type: " + tok[0])
    assert dim in (2, 3)
    m = len(p1)
    num_poses = max(max(p1), max(p2)) + 1
    z = np.zeros(m, np.int64)
    pp = PosePose(dim, z, p1, z, p2, np.array(Rs), np.array(ts), kap, tau)
    return Measurements(dim, pp=pp), num_poses, gt


def _sym_robot_state(sym):
    """getRobotAndStateIDFromSymbol / getStateTypeFromSymbol (:584-628)."""
    if sym[0] == "L":
        if sym[1].isupper():
            return ord(sym[1]) - ord("A"), int(sym[2:]), LANDMARK
        return MAP_ID, int(sym[1:]), LANDMARK
    if sym[0].isupper():
        return ord(sym[0]) - ord("A"), int(sym[1:]), POSE
    raise ValueError("bad symbol " + sym)


def _read_sym(vals, dim):
    cov = np.zeros((dim, dim))
    it = iter(vals)
    for i in range(dim):
        for j in range(i, dim):
            cov[i, j] = cov[j, i] = next(it)
    return cov


def _tau(cov):      # getTau :551-568
    return cov.shape[0] / np.trace(cov)


def _kappa(cov):    # getKappa :570-588
    if cov.shape[0] == 1:
        return 1.0 / cov[0, 0]
    return 3.0 / (2.0 * np.trace(cov))


class PyFG:
    """Parsed pyfg dataset in LOCAL (robot, frame) ids + ground truth."""

    def __init__(self):
        self.dim = 0
        self.robot_ids = set()
        self.gt_poses = {}        # (robot, frame) -> (R, t)
        self.gt_landmarks = {}    # (robot, frame) -> t
        self.gt_spheres = {}      # (robot, idx) -> unit vector
        self.pose_priors = []     # parsed but dropped by getGlobalMeasurements
        self.landmark_priors = []
        self.meas = None          # Measurements in local ids


def read_pyfg(filename):
    """read_pyfg_file (DCORA_utils.cpp:437-1167)."""
    ds = PyFG()
    pp, pl, rg = [], [], []
    sphere_ctr = {}
    seen_ranges = set()
    with open(filename) as fh:
        lines = [ln.split() for ln in fh if ln.strip()]
    first = lines[0][0]
    # getDimFromPyfgFirstLine (:376-435): dimension from the first item type
    if first == "EDGE_RANGE":
        raise ValueError("first pyfg line carries no dimension information")
    ds.dim = 3 if ("SE3" in first or "XYZ" in first) else 2
    d = ds.dim
    for tok in lines:
        ty = tok[0]
        if ty in ("VERTEX_SE2", "VERTEX_SE3:QUAT"):
            rid, sid, _ = _sym_robot_state(tok[2])
            v = list(map(float, tok[3:]))
            if ty == "VERTEX_SE2":
                t, R = np.array(v[0:2]), rot2d(v[2])
            else:
                t, R = np.array(v[0:3]), quat_to_rot(v[3], v[4], v[5], v[6])
            ds.robot_ids.add(rid)
            assert (rid, sid) not in ds.gt_poses
            ds.gt_poses[(rid, sid)] = (R, t)
        elif ty in ("VERTEX_XY", "VERTEX_XYZ"):
            rid, sid, _ = _sym_robot_state(tok[1])
            ds.robot_ids.add(rid)
            assert (rid, sid) not in ds.gt_landmarks
            ds.gt_landmarks[(rid, sid)] = np.array(list(map(float, tok[2:2 + d])))
        elif ty in ("VERTEX_SE2:PRIOR", "VERTEX_SE3:QUAT:PRIOR"):
            rid, sid, _ = _sym_robot_state(tok[2])
            v = list(map(float, tok[3:]))
            if d == 2:
                t, R, cov = np.array(v[0:2]), rot2d(v[2]), _read_sym(v[3:], 3)
                cov_t, cov_R = cov[:2, :2], cov[2:, 2:]
            else:
                t, R = np.array(v[0:3]), quat_to_rot(v[3], v[4], v[5], v[6])
                cov = _read_sym(v[7:], 6)
                cov_t, cov_R = cov[:3, :3], cov[3:, 3:]
            ds.pose_priors.append((rid, sid, R, t, _kappa(cov_R), _tau(cov_t)))
        elif ty in ("VERTEX_XY:PRIOR", "VERTEX_XYZ:PRIOR"):
            rid, sid, _ = _sym_robot_state(tok[2])
            v = list(map(float, tok[3:]))
            ds.landmark_priors.append((rid, sid, np.array(v[0:d]),
                                       _tau(_read_sym(v[d:], d))))
        elif ty in ("EDGE_SE2", "EDGE_SE3:QUAT"):
            r1, s1, _ = _sym_robot_state(tok[2])
            r2, s2, _ = _sym_robot_state(tok[3])
            v = list(map(float, tok[4:]))
            if d == 2:
                t, R, cov = np.array(v[0:2]), rot2d(v[2]), _read_sym(v[3:], 3)
                cov_t, cov_R = cov[:2, :2], cov[2:, 2:]
            else:
                t, R = np.array(v[0:3]), quat_to_rot(v[3], v[4], v[5], v[6])
                cov = _read_sym(v[7:], 6)
                cov_t, cov_R = cov[:3, :3], cov[3:, 3:]
            pp.append((r1, s1, r2, s2, R, t, _kappa(cov_R), _tau(cov_t)))
        elif ty in ("EDGE_SE2_XY", "EDGE_SE3_XYZ"):
            r1, s1, _ = _sym_robot_state(tok[2])
            r2, s2, _ = _sym_robot_state(tok[3])
            v = list(map(float, tok[4:]))
            pl.append((r1, s1, r2, s2, np.array(v[0:d]),
                       _tau(_read_sym(v[d:], d))))
        elif ty == "EDGE_RANGE":
            r1, s1, t1 = _sym_robot_state(tok[2])
            r2, s2, t2 = _sym_robot_state(tok[3])
            rng, cov = float(tok[4]), float(tok[5])
            assert rng > 0.0
            key = (t1, r1, s1, t2, r2, s2)
            if key in seen_ranges:       # duplicate ranges skipped :1078-1086
                continue
            seen_ranges.add(key)
            l = sphere_ctr.get(r1, 0)    # source robot owns the sphere :1092-1094
            sphere_ctr[r1] = l + 1
            a = (ds.gt_poses[(r1, s1)][1] if t1 == POSE
                 else ds.gt_landmarks[(r1, s1)])
            b = (ds.gt_poses[(r2, s2)][1] if t2 == POSE
                 else ds.gt_landmarks[(r2, s2)])
            u = a - b
            ds.gt_spheres[(r1, l)] = u / np.linalg.norm(u)   # :1136-1137
            rg.append((r1, s1, r2, s2, t1, t2, l, rng, 1.0 / cov))
        else:
            raise ValueError("unknown item type: " + ty)

    def col(rows, k, dt=None):
        return np.array([r[k] for r in rows], dtype=dt) if rows else np.zeros(0, dt)

    PP = (PosePose(d, col(pp, 0, np.int64), col(pp, 1, np.int64),
                   col(pp, 2, np.int64), col(pp, 3, np.int64),
                   np.array([r[4] for r in pp]).reshape(-1, d, d),
                   np.array([r[5] for r in pp]).reshape(-1, d),
                   col(pp, 6), col(pp, 7)) if pp else PosePose.empty(d))
    PL = (PoseLandmark(d, col(pl, 0, np.int64), col(pl, 1, np.int64),
                       col(pl, 2, np.int64), col(pl, 3, np.int64),
                       np.array([r[4] for r in pl]).reshape(-1, d), col(pl, 5))
          if pl else PoseLandmark.empty(d))
    RG = (Range(*[col(rg, k) for k in range(9)]) if rg else Range.empty())
    ds.meas = Measurements(d, PP, PL, RG)
    return ds


def global_measurements(ds: PyFG):
    """getGlobalMeasurements (DCORA_utils.cpp:1230-1352): re-index everything
    onto robot 0 with consecutive ids in std::map order (robot, frame); priors
    are dropped (:1244).  Returns (Measurements, n, l, b, X_gt) where X_gt is
    the d x ((d+1)n+l+b) ground truth in RA ordering."""
    d = ds.dim
    pose_map = {k: i for i, k in enumerate(sorted(ds.gt_poses))}
    lmk_map = {k: i for i, k in enumerate(sorted(ds.gt_landmarks))}
    sph_map = {k: i for i, k in enumerate(sorted(ds.gt_spheres))}
    n, l, b = len(pose_map), len(sph_map), len(lmk_map)
    m = ds.meas
    pp = m.pp
    z = np.zeros(len(pp), np.int64)
    g1 = [pose_map[(a, c)] for a, c in zip(pp.r1, pp.p1)]
    g2 = [pose_map[(a, c)] for a, c in zip(pp.r2, pp.p2)]
    PP = PosePose(d, z, g1, z, g2, pp.R, pp.t, pp.kappa, pp.tau, pp.weight)
    pl = m.pl
    z = np.zeros(len(pl), np.int64)
    g1 = [pose_map[(a, c)] for a, c in zip(pl.r1, pl.p1)]
    g2 = [lmk_map[(a, c)] for a, c in zip(pl.r2, pl.p2)]
    PL = PoseLandmark(d, z, g1, z, g2, pl.t, pl.tau, pl.weight)
    rg = m.rg
    z = np.zeros(len(rg), np.int64)
    g1 = [(pose_map if t == POSE else lmk_map)[(a, c)]
          for a, c, t in zip(rg.r1, rg.p1, rg.type1)]
    g2 = [(pose_map if t == POSE else lmk_map)[(a, c)]
          for a, c, t in zip(rg.r2, rg.p2, rg.type2)]
    gl = [sph_map[(a, c)] for a, c in zip(rg.r1, rg.l)]
    RG = Range(z, g1, z, g2, rg.type1, rg.type2, gl, rg.range, rg.precision,
               rg.weight)
    X = np.zeros((d, (d + 1) * n + l + b))
    for key, i in pose_map.items():
        R, t = ds.gt_poses[key]
        X[:, i * d:(i + 1) * d] = R
        X[:, d * n + l + i] = t
    for key, i in sph_map.items():
        X[:, d * n + i] = ds.gt_spheres[key]
    for key, i in lmk_map.items():
        X[:, d * n + l + n + i] = ds.gt_landmarks[key]
    return Measurements(d, PP, PL, RG), n, l, b, X


def robot_measurements(ds: PyFG):
    """getRobotMeasurements (DCORA_utils.cpp:1354-1512): per-robot measurement
    sets (every measurement touching the robot) with pose / landmark frame ids
    re-based to start at 0.  Returns {robot: Measurements}, first_pose_id,
    first_landmark_id."""
    m = ds.meas
    first_pose, first_lmk = {}, {}
    for rid in ds.robot_ids:
        ps = [s for (r, s) in ds.gt_poses if r == rid]
        ls = [s for (r, s) in ds.gt_landmarks if r == rid]
        first_pose[rid] = min(ps) if ps else 0
        first_lmk[rid] = min(ls) if ls else 0
    fp = np.vectorize(lambda r: first_pose.get(r, 0), otypes=[np.int64])
    fl = np.vectorize(lambda r: first_lmk.get(r, 0), otypes=[np.int64])
    out = {}
    for rid in sorted(ds.robot_ids):
        pp, pl, rg = m.pp, m.pl, m.rg
        s = np.nonzero((pp.r1 == rid) | (pp.r2 == rid))[0]
        q = pp.take(s)
        if len(q):
            q.p1 = q.p1 - fp(q.r1); q.p2 = q.p2 - fp(q.r2)
        s = np.nonzero((pl.r1 == rid) | (pl.r2 == rid))[0]
        u = pl.take(s)
        if len(u):
            u.p1 = u.p1 - fp(u.r1); u.p2 = u.p2 - fl(u.r2)
        s = np.nonzero((rg.r1 == rid) | (rg.r2 == rid))[0]
        g = rg.take(s)
        if len(g):
            g.p1 = g.p1 - np.where(g.type1 == POSE, fp(g.r1), fl(g.r1))
            g.p2 = g.p2 - np.where(g.type2 == POSE, fp(g.r2), fl(g.r2))
        out[rid] = Measurements(ds.dim, q, u, g)
    return out, first_pose, first_lmk
