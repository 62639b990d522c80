"""
Problem builders: the reference's known-answer problems and the synthetic BAL / pose-graph
workloads of BASELINE.json, lowered to `desc.Problem`.

Everything here is host-side numpy; nothing computes a linearization.
"""
import json
import os

import numpy as np

from . import desc as D

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class ValuesBuilder:
    """Flat insertion-ordered buffer, like sym::Values::data_ (symforce/opt/values.h:313)."""

    def __init__(self):
        self.chunks = []
        self.n = 0

    def add(self, arr):
        a = np.asarray(arr, dtype=np.float64).reshape(-1)
        off = self.n
        self.chunks.append(a)
        self.n += a.shape[0]
        return off

    def add_many(self, arr2d):
        """Adds rows of a [n, d] array contiguously; returns the offsets of each row."""
        a = np.ascontiguousarray(arr2d, dtype=np.float64)
        n, d = a.shape
        off = self.n
        self.chunks.append(a.reshape(-1))
        self.n += n * d
        return off + d * np.arange(n, dtype=np.int64)

    def data(self):
        return np.concatenate(self.chunks) if self.chunks else np.zeros(0)


# ------------------------------------------------------------------------------------------------
# quaternion helpers (storage order qx, qy, qz, qw like sym::Rot3)
# ------------------------------------------------------------------------------------------------
def quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack(
        [
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
            aw * bw - ax * bx - ay * by - az * bz,
        ],
        axis=-1,
    )


def quat_exp(v, eps=D.K_DEFAULT_EPSILON):
    th = np.sqrt(eps * eps + np.sum(v * v, axis=-1, keepdims=True))
    s = np.sin(0.5 * th) / th
    return np.concatenate([s * v, np.cos(0.5 * th)], axis=-1)


def quat_rotate(q, x):
    qv = q[..., :3]
    w = q[..., 3:4]
    t = 2.0 * np.cross(qv, x)
    return x + w * t + np.cross(qv, t)


def quat_from_matrix(R):
    """Batch rotation matrices [n,3,3] -> quaternions [n,4] (xyzw)."""
    n = R.shape[0]
    q = np.empty((n, 4))
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    for i in range(n):
        m = R[i]
        if tr[i] > 0:
            s = np.sqrt(tr[i] + 1.0) * 2
            q[i] = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
            q[i] = [0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s]
        elif m[1, 1] > m[2, 2]:
            s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
            q[i] = [(m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s]
        else:
            s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
            q[i] = [(m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s, (m[1, 0] - m[0, 1]) / s]
    return q / np.linalg.norm(q, axis=1, keepdims=True)


# ------------------------------------------------------------------------------------------------
# Known-answer problems of the reference's tests
# ------------------------------------------------------------------------------------------------
def _kat_init():
    with open(os.path.join(_GOLDEN, "kat_initial_values.json")) as f:
        return json.load(f)


def pose_smoothing(params=None):
    """test/symforce_optimizer_test.cc:79-134 (CreatePoseSmoothingProblem)."""
    init = np.array(_kat_init()["pose_smoothing"]).reshape(10, 7)
    vb = ValuesBuilder()
    pose_off = vb.add_many(init)
    eps_off = vb.add([1e-10])
    prior_start = vb.add([0, 0, 0, 1, 0, 0, 0])
    # FromYawPitchRoll(pi, 0, 0) (gen/cpp/sym/rot3.cc:170-200), normalised by the Rot3 ctor
    q = np.array([0.0, 0.0, np.sin(np.pi / 2), np.cos(np.pi / 2)])
    q = q / np.sqrt(np.sum(q * q))
    prior_last = vb.add([q[0], q[1], q[2], q[3], 5, 0, 0])
    si_prior = vb.add(np.diag(np.full(6, 1 / 0.1)).reshape(-1, order="F"))
    si_between = vb.add(np.diag(np.full(6, 1 / 0.5)).reshape(-1, order="F"))
    identity = vb.add([0, 0, 0, 1, 0, 0, 0])
    keys = [(D.TYPE_POSE3, int(pose_off[i]), 7, 6) for i in range(10)]
    # factor order: prior(P0), prior(P9), between(0,1) ... between(8,9)
    prior = (
        D.KIND_PRIOR_POSE3,
        np.array([[pose_off[0], pose_off[9]], [prior_start, prior_last], [si_prior, si_prior], [eps_off, eps_off]]),
        np.array([[0, 9]]),
        np.array([0, 1]),
    )
    n = 9
    between = (
        D.KIND_BETWEEN_POSE3,
        np.array([pose_off[:-1], pose_off[1:], np.full(n, identity), np.full(n, si_between), np.full(n, eps_off)]),
        np.array([np.arange(0, 9), np.arange(1, 10)]),
        np.arange(2, 11),
    )
    p = params if params is not None else D.default_params()
    return D.Problem(vb.data(), keys, [prior, between], params=p)


def rotation_smoothing(params=None):
    """test/symforce_optimizer_test.cc:183-236."""
    init = np.array(_kat_init()["rotation_smoothing"]).reshape(10, 4)
    vb = ValuesBuilder()
    off = vb.add_many(init)
    eps_off = vb.add([1e-15])
    prior_start = vb.add([0, 0, 0, 1])
    q = np.array([0.0, 0.0, np.sin(np.pi / 2), np.cos(np.pi / 2)])
    q = q / np.sqrt(np.sum(q * q))
    prior_last = vb.add(q)
    si_prior = vb.add(np.diag(np.full(3, 1 / 0.1)).reshape(-1, order="F"))
    si_between = vb.add(np.diag(np.full(3, 1 / 0.5)).reshape(-1, order="F"))
    identity = vb.add([0, 0, 0, 1])
    keys = [(D.TYPE_ROT3, int(off[i]), 4, 3) for i in range(10)]
    prior = (
        D.KIND_PRIOR_ROT3,
        np.array([[off[0], off[9]], [prior_start, prior_last], [si_prior, si_prior], [eps_off, eps_off]]),
        np.array([[0, 9]]),
        np.array([0, 1]),
    )
    n = 9
    between = (
        D.KIND_BETWEEN_ROT3,
        np.array([off[:-1], off[1:], np.full(n, identity), np.full(n, si_between), np.full(n, eps_off)]),
        np.array([np.arange(0, 9), np.arange(1, 10)]),
        np.arange(2, 11),
    )
    p = params if params is not None else D.default_params()
    return D.Problem(vb.data(), keys, [prior, between], params=p)


def frozen_keys(params=None):
    """test/symforce_optimizer_test.cc:276-331: 3 Rot3, all ordered pairs, R_0 frozen."""
    init = np.array(_kat_init()["frozen_keys"]).reshape(3, 4)
    vb = ValuesBuilder()
    off = vb.add_many(init)
    eps_off = vb.add([1e-15])
    si_between = vb.add(np.diag(np.full(3, 1 / 0.5)).reshape(-1, order="F"))
    identity = vb.add([0, 0, 0, 1])
    keys = [(D.TYPE_ROT3, int(off[i]), 4, 3) for i in (1, 2)]  # optimized: R_1, R_2
    key_of = {0: -1, 1: 0, 2: 1}
    pairs = [(i, j) for i in range(3) for j in range(3) if i != j]
    n = len(pairs)
    between = (
        D.KIND_BETWEEN_ROT3,
        np.array([[off[i] for i, _ in pairs], [off[j] for _, j in pairs], np.full(n, identity),
                  np.full(n, si_between), np.full(n, eps_off)]),
        np.array([[key_of[i] for i, _ in pairs], [key_of[j] for _, j in pairs]]),
        np.arange(n),
    )
    p = params if params is not None else D.default_params()
    return D.Problem(vb.data(), keys, [between], params=p)


def robot_3d_localization(params=None):
    """
    symforce/examples/robot_3d_localization/{common.h:24-62, run_dynamic_size.cc:25-70}; data from
    gen/measurements.cc via tests/golden/robot3d_measurements.json.
    """
    with open(os.path.join(_GOLDEN, "robot3d_measurements.json")) as f:
        m = json.load(f)
    P, L = m["num_poses"], m["num_landmarks"]
    body = np.array(m["body_t_landmark_measurements"]).reshape(P, L, 3)
    odom = np.array(m["odometry_relative_pose_measurements"]).reshape(P - 1, 7)
    odom[:, :4] /= np.linalg.norm(odom[:, :4], axis=1, keepdims=True)  # Pose3 ctor normalises
    land = np.array(m["landmark_positions"]).reshape(L, 3)
    vb = ValuesBuilder()
    pose_off = vb.add_many(np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (P, 1)))
    land_off = vb.add_many(land)
    sig_off = vb.add([0.05, 0.05, 0.05, 0.2, 0.2, 0.2])
    odom_off = vb.add_many(odom)
    msig_off = vb.add([0.1])
    body_off = vb.add_many(body.reshape(P * L, 3)).reshape(P, L)
    eps_off = vb.add([D.K_DEFAULT_EPSILON])
    keys = [(D.TYPE_POSE3, int(pose_off[i]), 7, 6) for i in range(P)]
    ii, jj = np.meshgrid(np.arange(P), np.arange(L), indexing="ij")
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    matching = (
        D.KIND_MATCHING,
        np.array([pose_off[ii], land_off[jj], body_off[ii, jj], np.full(P * L, msig_off)]),
        np.array([ii]),
        np.arange(P * L),
    )
    k = np.arange(P - 1)
    odometry = (
        D.KIND_ODOMETRY,
        np.array([pose_off[k], pose_off[k + 1], odom_off[k], np.full(P - 1, sig_off), np.full(P - 1, eps_off)]),
        np.array([k, k + 1]),
        P * L + k,
    )
    if params is None:
        params = D.default_params()
        params.initial_lambda = 1e4
        params.lambda_down_factor = 0.5
    return D.Problem(vb.data(), keys, [matching, odometry], params=params)


# ------------------------------------------------------------------------------------------------
# Synthetic BAL
# ------------------------------------------------------------------------------------------------
BAL_SHAPES = {
    "ladybug": dict(n_cams=49, n_pts=7776, n_obs=31843, window=12),
    "final": dict(n_cams=1778, n_pts=993923, n_obs=5001946, window=40),
    "tiny": dict(n_cams=6, n_pts=40, n_obs=150, window=2),
    "small": dict(n_cams=16, n_pts=600, n_obs=2400, window=4),
}


def bal_structure(n_cams, n_pts, n_obs, window, seed=0xBA1):
    """Returns (cam_idx, pt_idx) of the observations sorted by (camera, point)."""
    rng = np.random.default_rng(seed)
    W = min(window, (n_cams - 1) // 2)
    slots = 2 * W + 1
    n_extra = n_obs - 2 * n_pts
    assert n_extra >= 0
    mean = n_extra / n_pts
    k = rng.poisson(mean, n_pts)
    k = np.clip(k, 0, slots - 2)
    # fix up to the exact total
    diff = int(n_extra - k.sum())
    while diff != 0:
        step = 1 if diff > 0 else -1
        cand = np.flatnonzero((k + step >= 0) & (k + step <= slots - 2))
        if cand.shape[0] == 0:
            raise ValueError("n_obs=%d cannot be realised with %d points and tracks of 2..%d cameras" % (n_obs, n_pts, slots))
        pick = rng.choice(cand, size=min(abs(diff), cand.shape[0]), replace=False)
        k[pick] += step
        diff = int(n_extra - k.sum())
    k = k + 2
    centre = rng.integers(0, n_cams, n_pts)
    cam_chunks, pt_chunks = [], []
    chunk = 1 << 18
    for s in range(0, n_pts, chunk):
        e = min(n_pts, s + chunk)
        r = rng.random((e - s, slots), dtype=np.float32)
        order = np.argsort(r, axis=1)  # random permutation of window slots per point
        kk = k[s:e]
        mask = np.arange(slots)[None, :] < kk[:, None]
        rows = np.nonzero(mask)[0]
        sel = order[mask]
        cams = (centre[s:e][rows] + sel - W) % n_cams
        cam_chunks.append(cams.astype(np.int32))
        pt_chunks.append((rows + s).astype(np.int32))
    cam = np.concatenate(cam_chunks)
    pt = np.concatenate(pt_chunks)
    o = np.lexsort((pt, cam))
    return cam[o], pt[o]


def bal_problem(shape="ladybug", solver=D.SOLVER_SCHUR, params=None, seed_structure=0xBA1, seed_noise=0xBA2,
                n_cams=None, n_pts=None, n_obs=None, window=None, pt_range=None):
    """
    Synthetic BAL-shaped problem; keys/letters as the reference example
    (symforce/examples/bundle_adjustment_in_the_large/bundle_adjustment_in_the_large.cc:27-118):
    values = [pixels (obs order) | per camera (pose7, intrinsics3) | points | epsilon],
    factors in observation order sorted by (camera, point), optimized keys in lexical order
    c_0.., i_0.., p_0...  params default to DefaultOptimizerParams + DYNAMIC lambda (:133-135).

    pt_range=(lo, hi): keep only the observations of points lo..hi-1 (landmark sharding for
    multi-GPU) -- keys and values stay replicated.
    """
    if n_cams is None:
        s = BAL_SHAPES[shape]
        n_cams, n_pts, n_obs, window = s["n_cams"], s["n_pts"], s["n_obs"], s["window"]
    cam, pt = bal_structure(n_cams, n_pts, n_obs, window, seed_structure)
    rng = np.random.default_rng(seed_noise)
    # truth
    X = np.stack([rng.uniform(-10, 10, n_pts), rng.uniform(-10, 10, n_pts), rng.uniform(10, 30, n_pts)], axis=1)
    th = 2 * np.pi * np.arange(n_cams) / n_cams
    C = np.stack([5 * np.cos(th), 5 * np.sin(th), np.zeros(n_cams)], axis=1)
    look = np.array([0.0, 0.0, 20.0])[None, :] - C
    look /= np.linalg.norm(look, axis=1, keepdims=True)
    zc = -look  # camera looks down -z (BAL convention)
    up = np.array([0.0, 1.0, 0.0])
    xc = np.cross(np.tile(up, (n_cams, 1)), zc)
    xc /= np.linalg.norm(xc, axis=1, keepdims=True)
    yc = np.cross(zc, xc)
    R_wc = np.stack([xc, yc, zc], axis=2)  # columns = camera axes in world
    R_cw = np.transpose(R_wc, (0, 2, 1))
    q_cw = quat_from_matrix(R_cw)
    t_cw = -np.einsum("nij,nj->ni", R_cw, C)
    f = rng.uniform(500, 1500, n_cams)
    k1 = rng.normal(0, 1e-2, n_cams) * 1e-1
    k2 = rng.normal(0, 1e-3, n_cams) * 1e-2
    # observations
    pc = quat_rotate(q_cw[cam], X[pt]) + t_cw[cam]
    p = -pc[:, :2] / pc[:, 2:3]
    n2 = np.sum(p * p, axis=1)
    r = 1 + k1[cam] * n2 + k2[cam] * n2 * n2
    pix = (f[cam] * r)[:, None] * p + rng.normal(0, 0.5, (cam.shape[0], 2))
    # initial guess
    dq = quat_exp(rng.normal(0, 0.01, (n_cams, 3)))
    q0 = quat_mul(q_cw, dq)
    q0 /= np.linalg.norm(q0, axis=1, keepdims=True)
    t0 = t_cw + rng.normal(0, 0.05, (n_cams, 3))
    X0 = X + rng.normal(0, 0.1, (n_pts, 3))

    camblock = np.concatenate([q0, t0, f[:, None], k1[:, None], k2[:, None]], axis=1)  # [n_cams, 10]
    return _bal_from_arrays(cam, pt, pix, camblock, X0, solver, params, pt_range)


def _bal_from_arrays(cam, pt, pix, camblock, X0, solver=D.SOLVER_SCHUR, params=None, pt_range=None):
    """Flat BAL problem from observation lists (camera index, point index, pixel), per-camera
    [qx, qy, qz, qw, tx, ty, tz, f, k1, k2] and initial points; layout documented at bal_problem."""
    n_cams, n_pts = camblock.shape[0], X0.shape[0]
    vb = ValuesBuilder()
    n_all = cam.shape[0]
    pix_off = vb.add_many(pix)
    cam_off = vb.add_many(camblock)
    pose_off = cam_off
    intr_off = cam_off + 7
    pt_off = vb.add_many(X0)
    eps_off = vb.add([D.K_DEFAULT_EPSILON])

    keys = np.empty((2 * n_cams + n_pts, 4), dtype=np.int32)
    keys[:n_cams] = np.stack([np.full(n_cams, D.TYPE_POSE3), pose_off, np.full(n_cams, 7), np.full(n_cams, 6)], 1)
    keys[n_cams:2 * n_cams] = np.stack([np.full(n_cams, D.TYPE_VECTOR), intr_off, np.full(n_cams, 3), np.full(n_cams, 3)], 1)
    keys[2 * n_cams:] = np.stack([np.full(n_pts, D.TYPE_VECTOR), pt_off, np.full(n_pts, 3), np.full(n_pts, 3)], 1)

    sel = np.arange(n_all)
    if pt_range is not None:
        sel = np.flatnonzero((pt >= pt_range[0]) & (pt < pt_range[1]))
    cs, ps = cam[sel], pt[sel]
    n = sel.shape[0]
    batch = (
        D.KIND_SNAVELY,
        np.array([pose_off[cs], intr_off[cs], pt_off[ps], pix_off[sel], np.full(n, eps_off)]),
        np.array([cs, n_cams + cs, 2 * n_cams + ps]),
        np.arange(n),
    )
    if params is None:
        params = D.default_params()
        params.lambda_update_type = D.LAMBDA_DYNAMIC
    prob = D.Problem(vb.data(), keys, [batch], solver=solver,
                     schur_num_keys=n_pts if solver == D.SOLVER_SCHUR else 0, params=params)
    prob.meta = dict(n_cams=n_cams, n_pts=n_pts, n_obs=int(n), cam=cs, pt=ps, cam_off=int(cam_off[0]),
                     pt_off=int(pt_off[0]), pix_off=int(pix_off[0]))
    return prob


def read_bal(path, solver=D.SOLVER_SCHUR, params=None, pt_range=None):
    """
    A "Bundle Adjustment in the Large" text file (https://grail.cs.washington.edu/projects/bal/) as a flat problem,
    the way the reference example reads it (bundle_adjustment_in_the_large.cc:61-118): header `cameras points
    observations`; one `camera point x y` line per observation; 9 numbers per camera (Rodrigues rotation, translation,
    f, k1, k2; the pose is Pose3(Rot3::FromTangent(r), t)); 3 numbers per point.  Factors in file order, keys c_j, i_j, p_k.
    The whole file is whitespace-separated numbers, so it is parsed in one numpy call (5 M observations in seconds).
    """
    data = np.fromfile(path, sep=" ", dtype=np.float64)
    if data.shape[0] < 3:
        raise ValueError(f"{path}: not a BAL problem file")
    n_cams, n_pts, n_obs = (int(x) for x in data[:3])
    want = 3 + 4 * n_obs + 9 * n_cams + 3 * n_pts
    if data.shape[0] != want:
        raise ValueError(f"{path}: expected {want} numbers for {n_cams} cameras / {n_pts} points / {n_obs} observations, "
                         f"found {data.shape[0]}")
    obs = data[3:3 + 4 * n_obs].reshape(n_obs, 4)
    cam = obs[:, 0].astype(np.int32)
    pt = obs[:, 1].astype(np.int32)
    if n_obs and (cam.min() < 0 or cam.max() >= n_cams or pt.min() < 0 or pt.max() >= n_pts):
        raise ValueError(f"{path}: observation refers to a camera or point that is not in the file")
    cams = data[3 + 4 * n_obs:3 + 4 * n_obs + 9 * n_cams].reshape(n_cams, 9)
    X0 = data[3 + 4 * n_obs + 9 * n_cams:].reshape(n_pts, 3).copy()
    q = quat_exp(cams[:, :3])
    q /= np.linalg.norm(q, axis=1, keepdims=True)  # the Rot3 constructor normalises
    camblock = np.concatenate([q, cams[:, 3:]], axis=1)
    return _bal_from_arrays(cam, pt, obs[:, 2:4].copy(), camblock, X0, solver, params, pt_range)


def write_bal(path, prob):
    """Writes a flat BAL problem (bal_problem / read_bal layout) as a BAL text file; %.17g keeps every double."""
    from .geo import Rot3

    m = prob.meta
    n_cams, n_pts, n_obs = m["n_cams"], m["n_pts"], m["n_obs"]
    v = prob.values
    pix = v[m["pix_off"]:m["pix_off"] + 2 * n_obs].reshape(n_obs, 2)
    cams = v[m["cam_off"]:m["cam_off"] + 10 * n_cams].reshape(n_cams, 10)
    pts = v[m["pt_off"]:m["pt_off"] + 3 * n_pts].reshape(n_pts, 3)
    with open(path, "w") as f:
        f.write(f"{n_cams} {n_pts} {n_obs}\n")
        for c, p, (x, y) in zip(m["cam"], m["pt"], pix):
            f.write(f"{int(c)} {int(p)} {x:.17g} {y:.17g}\n")
        for row in cams:
            r = Rot3(row[:4]).to_tangent()
            for x in (*r, *row[4:]):
                f.write(f"{x:.17g}\n")
        for row in pts:
            for x in row:
                f.write(f"{x:.17g}\n")


# ------------------------------------------------------------------------------------------------
# Synthetic pose graph
# ------------------------------------------------------------------------------------------------
def pose_graph_problem(n_poses=100000, n_loops=20000, seed=0x51A4, params=None, ordering=D.ORDERING_METIS_SCALAR):
    """
    Synthetic Pose3 pose-graph SLAM (config E): a Manhattan-world trajectory (1 m steps, 90-degree
    turns, confined to a square region so that places are revisited), odometry edges (i, i+1) and
    `n_loops` loop closures between poses that are spatial neighbours (<= 5 m) but far apart in time;
    measurements = true relative pose (+) N(0, diag(0.01^2 rad, 0.05^2 m)); sqrt_info =
    diag(100,100,100,20,20,20) stored as one Matrix66 Values key per edge; one PriorFactorPose3 on P_0;
    initial guess = odometry dead reckoning.
    """
    rng = np.random.default_rng(seed)
    side = max(8, int(np.sqrt(n_poses) * 0.9))  # region size in metres: cells get revisited
    pos = np.zeros((n_poses, 2), dtype=np.int64)
    head = np.zeros(n_poses, dtype=np.int64)  # heading in quarter turns
    dirs = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]])
    turn = rng.choice([0, 0, 0, 0, 0, 0, 1, 3], size=n_poses)
    x, y, h = side // 2, side // 2, 0
    for i in range(n_poses):
        pos[i] = (x, y)
        head[i] = h
        h = (h + turn[i]) % 4
        nx, ny = x + dirs[h][0], y + dirs[h][1]
        tries = 0
        while not (0 <= nx < side and 0 <= ny < side) and tries < 4:
            h = (h + 1) % 4
            nx, ny = x + dirs[h][0], y + dirs[h][1]
            tries += 1
        x, y = nx, ny
    theta = head * (np.pi / 2)
    q = np.stack([np.zeros(n_poses), np.zeros(n_poses), np.sin(theta / 2), np.cos(theta / 2)], axis=1)
    t = np.stack([pos[:, 0].astype(float), pos[:, 1].astype(float), np.zeros(n_poses)], axis=1)
    # loop closures: spatial neighbours through a cell hash (5 m cells), far apart in time
    min_sep = min(100, max(1, n_poses // 4))
    cell = (pos // 5)
    keyc = cell[:, 0] * 100003 + cell[:, 1]
    order = np.argsort(keyc, kind="stable")
    sorted_keys = keyc[order]
    la, lb = [], []
    src = rng.permutation(n_poses)
    seen = set()
    for i in src:
        if len(la) >= n_loops:
            break
        lo = np.searchsorted(sorted_keys, keyc[i], side="left")
        hi = np.searchsorted(sorted_keys, keyc[i], side="right")
        cand = order[lo:hi]
        cand = cand[np.abs(cand - i) > min_sep]
        if cand.shape[0] == 0:
            continue
        j = int(cand[rng.integers(0, cand.shape[0])])
        a, b = (int(i), j) if i < j else (j, int(i))
        if (a, b) in seen:
            continue
        seen.add((a, b))
        la.append(a)
        lb.append(b)
    ea = np.concatenate([np.arange(n_poses - 1), np.array(la, dtype=np.int64)])
    eb = np.concatenate([np.arange(1, n_poses), np.array(lb, dtype=np.int64)])
    ne = ea.shape[0]
    # measurements: a_T_b = a^-1 b (+) noise
    qa_inv = q[ea] * np.array([-1, -1, -1, 1.0])
    q_ab = quat_mul(qa_inv, q[eb])
    t_ab = quat_rotate(qa_inv, t[eb] - t[ea])
    noise = np.concatenate([rng.normal(0, 0.01, (ne, 3)), rng.normal(0, 0.05, (ne, 3))], axis=1)
    q_meas = quat_mul(q_ab, quat_exp(noise[:, :3]))
    q_meas /= np.linalg.norm(q_meas, axis=1, keepdims=True)
    t_meas = t_ab + noise[:, 3:]
    meas = np.concatenate([q_meas, t_meas], axis=1)
    # init: dead reckoning along odometry
    q0 = np.empty_like(q)
    t0 = np.empty_like(t)
    q0[0] = q[0]
    t0[0] = t[0]
    for i in range(1, n_poses):
        a = q0[i - 1]
        b = q_meas[i - 1]
        qq = np.array([a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1],
                       a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0],
                       a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3],
                       a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]])
        q0[i] = qq / np.sqrt(qq @ qq)
        v = t_meas[i - 1]
        qv = a[:3]
        tt = 2.0 * np.cross(qv, v)
        t0[i] = t0[i - 1] + v + a[3] * tt + np.cross(qv, tt)
    vb = ValuesBuilder()
    pose_off = vb.add_many(np.concatenate([q0, t0], axis=1))
    sqrt_info = np.diag([100, 100, 100, 20, 20, 20.0]).reshape(-1, order="F")
    meas_off = vb.add_many(meas)
    si_off = vb.add_many(np.tile(sqrt_info, (ne, 1)))  # one Matrix66 Values key per edge
    prior_off = vb.add(np.concatenate([q[0], t[0]]))
    prior_si = vb.add((1e3 * np.eye(6)).reshape(-1, order="F"))
    eps_off = vb.add([D.K_DEFAULT_EPSILON])
    keys = np.stack([np.full(n_poses, D.TYPE_POSE3), pose_off, np.full(n_poses, 7), np.full(n_poses, 6)], 1)
    between = (
        D.KIND_BETWEEN_POSE3,
        np.array([pose_off[ea], pose_off[eb], meas_off, si_off, np.full(ne, eps_off)]),
        np.array([ea, eb]),
        np.arange(ne),
    )
    prior = (
        D.KIND_PRIOR_POSE3,
        np.array([[pose_off[0]], [prior_off], [prior_si], [eps_off]]),
        np.array([[0]]),
        np.array([ne]),
    )
    p = params if params is not None else D.default_params()
    prob = D.Problem(vb.data(), keys, [between, prior], params=p, ordering=ordering)
    prob.meta = dict(n_poses=n_poses, n_edges=int(ne), n_loops=len(la))
    return prob


# ------------------------------------------------------------------------------------------------
# BA example shape (config A): 2 views, 20 inverse-range landmarks, view 0 fixed
# ------------------------------------------------------------------------------------------------
def ba_example(seed=42, num_landmarks=20, params=None):
    """
    The bundle_adjustment example's problem structure (symforce/examples/bundle_adjustment/
    run_bundle_adjustment.cc:20-125, build_example_state.cc:25-140, example_utils/
    bundle_adjustment_util.h:119-136): keys v/c/T/s/l/P/S/m/M/W/u/C/e, relative-pose priors
    (BetweenFactorPose3) in both directions, an InverseRangeLandmarkPriorFactor and an
    InverseRangeLandmarkLinearGncFactor per landmark, VIEW 0 not optimized.  Inputs are drawn with
    numpy (the reference uses libstdc++'s RNG stream, which numpy cannot reproduce), so parity for
    this config is oracle-vs-GPU on these inputs plus the reference's acceptance check
    (final error < 10, SUCCESS).
    """
    rng = np.random.default_rng(seed)
    eps = 1e-10
    fx = fy = 740.0
    cx, cy = 639.5, 359.5
    q0 = quat_exp(rng.normal(0, 0.3, (1, 3)))[0]
    t0 = rng.normal(0, 1, 3)
    pert = np.array([0.1, -0.2, 0.1, 2.1, 0.4, -0.2]) * rng.normal(0, 0.3)
    q1 = quat_mul(q0, quat_exp(pert[None, :3])[0])
    q1 /= np.linalg.norm(q1)
    t1 = t0 + pert[3:]
    noise = 0.1 * rng.normal(0, 1, 6) * 0.3
    q1n = quat_mul(q1, quat_exp(noise[None, :3])[0])
    q1n /= np.linalg.norm(q1n)
    t1n = t1 + noise[3:]
    # correspondences: grid pixels in the source view, inverse ranges 1/U(2.5, 30)
    gx, gy = np.meshgrid(np.arange(100, 1200, 100), np.arange(100, 700, 100))
    grid = np.stack([gx.reshape(-1), gy.reshape(-1)], 1).astype(float)
    src = grid[rng.permutation(grid.shape[0])[:num_landmarks]]
    inv_range = 1.0 / rng.uniform(2.5, 30, num_landmarks)
    ray = np.stack([(src[:, 0] - cx) / fx, (src[:, 1] - cy) / fy, np.ones(num_landmarks)], 1)
    ray /= np.linalg.norm(ray, axis=1, keepdims=True)
    p_cam0 = ray / inv_range[:, None]
    p_w = quat_rotate(np.tile(q0, (num_landmarks, 1)), p_cam0) + t0
    q1_inv = q1 * np.array([-1, -1, -1, 1.0])
    p_cam1 = quat_rotate(np.tile(q1_inv, (num_landmarks, 1)), p_w - t1)
    tgt = np.stack([fx * p_cam1[:, 0] / p_cam1[:, 2] + cx, fy * p_cam1[:, 1] / p_cam1[:, 2] + cy], 1)
    tgt += 1.0 * rng.normal(0, 1, tgt.shape)
    range_pert = np.clip(1 + rng.normal(0, 0.5, num_landmarks), 0.5, 2.0)
    lm_init = 1.0 / ((1.0 / inv_range) * range_pert)
    # relative pose prior 0->1: between(view0, view1) (+) noise
    q0_inv = q0 * np.array([-1, -1, -1, 1.0])
    q01 = quat_mul(q0_inv, q1)
    t01 = quat_rotate(q0_inv[None, :], (t1 - t0)[None, :])[0]
    pn = 0.3 * rng.normal(0, 1, 6) * 0.1
    q01 = quat_mul(q01, quat_exp(pn[None, :3])[0])
    q01 /= np.linalg.norm(q01)
    t01 = t01 + pn[3:]

    vb = ValuesBuilder()
    e_off = vb.add([eps])
    scale_off = vb.add([10.0])
    mu_off = vb.add([0.0])
    v0 = vb.add(np.concatenate([q0, t0]))
    c0 = vb.add([fx, fy, cx, cy])
    v1 = vb.add(np.concatenate([q1n, t1n]))
    c1 = vb.add([fx, fy, cx, cy])
    ident = [0, 0, 0, 1, 0, 0, 0]
    T = {}
    S = {}
    for i in range(2):
        for j in range(2):
            T[i, j] = vb.add(ident)
            S[i, j] = vb.add(np.zeros(36))
    vals_T01 = np.concatenate([q01, t01])
    vb.chunks[-4 * 2 + 2] = vals_T01  # (0,1) pose prior
    vb.chunks[-4 * 2 + 3] = (np.eye(6) / 0.3).reshape(-1, order="F")
    lm = vb.add_many(lm_init[:, None])
    src_off = vb.add_many(src)
    tgt_off = vb.add_many(tgt)
    w_off = vb.add_many(np.ones((num_landmarks, 1)))
    prior_off = vb.add_many(inv_range[:, None])
    sig_off = vb.add_many(np.full((num_landmarks, 1), 100.0))
    # optimized keys, lexical: l_0.. ('l' < 'v'), then v_1
    keys = [(D.TYPE_VECTOR, int(lm[i]), 1, 1) for i in range(num_landmarks)] + [(D.TYPE_POSE3, int(v1), 7, 6)]
    kv1 = num_landmarks
    view_off = {0: v0, 1: v1}
    view_key = {0: -1, 1: kv1}
    pairs = [(0, 1), (1, 0)]
    between = (
        D.KIND_BETWEEN_POSE3,
        np.array([[view_off[i] for i, _ in pairs], [view_off[j] for _, j in pairs], [T[p] for p in pairs],
                  [S[p] for p in pairs], [e_off] * 2]),
        np.array([[view_key[i] for i, _ in pairs], [view_key[j] for _, j in pairs]]),
        np.array([0, 1]),
    )
    n = num_landmarks
    prior = (
        D.KIND_IRL_PRIOR,
        np.array([lm, prior_off, w_off, sig_off, np.full(n, e_off)]),
        np.array([np.arange(n)]),
        2 + np.arange(n),
    )
    gnc = (
        D.KIND_IRL_LINEAR_GNC,
        np.array([np.full(n, v0), np.full(n, c0), np.full(n, v1), np.full(n, c1), lm, src_off, tgt_off, w_off,
                  np.full(n, mu_off), np.full(n, scale_off), np.full(n, e_off)]),
        np.array([np.full(n, -1), np.full(n, kv1), np.arange(n)]),
        2 + n + np.arange(n),
    )
    if params is None:
        params = D.default_params()
        params.iterations = 50
        params.lambda_up_factor = 10.0
        params.lambda_down_factor = 0.1
        params.lambda_lower_bound = 1e-8
    prob = D.Problem(vb.data(), keys, [between, prior, gnc], params=params, epsilon=eps)
    prob.meta = dict(mu_off=int(mu_off), scale_off=int(scale_off), num_landmarks=num_landmarks)
    return prob


# ------------------------------------------------------------------------------------------------
# The reference's GNC test problem
# ------------------------------------------------------------------------------------------------
def gnc_test(params=None):
    """
    test/symforce_gnc_test.cc:23-57: x = ones(5), 20 samples y_i (3 outliers near 10, mt19937(42): tests/golden/
    kat_initial_values.json "gnc_test"), one gnc_factors::BarronFactor(x, y_i, u, e) each, x optimized, optimizer
    epsilon 1e-12.  Values order as the test fills them: x, e, y_0..y_19, then u (set by GncOptimizer::Optimize).
    """
    ys = np.array(_kat_init()["gnc_test"]).reshape(20, 5)
    vb = ValuesBuilder()
    x_off = vb.add(np.ones(5))
    e_off = vb.add([D.K_DEFAULT_EPSILON])
    y_off = vb.add_many(ys)
    mu_off = vb.add([0.0])
    n = 20
    keys = [(D.TYPE_VECTOR, x_off, 5, 5)]
    barron = (
        D.KIND_BARRON,
        np.array([np.full(n, x_off), y_off, np.full(n, mu_off), np.full(n, e_off)]),
        np.zeros((1, n), dtype=np.int32),
        np.arange(n),
    )
    p = params if params is not None else D.default_params()
    prob = D.Problem(vb.data(), keys, [barron], params=p, epsilon=1e-12)
    prob.meta = dict(mu_off=int(mu_off), x_off=int(x_off))
    return prob
