"""Synthetic graphs of the shapes BASELINE.json names (C1..C5).  Workload definitions only: they build a
:class:`g2o_b200.graph.Graph` on the host; no solver code here.

C1  ``ba_demo``      restates g2o/examples/ba/ba_demo.cpp:56-86,185-299 (glibc ``rand()``, unseeded)
C2  ``sphere``       restates g2o/examples/sphere/create_sphere.cpp:84-198 (libstdc++ mt19937 + normal_distribution)
C3  ``bal_venice``   BAL-shaped generator (ours, seeded): 1778 cameras / 993 923 points / 5 001 946 observations
C4  ``bal_large``    same generator, 10 000 / 4 000 000 / 20 000 000, banded co-visibility
C5  ``slam2d``       simulator2d-shaped 2-D SLAM (apps/g2o_simulator/test_simulator2d.cpp:78-250 noise model)
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from .graph import (EDGE_BAL, EDGE_PROJECT_XYZ2UV, EDGE_SE2, EDGE_SE2_POINT_XY, EDGE_SE3, EDGE_SE3_EXPMAP, EDGE_SE3_PROJECT_XYZ,
                    KERNEL_HUBER, VERTEX_CAM_BAL, VERTEX_POINT_BAL, VERTEX_POINT_XY, VERTEX_POINT_XYZ, VERTEX_SE2,
                    VERTEX_SE3, VERTEX_SE3_EXPMAP, Graph)


# ----------------------------------------------------------------------------------------------------
# C1: ba_demo
# ----------------------------------------------------------------------------------------------------
class _LibcRand:
    """glibc ``std::rand()`` with its default seed — ba_demo.cpp never calls ``srand``."""

    def __init__(self, seed: int = 1):
        self._libc = ctypes.CDLL("libc.so.6")
        self._libc.srand(seed)
        self.RAND_MAX = 2147483647

    def uniform_rand(self, lo: float, hi: float) -> float:            # ba_demo.cpp:60-62
        return lo + (self._libc.rand() / (self.RAND_MAX + 1.0)) * (hi - lo)

    def gauss_rand(self, mean: float, sigma: float) -> float:         # ba_demo.cpp:64-72 (polar Box-Muller, uses y)
        while True:
            x = -1.0 + 2.0 * self.uniform_rand(0.0, 1.0)
            y = -1.0 + 2.0 * self.uniform_rand(0.0, 1.0)
            r2 = x * x + y * y
            if not (r2 > 1.0 or r2 == 0.0):
                break
        return mean + sigma * y * math.sqrt(-2.0 * math.log(r2) / r2)


def ba_demo(pixel_noise: float = 1.0, outlier_ratio: float = 0.0, robust_kernel: bool = False, num_cameras: int = 15,
            num_points: int = 300, edge_type: int = EDGE_SE3_PROJECT_XYZ, seed: int = 1) -> Graph:
    """C1.  Cameras at t=(0.04 i - 1, 0, 0), identity rotation, camera 0 fixed; points uniform in
    [-1.5,1.5]x[-0.5,0.5]x[3,4]; f=1000, pp=(320,240), 640x480 image; point initial noise N(0,1) per axis.
    ``edge_type`` = EDGE_SE3_PROJECT_XYZ (the fork's ba_demo.cpp:268) or EDGE_PROJECT_XYZ2UV (ba_demo_block / upstream).
    C++ leaves the evaluation order of the three ``Sample::`` calls inside a ``Vector3d(...)`` constructor
    unspecified; we fix left-to-right."""
    rnd = _LibcRand(seed)
    true_points = [((rnd.uniform_rand(0., 1.) - 0.5) * 3, rnd.uniform_rand(0., 1.) - 0.5, rnd.uniform_rand(0., 1.) + 3)
                   for _ in range(num_points)]
    f, cx, cy = 1000., 320., 240.
    cam_t = [(i * 0.04 - 1., 0., 0.) for i in range(num_cameras)]

    v_id, v_type, v_fixed, v_marg, est = [], [], [], [], []
    for i, t in enumerate(cam_t):
        v_id.append(i); v_type.append(VERTEX_SE3_EXPMAP); v_fixed.append(1 if i < 1 else 0); v_marg.append(0)
        est.extend([t[0], t[1], t[2], 0., 0., 0., 1.])
    e_v0, e_v1, meas, prm = [], [], [], []
    point_id = num_cameras
    true_of_point = []

    def cam_map(t, X):   # types_six_dof_expmap.cpp:74-80 with identity rotation
        P = (X[0] + t[0], X[1] + t[1], X[2] + t[2])
        return (P[0] / P[2] * f + cx, P[1] / P[2] * f + cy)

    for i, X in enumerate(true_points):
        p_est = (X[0] + rnd.gauss_rand(0., 1.), X[1] + rnd.gauss_rand(0., 1.), X[2] + rnd.gauss_rand(0., 1.))
        zs = [cam_map(t, X) for t in cam_t]
        vis = [(z[0] >= 0 and z[1] >= 0 and z[0] < 640 and z[1] < 480) for z in zs]
        if sum(vis) >= 2:
            vidx = len(v_id)
            v_id.append(point_id); v_type.append(VERTEX_POINT_XYZ); v_fixed.append(0); v_marg.append(1)
            est.extend(p_est)
            for j, z in enumerate(zs):
                if not vis[j]:
                    continue
                sam = rnd.uniform_rand(0., 1.)
                if sam < outlier_ratio:
                    z = (float(int(rnd.uniform_rand(0, 640))), float(int(rnd.uniform_rand(0, 480))))
                z = (z[0] + rnd.gauss_rand(0., pixel_noise), z[1] + rnd.gauss_rand(0., pixel_noise))
                e_v0.append(vidx); e_v1.append(j); meas.extend(z)
                prm.extend([f, f, cx, cy] if edge_type == EDGE_SE3_PROJECT_XYZ else [f, cx, cy])
            true_of_point.append(i)
            point_id += 1
    ne = len(e_v0)
    g = Graph(v_id=v_id, v_type=v_type, v_fixed=v_fixed, v_marginalized=v_marg, v_estimate=est,
              e_type=np.full(ne, edge_type), e_v0=e_v0, e_v1=e_v1, e_measurement=meas,
              e_information=np.tile([1., 0., 0., 1.], ne), e_param=prm,
              e_kernel=np.full(ne, KERNEL_HUBER if robust_kernel else 0), name="ba_demo",
              meta={"true_points": np.array([true_points[i] for i in true_of_point]), "num_cameras": num_cameras})
    return g


# ----------------------------------------------------------------------------------------------------
# C2: create_sphere
# ----------------------------------------------------------------------------------------------------
class _StdNormalShared:
    """libstdc++ ``std::normal_distribution<double>`` (Marsaglia polar, keeps one saved value) driven by
    ``std::mt19937`` engines through ``generate_canonical<double,53>`` (two 32-bit draws, low word first).
    One instance is shared by all engines, exactly like the static ``_univariateSampler`` in
    g2o/stuff/sampler.cpp:31,41-45 — the saved value crosses between the two samplers."""

    def __init__(self):
        self.saved = 0.0
        self.available = False

    @staticmethod
    def engine(seed: int = 5489):
        bg = np.random.MT19937()
        bg._legacy_seeding(seed)          # init_genrand(seed) == std::mt19937(seed); 5489 is the default seed
        return bg

    @staticmethod
    def _canonical(bg) -> float:
        lo, hi = (int(v) for v in bg.random_raw(2))
        r = (lo + hi * 4294967296.0) / 18446744073709551616.0
        return r if r < 1.0 else math.nextafter(1.0, 0.0)

    def __call__(self, bg) -> float:
        if self.available:
            self.available = False
            return self.saved
        while True:
            x = 2.0 * self._canonical(bg) - 1.0
            y = 2.0 * self._canonical(bg) - 1.0
            r2 = x * x + y * y
            if not (r2 > 1.0 or r2 == 0.0):
                break
        mult = math.sqrt(-2.0 * math.log(r2) / r2)
        self.saved = x * mult
        self.available = True
        return y * mult


def _quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _iso_pack(R, t):
    return np.concatenate([np.asarray(R).ravel(order="F"), np.asarray(t)])


def sphere(nodes_per_level: int = 100, laps: int = 100, radius: float = 100.0,
           noise_translation=(0.01, 0.01, 0.01), noise_rotation=(0.005, 0.005, 0.005), fix_first: bool = True) -> Graph:
    """C2.  ``create_sphere -nodesPerLevel N -laps L``: N*L VertexSE3, (N*L-1) odometry + loop-closure EdgeSE3,
    information diag(1/sigma_t^2, 1/sigma_r^2), measurement noise from two default-seeded mt19937 engines,
    initial guess = odometry chaining (create_sphere.cpp:185-192).  Vertex 0 is fixed explicitly
    (``findGauge`` in the reference is unordered_map-order dependent, sparse_optimizer.cpp:118-137)."""
    n = nodes_per_level * laps
    Rs, ts = [], []
    idc = 0
    for f in range(laps):
        for k in range(nodes_per_level):
            idc += 1
            az = -math.pi + 2 * k * math.pi / nodes_per_level
            ay = -0.5 * math.pi + idc * math.pi / (laps * nodes_per_level)
            Rz = np.array([[math.cos(az), -math.sin(az), 0], [math.sin(az), math.cos(az), 0], [0, 0, 1]])
            Ry = np.array([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]])
            R = Rz @ Ry
            Rs.append(R); ts.append(R @ np.array([radius, 0, 0]))
    pairs = [(i - 1, i) for i in range(1, n)]
    n_odom = len(pairs)
    for f in range(1, laps):
        for nn in range(nodes_per_level):
            a = (f - 1) * nodes_per_level + nn
            for d in (-1, 0, 1):
                if f == laps - 1 and d == 1:
                    continue
                pairs.append((a, f * nodes_per_level + nn + d))
    info = np.zeros((6, 6))
    for i in range(3):
        info[i, i] = 1.0 / noise_translation[i] ** 2
        info[3 + i, 3 + i] = 1.0 / noise_rotation[i] ** 2
    nd = _StdNormalShared()
    gen_trans, gen_rot = nd.engine(), nd.engine()
    meas = np.zeros((len(pairs), 12))
    from scipy.spatial.transform import Rotation
    for e, (a, b) in enumerate(pairs):
        Rgt = Rs[a].T @ Rs[b]
        tgt = Rs[a].T @ (ts[b] - ts[a])
        qn = np.array([nd(gen_rot) * noise_rotation[i] for i in range(3)])
        qw = 1.0 - float(np.linalg.norm(qn))
        if qw < 0:
            qw = 0.0
        q = np.array([qn[0], qn[1], qn[2], qw]); q /= np.linalg.norm(q)
        tn = np.array([nd(gen_trans) * noise_translation[i] for i in range(3)])
        qgt = Rotation.from_matrix(Rgt).as_quat()          # x y z w
        rot = (Rotation.from_quat(qgt) * Rotation.from_quat(q)).as_quat()
        meas[e] = _iso_pack(_quat_to_R(rot / np.linalg.norm(rot)), tgt + tn)
    # odometry chaining: to = from * measurement (edge_se3.cpp:107-118)
    est = np.zeros((n, 12))
    Rc, tc = Rs[0], ts[0]
    est[0] = _iso_pack(Rc, tc)
    for e in range(n_odom):
        Rm = meas[e, :9].reshape(3, 3, order="F"); tm = meas[e, 9:]
        tc = Rc @ tm + tc
        Rc = Rc @ Rm
        est[e + 1] = _iso_pack(Rc, tc)
    ne = len(pairs)
    fixed = np.zeros(n, dtype=np.uint8)
    if fix_first:
        fixed[0] = 1
    return Graph(v_id=np.arange(n), v_type=np.full(n, VERTEX_SE3), v_fixed=fixed, v_marginalized=np.zeros(n),
                 v_estimate=est.ravel(), e_type=np.full(ne, EDGE_SE3), e_v0=[p[0] for p in pairs], e_v1=[p[1] for p in pairs],
                 e_measurement=meas.ravel(), e_information=np.tile(info.ravel(order="F"), ne), name="sphere",
                 meta={"nodes_per_level": nodes_per_level, "laps": laps})


def _se3quat_inverse_of_iso(iso12: np.ndarray) -> np.ndarray:
    """Isometry (R column-major 9, t 3)  ->  SE3Quat::toVector (t, qx qy qz qw) of its inverse."""
    R = iso12[:9].reshape(3, 3, order="F"); t = iso12[9:]
    Ri = R.T
    return np.concatenate([-Ri @ t, _r_to_quat(Ri.ravel(order="F"))])


def sphere_expmap(nodes_per_level: int = 16, laps: int = 8, **kw) -> Graph:
    """The sphere pose graph of C2 expressed with the sba types (types_six_dof_expmap.h:84-127): VertexSE3Expmap holds the world-to-body
    transform T_i = X_i^-1, EdgeSE3Expmap (v1 = i, v2 = j) has error log(T_j^-1 C T_i), so C = Z_ij^-1 for the EdgeSE3 measurement
    Z_ij = X_i^-1 X_j.  SE3Quat::log orders the error (rotation, translation): the information blocks are swapped accordingly."""
    g = sphere(nodes_per_level=nodes_per_level, laps=laps, **kw)
    est = np.concatenate([_se3quat_inverse_of_iso(v) for v in g.v_estimate.reshape(-1, 12)])
    meas = np.concatenate([_se3quat_inverse_of_iso(m) for m in g.e_measurement.reshape(-1, 12)])
    info = g.e_information.reshape(-1, 6, 6)
    perm = [3, 4, 5, 0, 1, 2]
    info = info[:, perm][:, :, perm]
    return Graph(v_id=g.v_id, v_type=np.full(g.n_vertices, VERTEX_SE3_EXPMAP), v_fixed=g.v_fixed, v_marginalized=g.v_marginalized, v_estimate=est,
                 e_type=np.full(g.n_edges, EDGE_SE3_EXPMAP), e_v0=g.e_v0, e_v1=g.e_v1, e_measurement=meas, e_information=info.ravel(),
                 name="sphere_expmap", meta=dict(g.meta))


def ba_demo_with_pose_constraints(sigma_t: float = 0.01, sigma_r: float = 0.005, seed: int = 3, **kw) -> Graph:
    """ba_demo plus an EdgeSE3Expmap between consecutive cameras (a visual-inertial style relative-pose prior): Hpp gets off-diagonal blocks
    next to the Schur complement of the points.  The cameras of ba_demo are T_i = (I, t_i); C = T_j T_i^-1 + noise."""
    g = ba_demo(**kw)
    rng = np.random.default_rng(seed)
    cams = np.flatnonzero(np.asarray(g.v_type) == VERTEX_SE3_EXPMAP)
    off = g.estimate_offsets()
    e_v0, e_v1, meas, info = [], [], [], []
    I6 = np.diag([1 / sigma_r ** 2] * 3 + [1 / sigma_t ** 2] * 3)
    for a, b in zip(cams[:-1], cams[1:]):
        ta, tb = g.v_estimate[off[a]:off[a] + 3], g.v_estimate[off[b]:off[b] + 3]
        w = rng.normal(0, sigma_r, 3); q = np.concatenate([0.5 * w, [1.0]]); q /= np.linalg.norm(q)
        e_v0.append(a); e_v1.append(b); meas.extend(list(tb - ta + rng.normal(0, sigma_t, 3)) + list(q)); info.extend(I6.ravel())
    n = len(e_v0)
    # cameras start off their true place so that the priors and the projections pull against each other
    est = g.v_estimate.copy()
    for c in cams[1:]:
        est[off[c]:off[c] + 3] += rng.normal(0, 0.02, 3)
    return Graph(v_id=g.v_id, v_type=g.v_type, v_fixed=g.v_fixed, v_marginalized=g.v_marginalized, v_estimate=est,
                 e_type=np.concatenate([g.e_type, np.full(n, EDGE_SE3_EXPMAP)]), e_v0=np.concatenate([g.e_v0, e_v0]), e_v1=np.concatenate([g.e_v1, e_v1]),
                 e_measurement=np.concatenate([g.e_measurement, meas]), e_information=np.concatenate([g.e_information, info]),
                 e_param=g.e_param, e_kernel=np.concatenate([g.e_kernel, np.zeros(n)]), e_kernel_delta=np.concatenate([g.e_kernel_delta, np.ones(n)]),
                 name="ba_demo_pose_constraints", meta=dict(g.meta))


# ----------------------------------------------------------------------------------------------------
# C3 / C4: BAL-shaped bundle adjustment
# ----------------------------------------------------------------------------------------------------
def _track_lengths(rng, n_points: int, n_obs: int, k_max: int) -> np.ndarray:
    """Discrete power-law track lengths on [2, k_max] whose sum is exactly ``n_obs``."""
    mean = n_obs / n_points
    if not (2.0 <= mean <= k_max):
        raise ValueError(f"mean track length {mean:.3f} is outside [2, {k_max}]")
    ks = np.arange(2, k_max + 1, dtype=np.float64)
    lo, hi = -4.0, 8.0
    for _ in range(80):                      # bisection on the exponent for the requested mean
        s = 0.5 * (lo + hi)
        p = ks ** (-s); p /= p.sum()
        if (p * ks).sum() > mean:
            lo = s
        else:
            hi = s
    k = rng.choice(ks.astype(np.int64), size=n_points, p=p)
    diff = int(n_obs - k.sum())
    for _ in range(100000):                   # exact total: nudge random points by +-1
        if diff == 0:
            break
        idx = rng.integers(0, n_points, size=abs(diff))
        if diff > 0:
            ok = idx[k[idx] < k_max]
            ok = np.unique(ok)
            k[ok] += 1
        else:
            ok = idx[k[idx] > 2]
            ok = np.unique(ok)
            k[ok] -= 1
        diff = int(n_obs - k.sum())
    if diff != 0:
        raise RuntimeError("could not reach the requested number of observations")
    return k


def bal_synthetic(n_cameras: int = 1778, n_points: int = 993_923, n_obs: int = 5_001_946, seed: int = 20260101,
                  k_max: int = 500, window_scale: float = 0.75, min_window: int = 8, huber_delta: float | None = 1.0,
                  outlier_fraction: float = 0.01, pixel_sigma: float = 1.0, name: str = "bal_venice") -> Graph:
    """C3 (defaults) / C4.  Cameras on a closed ring looking at the scene inside it (BAL model, bal_example.cpp:192-244: angle-axis,
    t, f, k1, k2; the camera looks down -z).  Every point has a centre camera and a track length k from a truncated
    power law (mean n_obs/n_points, max ``k_max``); its observers are k distinct cameras drawn from the ring window
    of half-width max(min_window, window_scale*k) around the centre.  Observations = exact projection + N(0, pixel_sigma) (+ a few gross outliers); initial
    cameras/points are perturbed ground truth.  Edges are added point by point (the order of a BAL file).
    Vertex ids: cameras 0..Nc-1, then points (bal_example.cpp:336-377); points are marginalized; Huber kernel with
    ``huber_delta`` on every edge when not None; no fixed vertex (as in bal_example)."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    nc = n_cameras
    k = _track_lengths(rng, n_points, n_obs, min(k_max, nc))
    half = np.maximum(min_window, np.ceil(window_scale * k)).astype(np.int64)
    half = np.minimum(half, (nc - 1) // 2)
    half = np.maximum(half, (k + 1) // 2)                   # window 2*half+1 >= k
    centre = rng.integers(0, nc, size=n_points)
    ptr = np.zeros(n_points + 1, dtype=np.int64)
    np.cumsum(k, out=ptr[1:])
    cam_of_obs = np.empty(n_obs, dtype=np.int64)
    # choose k distinct offsets in [-half, half] per point, grouped by (k, half) so it vectorises
    key = k * (nc + 1) + half
    order = np.argsort(key, kind="stable")
    bounds = np.flatnonzero(np.diff(key[order])) + 1
    for grp in np.split(order, bounds):
        kk, hh = int(k[grp[0]]), int(half[grp[0]])
        w = 2 * hh + 1
        for s in range(0, len(grp), 65536):
            g = grp[s:s + 65536]
            r = rng.random((len(g), w))
            sel = np.argpartition(r, kk - 1, axis=1)[:, :kk] if kk < w else np.tile(np.arange(w), (len(g), 1))
            sel = np.sort(sel, axis=1) - hh
            cams = (centre[g][:, None] + sel) % nc
            idx = ptr[g][:, None] + np.arange(kk)[None, :]
            cam_of_obs[idx.ravel()] = cams.ravel()
    pt_of_obs = np.repeat(np.arange(n_points), k)

    # --- ground-truth geometry: cameras on a ring of radius rho looking at the scene in the middle ---
    rho = 50.0
    phi = 2 * np.pi * np.arange(nc) / nc
    C = np.stack([rho * np.cos(phi), rho * np.sin(phi), np.zeros(nc)], axis=1) + rng.normal(0, 0.05, (nc, 3))
    d = np.stack([np.cos(phi), np.sin(phi), np.zeros(nc)], axis=1)
    xc = np.stack([-np.sin(phi), np.cos(phi), np.zeros(nc)], axis=1)
    yc = np.tile([0., 0., 1.], (nc, 1))
    R = np.stack([xc, yc, d], axis=1)                       # rows = camera axes; the camera looks down -z = -d (inward)
    R = Rotation.from_rotvec(rng.normal(0, 0.02, (nc, 3))).as_matrix() @ R
    tvec = -np.einsum("nij,nj->ni", R, C)
    rotvec = Rotation.from_matrix(R).as_rotvec()
    cams_true = np.concatenate([rotvec, tvec, rng.uniform(800, 1200, (nc, 1)), rng.normal(0, 1e-2, (nc, 1)),
                                rng.normal(0, 1e-3, (nc, 1))], axis=1)
    # points fill a disc of radius 0.35 rho, biased towards the side of their centre camera
    phi_p = 2 * np.pi * (centre + rng.uniform(-0.5, 0.5, n_points)) / nc
    rad = 0.35 * rho * np.sqrt(rng.uniform(0, 1, n_points))
    X = np.stack([rad * np.cos(phi_p), rad * np.sin(phi_p), rng.uniform(-0.2, 0.2, n_points) * rho], axis=1)
    depth = rho - rad

    # --- observations: exact BAL projection of the truth + pixel noise ---
    meas = _bal_project(cams_true[cam_of_obs], X[pt_of_obs])
    meas += rng.normal(0, pixel_sigma, meas.shape)
    if outlier_fraction > 0:
        n_out = int(outlier_fraction * n_obs)
        oi = rng.choice(n_obs, size=n_out, replace=False)
        meas[oi] += rng.normal(0, 25.0, (n_out, 2))

    cams0 = cams_true.copy()
    cams0[:, 0:3] += rng.normal(0, 1e-3, (nc, 3))
    cams0[:, 3:6] += rng.normal(0, 1e-2, (nc, 3))
    cams0[:, 6] *= 1 + rng.normal(0, 1e-3, nc)
    X0 = X + rng.normal(0, 1.0, X.shape) * (1e-2 * depth / 10.0)[:, None]

    nv = nc + n_points
    v_type = np.concatenate([np.full(nc, VERTEX_CAM_BAL), np.full(n_points, VERTEX_POINT_BAL)])
    marg = np.concatenate([np.zeros(nc), np.ones(n_points)])
    g = Graph(v_id=np.arange(nv), v_type=v_type, v_fixed=np.zeros(nv), v_marginalized=marg,
              v_estimate=np.concatenate([cams0.ravel(), X0.ravel()]),
              e_type=np.full(n_obs, EDGE_BAL), e_v0=cam_of_obs, e_v1=nc + pt_of_obs, e_measurement=meas.ravel(),
              e_information=np.tile([1., 0., 0., 1.], n_obs),
              e_kernel=np.full(n_obs, KERNEL_HUBER if huber_delta is not None else 0),
              e_kernel_delta=np.full(n_obs, huber_delta if huber_delta is not None else 1.0), name=name,
              meta={"n_cameras": nc, "n_points": n_points, "n_obs": n_obs, "k_max_drawn": int(k.max()),
                    "sum_pairs": int((k * (k + 1) // 2).sum()), "seed": seed})
    return g


def _bal_project(cam: np.ndarray, X: np.ndarray) -> np.ndarray:
    """Vectorised BAL projection used only to synthesise measurements."""
    w = cam[:, 0:3]
    th = np.linalg.norm(w, axis=1, keepdims=True)
    v = w / np.maximum(th, 1e-300)
    c, s = np.cos(th), np.sin(th)
    p = X * c + np.cross(v, X) * s + v * (np.sum(v * X, axis=1, keepdims=True)) * (1 - c)
    p = p + cam[:, 3:6]
    u = -p[:, 0:2] / p[:, 2:3]
    r2 = np.sum(u * u, axis=1, keepdims=True)
    return cam[:, 6:7] * (1 + cam[:, 7:8] * r2 + cam[:, 8:9] * r2 * r2) * u


def bal_venice(**kw) -> Graph:
    return bal_synthetic(**kw)


def bal_large(**kw) -> Graph:
    args = dict(n_cameras=10_000, n_points=4_000_000, n_obs=20_000_000, k_max=200, window_scale=0.5, min_window=8,
                name="bal_large", seed=20260102)
    args.update(kw)
    return bal_synthetic(**args)


def bal_small(n_cameras=12, n_points=200, n_obs=900, seed=7, **kw) -> Graph:
    args = dict(k_max=min(10, n_cameras), min_window=3, name="bal_small", outlier_fraction=0.02)
    args.update(kw)
    return bal_synthetic(n_cameras=n_cameras, n_points=n_points, n_obs=n_obs, seed=seed, **args)


def read_bal(path: str, huber_delta: float | None = None) -> Graph:
    """BAL text format reader, the format parsed at bal_example.cpp:336-414."""
    with open(path) as fh:
        tok = fh.read().split()
    nc, npnt, nobs = int(tok[0]), int(tok[1]), int(tok[2])
    o = np.array(tok[3:3 + 4 * nobs], dtype=np.float64).reshape(nobs, 4)
    rest = np.array(tok[3 + 4 * nobs:3 + 4 * nobs + 9 * nc + 3 * npnt], dtype=np.float64)
    cams, pts = rest[:9 * nc], rest[9 * nc:]
    nv = nc + npnt
    return Graph(v_id=np.arange(nv), v_type=np.concatenate([np.full(nc, VERTEX_CAM_BAL), np.full(npnt, VERTEX_POINT_BAL)]),
                 v_fixed=np.zeros(nv), v_marginalized=np.concatenate([np.zeros(nc), np.ones(npnt)]),
                 v_estimate=np.concatenate([cams, pts]), e_type=np.full(nobs, EDGE_BAL), e_v0=o[:, 0].astype(np.int32),
                 e_v1=nc + o[:, 1].astype(np.int32), e_measurement=o[:, 2:4].ravel(),
                 e_information=np.tile([1., 0., 0., 1.], nobs),
                 e_kernel=np.full(nobs, KERNEL_HUBER if huber_delta is not None else 0),
                 e_kernel_delta=np.full(nobs, huber_delta if huber_delta is not None else 1.0), name="bal_file")


def write_bal(g: Graph, path: str) -> None:
    nc = int((g.v_type == VERTEX_CAM_BAL).sum()); npnt = int((g.v_type == VERTEX_POINT_BAL).sum())
    with open(path, "w") as fh:
        fh.write(f"{nc} {npnt} {g.n_edges}\n")
        m = g.e_measurement.reshape(-1, 2)
        for i in range(g.n_edges):
            fh.write(f"{int(g.e_v0[i])} {int(g.e_v1[i]) - nc} {float(m[i, 0])!r} {float(m[i, 1])!r}\n")
        for v in g.v_estimate:
            fh.write(f"{float(v)!r}\n")


# ----------------------------------------------------------------------------------------------------
# C5: 2-D SLAM with landmarks
# ----------------------------------------------------------------------------------------------------
def slam2d(n_poses: int = 100_000, n_landmarks: int = 20_000, world_size: float = 250.0, seed: int = 5,
           huber_delta: float | None = 1.0, sensor_range: float = 5.0, fov: float = 0.75 * math.pi,
           marginalize_landmarks: bool = True) -> Graph:
    """C5.  A robot random-walks on a Manhattan grid inside a ``world_size`` square (1 m steps, 90-degree turns), with
    odometry EdgeSE2 (information diag(500,500,5000)) and EdgeSE2PointXY observations (information 1000*I) of the
    landmarks within ``sensor_range`` and ``fov`` — the noise model of test_simulator2d.cpp:128-158.  As in the
    simulator, landmarks get ids 0..nl-1 and poses follow (apps/g2o_simulator/simulator.cpp:89-99); landmarks never
    observed have no edges and so are inactive.  The initial guess is odometry chaining for poses and the first
    observation for landmarks (``g2o -guessOdometry``); pose 0 is fixed.  The simulator's own edge order is
    pointer-dependent (sensor_pointxy.cpp:78-90), so this is a seeded restatement of the shape, not of its bits."""
    rng = np.random.default_rng(seed)
    lm = rng.uniform(-world_size / 2, world_size / 2, (n_landmarks, 2))
    # ground-truth trajectory
    gt = np.zeros((n_poses, 3))
    x, y, th = 0.0, 0.0, 0.0
    turn = rng.random(n_poses)
    for i in range(1, n_poses):
        if turn[i] < 0.15:
            th += math.pi / 2 if turn[i] < 0.075 else -math.pi / 2
        nx, ny = x + math.cos(th), y + math.sin(th)
        if abs(nx) > world_size / 2 or abs(ny) > world_size / 2:
            th += math.pi
            nx, ny = x + math.cos(th), y + math.sin(th)
        x, y = nx, ny
        th = (th + math.pi) % (2 * math.pi) - math.pi
        gt[i] = (x, y, th)
    # odometry measurements
    c, s = np.cos(gt[:-1, 2]), np.sin(gt[:-1, 2])
    dx, dy = gt[1:, 0] - gt[:-1, 0], gt[1:, 1] - gt[:-1, 1]
    odo = np.stack([c * dx + s * dy, -s * dx + c * dy, (gt[1:, 2] - gt[:-1, 2] + np.pi) % (2 * np.pi) - np.pi], axis=1)
    odo += rng.normal(0, 1, odo.shape) * np.array([1 / math.sqrt(500), 1 / math.sqrt(500), 1 / math.sqrt(5000)])
    # observations through a uniform grid
    cell = sensor_range
    gx = np.floor((lm[:, 0] + world_size / 2) / cell).astype(np.int64)
    gy = np.floor((lm[:, 1] + world_size / 2) / cell).astype(np.int64)
    ncell = int(math.ceil(world_size / cell)) + 1
    cell_id = gx * ncell + gy
    order = np.argsort(cell_id, kind="stable")
    starts = np.searchsorted(cell_id[order], np.arange(ncell * ncell + 1))
    obs_pose, obs_lm = [], []
    pgx = np.floor((gt[:, 0] + world_size / 2) / cell).astype(np.int64)
    pgy = np.floor((gt[:, 1] + world_size / 2) / cell).astype(np.int64)
    for ddx in (-1, 0, 1):
        for ddy in (-1, 0, 1):
            cx_, cy_ = pgx + ddx, pgy + ddy
            ok = (cx_ >= 0) & (cx_ < ncell) & (cy_ >= 0) & (cy_ < ncell)
            cid = np.where(ok, cx_ * ncell + cy_, 0)
            cnt = np.where(ok, starts[cid + 1] - starts[cid], 0)
            tot = int(cnt.sum())
            if tot == 0:
                continue
            pidx = np.repeat(np.arange(n_poses), cnt)
            off = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            lidx = order[np.repeat(starts[cid], cnt) + off]
            obs_pose.append(pidx); obs_lm.append(lidx)
    obs_pose = np.concatenate(obs_pose) if obs_pose else np.zeros(0, dtype=np.int64)     # n_landmarks = 0: an odometry-only pose chain
    obs_lm = np.concatenate(obs_lm) if obs_lm else np.zeros(0, dtype=np.int64)
    rel = lm[obs_lm] - gt[obs_pose, :2]
    cp, sp = np.cos(gt[obs_pose, 2]), np.sin(gt[obs_pose, 2])
    local = np.stack([cp * rel[:, 0] + sp * rel[:, 1], -sp * rel[:, 0] + cp * rel[:, 1]], axis=1)
    keep = (np.sum(local ** 2, axis=1) <= sensor_range ** 2) & (np.abs(np.arctan2(local[:, 1], local[:, 0])) <= fov / 2)
    obs_pose, obs_lm, local = obs_pose[keep], obs_lm[keep], local[keep]
    o2 = np.lexsort((obs_lm, obs_pose))
    obs_pose, obs_lm, local = obs_pose[o2], obs_lm[o2], local[o2]
    zl = local + rng.normal(0, 1 / math.sqrt(1000), local.shape)
    # interleave edges pose by pose: odometry into pose i, then that pose's observations
    n_odo, n_obs = n_poses - 1, len(obs_pose)
    key_odo = np.arange(1, n_poses) * 2
    key_obs = obs_pose * 2 + 1
    keys = np.concatenate([key_odo, key_obs])
    eo = np.argsort(keys, kind="stable")
    e_type = np.concatenate([np.full(n_odo, EDGE_SE2), np.full(n_obs, EDGE_SE2_POINT_XY)])[eo]
    v0 = np.concatenate([n_landmarks + np.arange(0, n_poses - 1), n_landmarks + obs_pose])[eo]
    v1 = np.concatenate([n_landmarks + np.arange(1, n_poses), obs_lm])[eo]
    is_odo = (e_type == EDGE_SE2)
    meas_parts = np.zeros((len(eo), 3)); info_parts = np.zeros((len(eo), 9))
    src = np.concatenate([np.arange(n_odo), np.arange(n_obs)])[eo]
    meas_parts[is_odo] = odo[src[is_odo]]
    meas_parts[~is_odo, :2] = zl[src[~is_odo]]
    info_parts[is_odo] = np.diag([500., 500., 5000.]).ravel()
    info_parts[~is_odo, :4] = (1000. * np.eye(2)).ravel()
    width_m = np.where(is_odo, 3, 2); width_i = np.where(is_odo, 9, 4)
    mask_m = np.arange(3)[None, :] < width_m[:, None]; mask_i = np.arange(9)[None, :] < width_i[:, None]
    # initial guess
    est_p = np.zeros((n_poses, 3))
    for i in range(1, n_poses):
        xx, yy, tt = est_p[i - 1]
        cc, ss = math.cos(tt), math.sin(tt)
        est_p[i] = (xx + cc * odo[i - 1, 0] - ss * odo[i - 1, 1], yy + ss * odo[i - 1, 0] + cc * odo[i - 1, 1],
                    (tt + odo[i - 1, 2] + math.pi) % (2 * math.pi) - math.pi)
    est_l = np.zeros((n_landmarks, 2))
    first = np.full(n_landmarks, -1, dtype=np.int64)
    first[obs_lm[::-1]] = np.arange(n_obs)[::-1]
    seen = first >= 0
    fp = obs_pose[first[seen]]
    cc, ss = np.cos(est_p[fp, 2]), np.sin(est_p[fp, 2])
    zz = zl[first[seen]]
    est_l[seen] = np.stack([est_p[fp, 0] + cc * zz[:, 0] - ss * zz[:, 1], est_p[fp, 1] + ss * zz[:, 0] + cc * zz[:, 1]], axis=1)
    nv = n_landmarks + n_poses
    fixed = np.zeros(nv, dtype=np.uint8); fixed[n_landmarks] = 1
    marg = np.concatenate([np.full(n_landmarks, 1 if marginalize_landmarks else 0), np.zeros(n_poses)])
    ne = len(eo)
    return Graph(v_id=np.arange(nv), v_type=np.concatenate([np.full(n_landmarks, VERTEX_POINT_XY), np.full(n_poses, VERTEX_SE2)]),
                 v_fixed=fixed, v_marginalized=marg, v_estimate=np.concatenate([est_l.ravel(), est_p.ravel()]),
                 e_type=e_type, e_v0=v0, e_v1=v1, e_measurement=meas_parts[mask_m], e_information=info_parts[mask_i],
                 e_kernel=np.full(ne, KERNEL_HUBER if huber_delta is not None else 0),
                 e_kernel_delta=np.full(ne, huber_delta if huber_delta is not None else 1.0), name="slam2d",
                 meta={"n_poses": n_poses, "n_landmarks": n_landmarks, "n_observations": int(n_obs)})


# ----------------------------------------------------------------------------------------------------
# .g2o text files (optimizable_graph.cpp:397-640 and the per-type read/write of the reference)
# ----------------------------------------------------------------------------------------------------
def _r_to_quat(R: np.ndarray) -> np.ndarray:
    """Column-major 3x3 (9 values) -> unit quaternion (x, y, z, w) with Eigen's branch structure."""
    m = np.asarray(R, dtype=np.float64).reshape(3, 3).T
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = math.sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t
        q[0] = (m[2, 1] - m[1, 2]) * t; q[1] = (m[0, 2] - m[2, 0]) * t; q[2] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[i] = 0.5 * t; t = 0.5 / t
        q[3] = (m[k, j] - m[j, k]) * t; q[j] = (m[j, i] + m[i, j]) * t; q[k] = (m[k, i] + m[i, k]) * t
    return q / np.linalg.norm(q)


def write_g2o(g: Graph, path: str) -> None:
    """VERTEX_SE2 / VERTEX_XY / VERTEX_SE3:QUAT / EDGE_SE2 / EDGE_SE2_XY / EDGE_SE3:QUAT and FIX lines, 17 significant digits."""
    from .graph import EDGE_DIM, EDGE_MEAS_DIM, VERTEX_ESTIMATE_DIM
    est_off = np.concatenate([[0], np.cumsum(VERTEX_ESTIMATE_DIM[g.v_type])])
    meas_off = np.concatenate([[0], np.cumsum(EDGE_MEAS_DIM[g.e_type])])
    info_off = np.concatenate([[0], np.cumsum(EDGE_DIM[g.e_type] ** 2)])
    f = lambda xs: " ".join(repr(float(x)) for x in xs)
    with open(path, "w") as fh:
        for i in range(g.n_vertices):
            x = g.v_estimate[est_off[i]:est_off[i + 1]]; t = int(g.v_type[i]); vid = int(g.v_id[i])
            if t == VERTEX_SE2:
                fh.write(f"VERTEX_SE2 {vid} {f(x)}\n")
            elif t == VERTEX_POINT_XY:
                fh.write(f"VERTEX_XY {vid} {f(x)}\n")
            elif t == VERTEX_SE3:
                fh.write(f"VERTEX_SE3:QUAT {vid} {f(x[9:12])} {f(_r_to_quat(x[:9]))}\n")
            else:
                raise ValueError(f"vertex type {t} has no .g2o writer here")
            if g.v_fixed[i]:
                fh.write(f"FIX {vid}\n")
        for k in range(g.n_edges):
            z = g.e_measurement[meas_off[k]:meas_off[k + 1]]; t = int(g.e_type[k]); d = int(EDGE_DIM[t])
            info = g.e_information[info_off[k]:info_off[k + 1]].reshape(d, d)
            upper = [info[i, j] for i in range(d) for j in range(i, d)]
            a, b = int(g.v_id[g.e_v0[k]]), int(g.v_id[g.e_v1[k]])
            if t == EDGE_SE2:
                fh.write(f"EDGE_SE2 {a} {b} {f(z)} {f(upper)}\n")
            elif t == EDGE_SE2_POINT_XY:
                fh.write(f"EDGE_SE2_XY {a} {b} {f(z)} {f(upper)}\n")
            elif t == EDGE_SE3:
                fh.write(f"EDGE_SE3:QUAT {a} {b} {f(z[9:12])} {f(_r_to_quat(z[:9]))} {f(upper)}\n")
            else:
                raise ValueError(f"edge type {t} has no .g2o writer here")
