"""Deterministic BAL-shaped synthetic bundle-adjustment problems.

No dataset ships with the reference (`data/` is git-ignored, fetched by crates/apex-io/src/utils.rs:209-248)
and there is no network, so every BASELINE.json config is generated: points in a unit ball, cameras on a
closed ring looking inward, track lengths and camera degrees heavy-tailed with ring locality, pixel noise
N(0, 0.5^2) + 2 % gross outliers, initial values = truth perturbed. BAL convention: p_cam = R p + t with
z < 0 in front (bal_pinhole.rs:154-156); Pinhole / Kannala-Brandt / Double-Sphere look down +z.
"""
from __future__ import annotations

import numpy as np

from . import _ffi as F
from .context import BAProblem

# name -> (ncam, npts, mean track length, camera model, loss, config index)
SHAPES = {
    "ladybug49": (49, 7776, 31843 / 7776, F.CAM_BAL, (F.LOSS_HUBER, 1.0), 1),
    "trafalgar257": (257, 65132, 225911 / 65132, F.CAM_BAL, (F.LOSS_HUBER, 1.0), 2),
    "venice1778": (1778, 993923, 5001946 / 993923, F.CAM_BAL, (F.LOSS_HUBER, 1.0), 3),
    "kb2000": (2000, 1000000, 6.0, F.CAM_KANNALA_BRANDT, (F.LOSS_CAUCHY, 1.0), 4),
    "ds2000": (2000, 1000000, 6.0, F.CAM_DOUBLE_SPHERE, (F.LOSS_CAUCHY, 1.0), 4),
    "final13682": (13682, 4456117, 28987644 / 4456117, F.CAM_BAL, (F.LOSS_HUBER, 1.0), 5),
}


def quat_from_matrix(R):
    """Unit quaternions (w,x,y,z) from rotation matrices [n,3,3] (Shepperd)."""
    R = np.asarray(R)
    n = R.shape[0]
    q = np.empty((n, 4))
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    cand = np.stack([tr, R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]], axis=1)
    best = np.argmax(cand, axis=1)
    for b in range(4):
        m = best == b
        if not m.any():
            continue
        Rm = R[m]
        if b == 0:
            s = np.sqrt(1.0 + tr[m]) * 2
            q[m] = np.stack([0.25 * s, (Rm[:, 2, 1] - Rm[:, 1, 2]) / s, (Rm[:, 0, 2] - Rm[:, 2, 0]) / s, (Rm[:, 1, 0] - Rm[:, 0, 1]) / s], 1)
        elif b == 1:
            s = np.sqrt(1.0 + Rm[:, 0, 0] - Rm[:, 1, 1] - Rm[:, 2, 2]) * 2
            q[m] = np.stack([(Rm[:, 2, 1] - Rm[:, 1, 2]) / s, 0.25 * s, (Rm[:, 0, 1] + Rm[:, 1, 0]) / s, (Rm[:, 0, 2] + Rm[:, 2, 0]) / s], 1)
        elif b == 2:
            s = np.sqrt(1.0 + Rm[:, 1, 1] - Rm[:, 0, 0] - Rm[:, 2, 2]) * 2
            q[m] = np.stack([(Rm[:, 0, 2] - Rm[:, 2, 0]) / s, (Rm[:, 0, 1] + Rm[:, 1, 0]) / s, 0.25 * s, (Rm[:, 1, 2] + Rm[:, 2, 1]) / s], 1)
        else:
            s = np.sqrt(1.0 + Rm[:, 2, 2] - Rm[:, 0, 0] - Rm[:, 1, 1]) * 2
            q[m] = np.stack([(Rm[:, 1, 0] - Rm[:, 0, 1]) / s, (Rm[:, 0, 2] + Rm[:, 2, 0]) / s, (Rm[:, 1, 2] + Rm[:, 2, 1]) / s, 0.25 * s], 1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    return q


def quat_to_matrix(q):
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = w * w + x * x - y * y - z * z; R[:, 0, 1] = 2 * (x * y - w * z); R[:, 0, 2] = 2 * (w * y + x * z)
    R[:, 1, 0] = 2 * (w * z + x * y); R[:, 1, 1] = w * w - x * x + y * y - z * z; R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y); R[:, 2, 1] = 2 * (w * x + y * z); R[:, 2, 2] = w * w - x * x - y * y + z * z
    return R


def quat_mul(a, b):
    aw, ax, ay, az = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    bw, bx, by, bz = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], 1)


def axis_angle_to_quat(aa):
    """bin/bundle_adjustment.rs:200-208: angle < 1e-10 -> identity, else from_axis_angle."""
    aa = np.asarray(aa, dtype=np.float64).reshape(-1, 3)
    ang = np.linalg.norm(aa, axis=1)
    q = np.zeros((aa.shape[0], 4))
    q[:, 0] = 1.0
    m = ang >= 1e-10
    ax = aa[m] / ang[m, None]
    q[m, 0] = np.cos(ang[m] / 2)
    q[m, 1:] = ax * np.sin(ang[m] / 2)[:, None]
    return q


def quat_to_axis_angle(q):
    q = np.asarray(q, dtype=np.float64).reshape(-1, 4).copy()
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    v = np.linalg.norm(q[:, 1:], axis=1)
    ang = 2 * np.arctan2(v, q[:, 0])
    out = np.zeros((q.shape[0], 3))
    m = v > 1e-300
    out[m] = q[m, 1:] / v[m, None] * ang[m, None]
    return out


def project(model, intr, pc):
    """numpy restatement of CameraModel::project for generating measurements; returns uv, valid."""
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    if model == F.CAM_BAL:
        valid = z < -1e-6
        zz = np.where(valid, z, -1.0)
        xn, yn = -x / zz, -y / zz
        r2 = xn * xn + yn * yn
        d = 1.0 + intr[:, 1] * r2 + intr[:, 2] * r2 * r2
        return np.stack([intr[:, 0] * xn * d, intr[:, 0] * yn * d], 1), valid
    if model == F.CAM_PINHOLE:
        valid = z >= 1e-6
        zz = np.where(valid, z, 1.0)
        return np.stack([intr[:, 0] * x / zz + intr[:, 2], intr[:, 1] * y / zz + intr[:, 3]], 1), valid
    if model == F.CAM_KANNALA_BRANDT:
        valid = z > np.finfo(np.float64).eps
        r = np.sqrt(x * x + y * y)
        th = np.arctan2(r, z)
        t2 = th * th
        thd = th * (1 + t2 * (intr[:, 4] + t2 * (intr[:, 5] + t2 * (intr[:, 6] + t2 * intr[:, 7]))))
        rr = np.where(r < 1e-6, 1.0, r)
        s = np.where(r < 1e-6, 1.0 / np.where(valid, z, 1.0), thd / rr)
        return np.stack([intr[:, 0] * x * s + intr[:, 2], intr[:, 1] * y * s + intr[:, 3]], 1), valid
    if model == F.CAM_DOUBLE_SPHERE:
        xi, al = intr[:, 4], intr[:, 5]
        r2 = x * x + y * y
        d1 = np.sqrt(r2 + z * z)
        w1 = np.where(al > 0.5, (1 - al) / al, al / (1 - al))
        w2 = (w1 + xi) / np.sqrt(2 * w1 * xi + xi * xi + 1)
        k = xi * d1 + z
        d2 = np.sqrt(r2 + k * k)
        den = al * d2 + (1 - al) * k
        valid = (z > -w2 * d1) & (den >= 1e-6)
        den = np.where(valid, den, 1.0)
        return np.stack([intr[:, 0] * x / den + intr[:, 2], intr[:, 1] * y / den + intr[:, 3]], 1), valid
    if model == F.CAM_RADTAN:
        valid = z >= 1e-6
        zz = np.where(valid, z, 1.0)
        xp, yp = x / zz, y / zz
        r2 = xp * xp + yp * yp
        k1, k2, p1, p2, k3 = (intr[:, i] for i in range(4, 9))
        radial = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 ** 3
        dx = 2 * p1 * xp * yp + p2 * (r2 + 2 * xp * xp)
        dy = p1 * (r2 + 2 * yp * yp) + 2 * p2 * xp * yp
        return np.stack([intr[:, 0] * (radial * xp + dx) + intr[:, 2], intr[:, 1] * (radial * yp + dy) + intr[:, 3]], 1), valid
    if model in (F.CAM_UCM, F.CAM_EUCM):
        al = intr[:, 4]
        beta = intr[:, 5] if model == F.CAM_EUCM else np.ones_like(al)
        r2 = x * x + y * y
        d = np.sqrt(beta * r2 + z * z)
        den = al * d + (1 - al) * z
        if model == F.CAM_UCM:
            w = np.where(al <= 0.5, al / (1 - al), (1 - al) / al)
            valid = (z > -w * d) & (den >= 1e-6)
        else:
            valid = (den >= 1e-6) & ~((al > 0.5) & (z < den * (al - 1) / np.where(al > 0.5, 2 * al - 1, 1.0)))
        den = np.where(valid, den, 1.0)
        return np.stack([intr[:, 0] * x / den + intr[:, 2], intr[:, 1] * y / den + intr[:, 3]], 1), valid
    if model == F.CAM_FOV:
        valid = z >= np.sqrt(np.finfo(np.float64).eps)
        zz = np.where(valid, z, 1.0)
        r = np.sqrt(x * x + y * y)
        w = intr[:, 4]
        m2t = 2 * np.tan(w / 2)
        rr = np.where(r > 1e-6, r, 1.0)
        rd = np.where(r > 1e-6, np.arctan(m2t * rr / zz) / (rr * w), m2t / w)
        return np.stack([intr[:, 0] * x * rd + intr[:, 2], intr[:, 1] * y * rd + intr[:, 3]], 1), valid
    if model == F.CAM_FTHETA:
        valid = z >= 1e-6
        d = np.sqrt(x * x + y * y + z * z)
        th = np.arccos(np.clip(z / np.where(d > 0, d, 1.0), -1.0, 1.0))
        f = th * (intr[:, 2] + th * (intr[:, 3] + th * (intr[:, 4] + th * intr[:, 5])))
        rp = np.sqrt(x * x + y * y)
        rr = np.where(rp < 1e-6, 1.0, rp)
        s_ = np.where(rp < 1e-6, 0.0, f / rr)
        return np.stack([intr[:, 0] + s_ * x, intr[:, 1] + s_ * y], 1), valid
    raise ValueError(f"camera model {model} not supported by the generator")


def _ring_cameras(rng, ncam, model, radius=5.0):
    phi = 2 * np.pi * (np.arange(ncam) + 0.15 * rng.standard_normal(ncam)) / ncam
    C = np.stack([radius * np.cos(phi), radius * np.sin(phi), 0.6 * np.sin(3 * phi) + 0.1 * rng.standard_normal(ncam)], 1)
    C *= (1.0 + 0.08 * rng.standard_normal(ncam))[:, None]
    target = 0.15 * rng.standard_normal((ncam, 3))
    fwd = target - C
    fwd /= np.linalg.norm(fwd, axis=1, keepdims=True)
    zc = -fwd if model == F.CAM_BAL else fwd         # camera z axis in world coordinates
    up = np.array([0.0, 0.0, 1.0])[None, :] + 0.05 * rng.standard_normal((ncam, 3))
    xc = np.cross(up, zc)
    xc /= np.linalg.norm(xc, axis=1, keepdims=True)
    yc = np.cross(zc, xc)
    R = np.stack([xc, yc, zc], axis=1)               # rows = camera axes => p_cam = R (p - C)
    t = -np.einsum("nij,nj->ni", R, C)
    return R, t


def _intrinsics(rng, ncam, model):
    if model == F.CAM_BAL:
        return np.stack([rng.uniform(400, 2000, ncam), 1e-3 * rng.standard_normal(ncam), 1e-4 * rng.standard_normal(ncam)], 1)
    j = lambda v: v * (1.0 + 0.02 * rng.standard_normal(ncam))
    if model == F.CAM_PINHOLE:
        return np.stack([j(500.0), j(500.0), j(320.0), j(240.0)], 1)
    if model == F.CAM_KANNALA_BRANDT:  # tests/camera_kannala_brandt_integration.rs:50-63
        return np.stack([j(200.0), j(200.0), j(300.0), j(200.0), j(0.5), j(0.1), 0.001 * rng.standard_normal(ncam), 0.0001 * rng.standard_normal(ncam)], 1)
    if model == F.CAM_DOUBLE_SPHERE:   # tests/camera_double_sphere_integration.rs:52-63
        return np.stack([j(200.0), j(200.0), j(300.0), j(200.0), j(0.5), j(0.5)], 1)
    if model == F.CAM_RADTAN:          # rad_tan.rs:895-903 (test camera), focal scaled like the other +z models
        return np.stack([j(200.0), j(200.0), j(300.0), j(200.0), j(0.1), j(0.01), j(0.001), j(0.002), j(0.001)], 1)
    if model == F.CAM_UCM:             # ucm.rs:715-717
        return np.stack([j(200.0), j(200.0), j(300.0), j(200.0), j(0.6)], 1)
    if model == F.CAM_EUCM:            # eucm.rs:863-868
        return np.stack([j(200.0), j(200.0), j(300.0), j(200.0), j(0.7), j(1.5)], 1)
    if model == F.CAM_FOV:             # fov.rs:757-759
        return np.stack([j(200.0), j(200.0), j(300.0), j(200.0), j(1.5)], 1)
    if model == F.CAM_FTHETA:          # ftheta.rs:453-455, scaled to the same image size
        return np.stack([j(300.0), j(200.0), j(200.0), j(-4.0), j(0.8), j(-0.04)], 1)
    raise ValueError(model)


def make_problem(ncam, npts, mean_track, camera_model=F.CAM_BAL, loss=(F.LOSS_HUBER, 1.0), seed=0, self_calibration=True,
                 pixel_sigma=0.5, outlier_frac=0.02, fix_first_pose=True, fix_first_intr=False, window_frac=0.04,
                 pose_sigma_t=0.02, pose_sigma_deg=0.5, point_sigma=0.02, intr_rel_sigma=0.01, shuffle_obs=False) -> BAProblem:
    rng = np.random.Generator(np.random.PCG64(seed))
    R, t = _ring_cameras(rng, ncam, camera_model)
    intr = _intrinsics(rng, ncam, camera_model)
    # points: uniform in the unit ball
    v = rng.standard_normal((npts, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    pts = v * np.cbrt(rng.uniform(0, 1, npts))[:, None]
    # tracks: heavy-tailed camera popularity, ring locality
    w = (np.arange(ncam) + 10.0) ** -0.7
    w = w[rng.permutation(ncam)]
    cdf = np.cumsum(w) / w.sum()
    c0 = np.minimum(np.searchsorted(cdf, rng.uniform(0, 1, npts)), ncam - 1)
    order = np.argsort(c0 + 0.35 * ncam * window_frac * rng.standard_normal(npts), kind="stable")
    pts, c0 = pts[order], c0[order]
    extra = np.maximum(mean_track - 2.0, 0.05)
    klen = 2 + rng.geometric(1.0 / (1.0 + extra * 1.12), npts) - 1        # >= 2, mean ~ mean_track (before de-duplication)
    klen = np.minimum(klen, ncam)
    pid = np.repeat(np.arange(npts, dtype=np.int64), klen)
    sigma = max(1.0, window_frac * ncam)
    off = np.rint(sigma * rng.standard_normal(pid.shape[0])).astype(np.int64)
    first = np.concatenate([[0], np.cumsum(klen)[:-1]])
    off[first] = 0
    if ncam > 1:
        off[first + 1] = np.where(off[first + 1] == 0, 1 + rng.integers(0, max(1, int(sigma)), npts), off[first + 1])
    cid = (c0[pid] + off) % ncam
    key = np.unique(pid * ncam + cid)
    obs_pt = (key // ncam).astype(np.uint32)
    obs_cam = (key % ncam).astype(np.uint32)
    # measurements from the ground truth
    pc = np.einsum("nij,nj->ni", R[obs_cam], pts[obs_pt]) + t[obs_cam]
    uv, valid = project(camera_model, intr[obs_cam], pc)
    if not valid.all():
        keep = valid
        obs_pt, obs_cam, uv = obs_pt[keep], obs_cam[keep], uv[keep]
    nobs = obs_pt.shape[0]
    uv = uv + pixel_sigma * rng.standard_normal((nobs, 2))
    out = rng.uniform(0, 1, nobs) < outlier_frac
    uv[out] += rng.uniform(-50, 50, (int(out.sum()), 2))
    # initial values = truth perturbed
    q_true = quat_from_matrix(R)
    dth = np.deg2rad(pose_sigma_deg) * rng.standard_normal((ncam, 3))
    q0 = quat_mul(q_true, axis_angle_to_quat(dth))
    t0 = t + pose_sigma_t * rng.standard_normal((ncam, 3))
    pose0 = np.concatenate([t0, q0], axis=1)
    pts0 = pts + point_sigma * rng.standard_normal((npts, 3))
    if self_calibration:
        scale = np.maximum(np.abs(intr), 1e-3)
        intr0 = intr + intr_rel_sigma * scale * rng.standard_normal(intr.shape)
    else:
        intr0 = intr.copy()
    if fix_first_pose:  # gauge: keep camera 0 at the truth
        pose0[0] = np.concatenate([t[0], q_true[0]])
    if shuffle_obs:
        perm = rng.permutation(nobs)
        obs_pt, obs_cam, uv = obs_pt[perm], obs_cam[perm], uv[perm]
    K = intr.shape[1]
    pose_fixed = np.zeros(ncam, np.uint8)
    intr_fixed = np.zeros(ncam, np.uint16)
    if fix_first_pose:
        pose_fixed[0] = 0x3F      # bin/bundle_adjustment.rs:294-298
    if fix_first_intr:
        intr_fixed[0] = (1 << K) - 1  # benches/bundle_adjustment_benchmark.rs:336-339
    opt = F.OPT_POSE | F.OPT_LANDMARK | (F.OPT_INTRINSIC if self_calibration else 0)
    lp = tuple(loss[1:]) + (0.0,) * (5 - len(loss))
    return BAProblem(camera_model=camera_model, opt_flags=opt, pose=pose0, intr=intr0, pt=pts0, obs_cam=obs_cam, obs_pt=obs_pt, obs_uv=uv,
                     loss_id=loss[0], loss_params=lp[:4], intr_vars_present=True, pose_fixed=pose_fixed, intr_fixed=intr_fixed,
                     meta={"seed": seed, "truth_pose": np.concatenate([t, q_true], 1), "truth_intr": intr, "truth_pt": pts})


def make_calibration_scene(camera_model=F.CAM_KANNALA_BRANDT, ncam=5, grid=(20, 10), spacing=0.1, wall_z=3.0, arc_spread=0.8, seed=100,
                           landmark_sigma=0.01, pose_sigma_t=0.02, pose_sigma_deg=1.0, intr_rel_sigma=0.02, scene="wall") -> BAProblem:
    """The graph of the reference's camera calibration tests (tests/camera_kannala_brandt_integration.rs:45-225,
    camera_double_sphere_integration.rs, tests/camera_test_utils.rs:215-285): a planar wall of grid[0] x grid[1] points at depth
    wall_z, ncam cameras on a horizontal arc looking down +z (identity rotation), every camera sees every point, exact
    projections; ONE ProjectionFactor with all its observations per camera over [pose_k, "landmarks", "intrinsics"] - i.e. one
    intrinsics variable shared by all cameras (APEX_OPT_SHARED_INTRINSICS), SelfCalibration, no loss, pose_0 fixed. Initial
    values = truth perturbed (1 cm, 2 cm / 1 deg, 2 %), own seeded noise."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nx, ny = grid
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny))
    pts = np.stack([ix.ravel() * spacing - (nx - 1) * spacing / 2, iy.ravel() * spacing - (ny - 1) * spacing / 2, np.full(nx * ny, wall_z)], 1)
    if scene == "hemisphere":   # generate_scene_points (tests/camera_test_utils.rs:24-43): index-driven points in front of the cameras
        i = np.arange(nx * ny, dtype=np.float64)
        a1, a2, depth = (i * 2.4) % (2 * np.pi), (i * 1.7) % np.pi, 2.0 + 3.0 * ((i * 0.17) % 1.0)
        pts = np.stack([depth * np.sin(a2) * np.cos(a1), depth * np.sin(a2) * np.sin(a1), depth * np.abs(np.cos(a2)) + 2.0], 1)
    tx = ((np.arange(ncam) / (ncam - 1)) - 0.5) * 2.0 * arc_spread if ncam > 1 else np.zeros(1)
    t = np.stack([tx, np.zeros(ncam), np.zeros(ncam)], 1)
    q_true = np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (ncam, 1))
    truth = {F.CAM_KANNALA_BRANDT: [200.0, 200.0, 300.0, 200.0, 0.5, 0.1, 0.0, 0.0],
             F.CAM_DOUBLE_SPHERE: [200.0, 200.0, 300.0, 200.0, 0.5, 0.5],
             F.CAM_PINHOLE: [500.0, 500.0, 320.0, 240.0]}[camera_model]
    intr_true = np.tile(np.array(truth), (ncam, 1))
    obs_cam = np.repeat(np.arange(ncam, dtype=np.uint32), pts.shape[0])
    obs_pt = np.tile(np.arange(pts.shape[0], dtype=np.uint32), ncam)
    pc = pts[obs_pt] + t[obs_cam]                       # identity rotation: p_cam = p_world + t
    uv, valid = project(camera_model, intr_true[obs_cam], pc)
    assert valid.all(), "every camera must see every calibration point"
    pts0 = pts + landmark_sigma * rng.standard_normal(pts.shape)
    dth = np.deg2rad(pose_sigma_deg) * rng.standard_normal((ncam, 3))
    pose0 = np.concatenate([t + pose_sigma_t * rng.standard_normal((ncam, 3)), quat_mul(q_true, axis_angle_to_quat(dth))], 1)
    pose0[0] = np.concatenate([t[0], q_true[0]])        # the anchor starts at the truth (it is fixed)
    shared0 = np.array(truth) * (1.0 + intr_rel_sigma * rng.standard_normal(len(truth)))
    pose_fixed = np.zeros(ncam, np.uint8)
    pose_fixed[0] = 0x3F
    return BAProblem(camera_model=camera_model, opt_flags=F.OPT_POSE | F.OPT_LANDMARK | F.OPT_INTRINSIC | F.OPT_SHARED_INTRINSICS, pose=pose0,
                     intr=np.tile(shared0, (ncam, 1)), pt=pts0, obs_cam=obs_cam, obs_pt=obs_pt, obs_uv=uv, loss_id=F.LOSS_NONE, loss_params=(0.0,) * 4,
                     intr_vars_present=True, pose_fixed=pose_fixed, intr_fixed=np.zeros(ncam, np.uint16),
                     meta={"truth_intr": np.array(truth), "truth_pose": np.concatenate([t, q_true], 1), "truth_pt": pts})


def make_shape(name: str, scale: float = 1.0, **kw) -> BAProblem:
    """One of the BASELINE.json configs (optionally scaled down for tests); seed = 0xA9E50000 + config#."""
    ncam, npts, mean_track, model, loss, cfg = SHAPES[name]
    if scale != 1.0:
        ncam = max(4, int(round(ncam * scale)))
        npts = max(16, int(round(npts * scale)))
    kw.setdefault("seed", 0xA9E50000 + cfg)
    kw.setdefault("camera_model", model)
    kw.setdefault("loss", loss)
    p = make_problem(ncam, npts, mean_track, **kw)
    p.meta["shape"] = name
    return p
