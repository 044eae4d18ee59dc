"""Deterministic synthetic depth / LiDAR scans for the five BASELINE.json configs.

The reference ships analytic worlds in voxblox/simulation/ (simulation_world.cc:49-117,
objects.h:20-404) and its tests render scans from them
(test/test_sdf_integrators.cc:28-84).  These generators follow the same idea --
closed-form ray/primitive intersection, seeded, float32 output -- but are written
for the scan shapes BASELINE.json names: pinhole 64x48 and
640x480 cameras with an optical (z-forward) frame so that the 1/z^2 point weight
(tsdf_integrator.cc:231-240) is meaningful, and a 2048x128 spinning LiDAR.

A scan is (points_C [N,3] f32, colors [N,4] u8, q_wxyz [4] f32, t [3] f32) with
T_G_C = (q, t): exactly the arguments of TsdfIntegratorBase::integratePointCloud
(tsdf_integrator.h:100-103).  Everything is numpy on the host; nothing here runs
on the timed path except as the producer of input buffers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

Scan = Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]


# --------------------------------------------------------------------------- poses
def quat_from_rpy(roll: float, pitch: float, yaw: float) -> np.ndarray:
    cr, sr = math.cos(roll / 2), math.sin(roll / 2)
    cp, sp = math.cos(pitch / 2), math.sin(pitch / 2)
    cy, sy = math.cos(yaw / 2), math.sin(yaw / 2)
    return np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], dtype=np.float64)


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def quat_to_matrix(q: np.ndarray) -> np.ndarray:
    w, x, y, z = (float(v) for v in q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def quat_from_matrix(R: np.ndarray) -> np.ndarray:
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    q = np.array(q)
    return q / np.linalg.norm(q)


def look_at(position: Sequence[float], target: Sequence[float],
            up: Sequence[float] = (0.0, 0.0, 1.0)) -> np.ndarray:
    """Optical-frame rotation (x right, y down, z forward) looking from position at target."""
    p = np.asarray(position, dtype=np.float64)
    f = np.asarray(target, dtype=np.float64) - p
    f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, dtype=np.float64))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    return quat_from_matrix(np.stack([r, d, f], axis=1))


# ---------------------------------------------------------------------- primitives
@dataclass
class Plane:
    normal: Tuple[float, float, float]
    offset: float  # n . x = offset
    color: Tuple[int, int, int] = (200, 200, 200)

    def hit(self, o, d):
        n = np.asarray(self.normal, dtype=np.float64)
        denom = d @ n
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (self.offset - o @ n) / denom
        t[~np.isfinite(t) | (t <= 1e-9)] = np.inf
        return t


@dataclass
class Sphere:
    center: Tuple[float, float, float]
    radius: float
    color: Tuple[int, int, int] = (230, 80, 60)

    def hit(self, o, d):
        # both roots: the near one from outside, the far one from inside.  (The
        # reference's Sphere::getRayIntersection returns only the near root,
        # simulation/objects.h:65-97, so a camera inside its sphere sees nothing.)
        oc = o - np.asarray(self.center, dtype=np.float64)
        a = np.einsum("ij,ij->i", d, d)
        b = 2.0 * (d @ oc)
        c = oc @ oc - self.radius ** 2
        disc = b * b - 4 * a * c
        ok = disc >= 0
        sq = np.sqrt(np.where(ok, disc, 0.0))
        t0 = (-b - sq) / (2 * a)
        t1 = (-b + sq) / (2 * a)
        t = np.where(t0 > 1e-9, t0, t1)
        t[~ok | (t <= 1e-9)] = np.inf
        return t


@dataclass
class Box:
    lo: Tuple[float, float, float]
    hi: Tuple[float, float, float]
    color: Tuple[int, int, int] = (60, 120, 220)

    def hit(self, o, d):
        lo = np.asarray(self.lo, dtype=np.float64)
        hi = np.asarray(self.hi, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
            ta = (lo - o) * inv
            tb = (hi - o) * inv
        tmin = np.nanmax(np.minimum(ta, tb), axis=1)
        tmax = np.nanmin(np.maximum(ta, tb), axis=1)
        ok = tmax >= np.maximum(tmin, 0.0)
        t = np.where(tmin > 1e-9, tmin, tmax)  # outside: entry; inside (a room): exit
        t[~ok | (t <= 1e-9)] = np.inf
        return t


@dataclass
class CylinderZ:
    center: Tuple[float, float]  # x, y
    radius: float
    zmin: float
    zmax: float
    color: Tuple[int, int, int] = (90, 200, 90)

    def hit(self, o, d):
        ox = o[0] - self.center[0]
        oy = o[1] - self.center[1]
        a = d[:, 0] ** 2 + d[:, 1] ** 2
        b = 2 * (d[:, 0] * ox + d[:, 1] * oy)
        c = ox * ox + oy * oy - self.radius ** 2
        disc = b * b - 4 * a * c
        ok = (disc >= 0) & (a > 1e-18)
        sq = np.sqrt(np.where(ok, disc, 0.0))
        with np.errstate(divide="ignore", invalid="ignore"):
            t0 = (-b - sq) / (2 * a)
        z0 = o[2] + t0 * d[:, 2]
        side = ok & (t0 > 1e-9) & (z0 >= self.zmin) & (z0 <= self.zmax)
        t = np.where(side, t0, np.inf)
        # caps
        for zc in (self.zmin, self.zmax):
            with np.errstate(divide="ignore", invalid="ignore"):
                tc = (zc - o[2]) / d[:, 2]
            xc = ox + tc * d[:, 0]
            yc = oy + tc * d[:, 1]
            capok = np.isfinite(tc) & (tc > 1e-9) & (xc * xc + yc * yc <= self.radius ** 2)
            t = np.minimum(t, np.where(capok, tc, np.inf))
        return t


def trace(prims: Sequence, origin: np.ndarray, dirs_G: np.ndarray):
    """Nearest hit per ray: (t [N] f64 in units of |dir|, prim id [N])."""
    best = np.full(dirs_G.shape[0], np.inf)
    who = np.full(dirs_G.shape[0], -1, dtype=np.int64)
    for i, p in enumerate(prims):
        t = p.hit(origin, dirs_G)
        better = t < best
        best = np.where(better, t, best)
        who[better] = i
    return best, who


# ------------------------------------------------------------------------- sensors
def pinhole_dirs(width: int, height: int, fx: float, fy: float, cx: float, cy: float) -> np.ndarray:
    """Optical-frame ray directions with z = 1, row-major (v outer, u inner)."""
    u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    return np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1).reshape(-1, 3)


def lidar_dirs(n_azimuth: int, n_rings: int, elev_deg: float) -> np.ndarray:
    """Unit directions, sensor frame x forward / z up, ring-major (ring outer)."""
    az = (np.arange(n_azimuth, dtype=np.float64) + 0.25) * (2 * math.pi / n_azimuth)
    el = np.deg2rad(np.linspace(-elev_deg, elev_deg, n_rings) + 0.013)
    azg, elg = np.meshgrid(az, el)
    return np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)],
                    axis=-1).reshape(-1, 3)


def render(prims: Sequence, dirs_C: np.ndarray, q_wxyz: np.ndarray, t: np.ndarray,
           min_range: float = 0.0, max_range: float = math.inf, drop_fraction: float = 0.0,
           rng: Optional[np.random.Generator] = None, depth_noise_sigma: float = 0.0,
           keep_misses_as: Optional[float] = None) -> Scan:
    """Cast dirs_C from pose (q, t) into the scene; returns the scan in the sensor frame.

    Rays that miss, fall outside [min_range, max_range] (Euclidean) or are dropped
    (drop_fraction, seeded) are removed -- like the NaN filter of the ROS front end
    (voxblox_ros/include/voxblox_ros/conversions.h:135-137) -- unless keep_misses_as
    gives a range at which to report them (free-space / clearing returns).
    """
    R = quat_to_matrix(q_wxyz)
    dirs_G = dirs_C @ R.T
    tt, who = trace(prims, np.asarray(t, dtype=np.float64), dirs_G)
    if depth_noise_sigma > 0.0:
        assert rng is not None
        tt = tt + rng.normal(0.0, depth_noise_sigma, size=tt.shape) / np.linalg.norm(dirs_C, axis=1)
    rng_len = tt * np.linalg.norm(dirs_C, axis=1)
    keep = np.isfinite(tt) & (rng_len >= min_range) & (rng_len <= max_range)
    if keep_misses_as is not None:
        miss = ~keep
        tt = np.where(miss, keep_misses_as / np.linalg.norm(dirs_C, axis=1), tt)
        who = np.where(miss, -1, who)
        keep = np.ones_like(keep)
    if drop_fraction > 0.0:
        assert rng is not None
        keep &= rng.random(tt.shape[0]) >= drop_fraction
    pts = (dirs_C * tt[:, None])[keep].astype(np.float32)
    who = who[keep]
    # colour: primitive base colour modulated by a deterministic world-position pattern
    hit_G = (np.asarray(t, dtype=np.float64) + dirs_G[keep] * tt[keep, None])
    base = np.array([p.color for p in prims] + [(0, 0, 0)], dtype=np.float64)[who]
    mod = 0.75 + 0.25 * np.sin(hit_G[:, 0:1] * 3.1 + hit_G[:, 1:2] * 2.3 + hit_G[:, 2:3] * 1.7)
    rgb = np.clip(base * mod, 0, 255).astype(np.uint8)
    colors = np.concatenate([rgb, np.full((rgb.shape[0], 1), 255, dtype=np.uint8)], axis=1)
    return (np.ascontiguousarray(pts), np.ascontiguousarray(colors),
            np.asarray(q_wxyz, dtype=np.float32), np.asarray(t, dtype=np.float32))


# -------------------------------------------------------------------- the configs
def c1_planar_wall() -> Scan:
    """C1: one 64x48 scan of a planar wall at z = 3 m, pose slightly off-axis."""
    dirs = pinhole_dirs(64, 48, 52.5, 52.5, 32.0, 24.0)
    q = quat_from_rpy(0.013, -0.021, 0.017)
    t = np.array([0.013, 0.021, 0.017])
    pts, _, qf, tf = render([Plane((0.0, 0.0, 1.0), 3.0)], dirs, q, t)
    u, v = np.meshgrid(np.arange(64), np.arange(48))
    u, v = u.reshape(-1), v.reshape(-1)
    colors = np.stack([(u * 4) & 255, (v * 5) & 255, (u + v) & 255, np.full_like(u, 255)],
                      axis=1).astype(np.uint8)
    return pts, np.ascontiguousarray(colors), qf, tf


def _orbit_pose(i: int, n: int, radius: float, height: float, target, rng, jitter: float):
    ang = 2 * math.pi * i / n
    pos = np.array([radius * math.cos(ang), radius * math.sin(ang), height])
    pos = pos + rng.normal(0.0, jitter, size=3)
    tgt = np.asarray(target, dtype=np.float64) + rng.normal(0.0, jitter, size=3)
    return look_at(pos, tgt), pos


def c2_sphere_scan(i: int, n_scans: int = 300, width: int = 640, height: int = 480,
                   seed: int = 0) -> Scan:
    """C2, scan i: camera orbiting inside a sphere of radius 3 m (far-root hits)."""
    rng = np.random.default_rng([seed, 2, i])
    f = 525.0 * width / 640.0
    dirs = pinhole_dirs(width, height, f, f, width / 2.0, height / 2.0)
    prims = [Sphere((0.0, 0.0, 0.0), 3.0, (180, 160, 140))]
    ang = 2 * math.pi * i / n_scans
    q, pos = _orbit_pose(i, n_scans, 0.9, 0.1,
                         (2.5 * math.cos(ang + 0.6), 2.5 * math.sin(ang + 0.6), 0.2), rng, 0.01)
    return render(prims, dirs, q, pos)


def c2_sphere_room(n_scans: int = 300, width: int = 640, height: int = 480, seed: int = 0,
                   total: int = 300) -> List[Scan]:
    return [c2_sphere_scan(i, total, width, height, seed) for i in range(n_scans)]


def room_prims() -> list:
    """Cow-and-Lady-shaped scene: a 4 x 4 x 3 m box room with a few analytic objects."""
    return [Box((-2.0, -2.0, 0.0), (2.0, 2.0, 3.0), (170, 170, 160)),
            Sphere((0.6, 0.4, 0.55), 0.55, (210, 90, 70)),
            CylinderZ((-0.7, 0.5), 0.3, 0.0, 1.4, (80, 190, 110)),
            Box((-0.2, -1.1, 0.0), (0.5, -0.5, 0.8), (70, 110, 220)),
            Sphere((-0.9, -0.9, 1.6), 0.35, (220, 200, 60))]


def c3_room_scan(i: int, width: int = 640, height: int = 480, seed: int = 0,
                 drop_fraction: float = 0.15, depth_noise_sigma: float = 0.0) -> Scan:
    """C3/C4, scan i of a handheld-style trajectory (30 Hz) through the room: depth
    0.5-4.5 m, a seeded fraction of pixels dropped.  Scans are independent of each other
    (per-scan RNG stream), so a sequence can be generated in parallel."""
    rng = np.random.default_rng([seed, 3, i])
    f = 525.0 * width / 640.0
    dirs = pinhole_dirs(width, height, f, f, width / 2.0 - 0.5, height / 2.0 - 0.5)
    s = i / 30.0
    pos = np.array([1.25 * math.cos(0.35 * s) + 0.1 * math.sin(1.3 * s),
                    1.25 * math.sin(0.35 * s) + 0.1 * math.cos(1.1 * s),
                    1.35 + 0.2 * math.sin(0.7 * s)]) + rng.normal(0.0, 0.004, size=3)
    tgt = np.array([0.25 * math.sin(0.21 * s), 0.25 * math.cos(0.17 * s),
                    0.7 + 0.25 * math.sin(0.5 * s)])
    return render(room_prims(), dirs, look_at(pos, tgt), pos, min_range=0.5, max_range=4.5,
                  drop_fraction=drop_fraction, rng=rng, depth_noise_sigma=depth_noise_sigma)


def c3_room_sequence(n_scans: int = 1000, width: int = 640, height: int = 480, seed: int = 0,
                     drop_fraction: float = 0.15, depth_noise_sigma: float = 0.0,
                     start: int = 0) -> List[Scan]:
    return [c3_room_scan(start + i, width, height, seed, drop_fraction, depth_noise_sigma)
            for i in range(n_scans)]


def lidar_prims() -> list:
    return [Box((-4.5, -4.5, 0.0), (4.5, 4.5, 4.0), (160, 160, 170)),
            CylinderZ((2.0, 1.5), 0.35, 0.0, 4.0, (200, 120, 60)),
            CylinderZ((-2.2, 2.4), 0.35, 0.0, 4.0, (60, 200, 120)),
            CylinderZ((-1.8, -2.6), 0.35, 0.0, 4.0, (120, 60, 200)),
            CylinderZ((2.6, -1.9), 0.35, 0.0, 4.0, (200, 200, 60))]


def c5_lidar_scan(i: int, n_azimuth: int = 2048, n_rings: int = 128, seed: int = 0) -> Scan:
    """C5, scan i: spinning LiDAR (360 x +-22.5 deg) inside a 9 x 9 x 4 m hall with pillars."""
    rng = np.random.default_rng([seed, 5, i])
    dirs = lidar_dirs(n_azimuth, n_rings, 22.5)
    pos = np.array([0.4 * math.cos(0.4 * i) + 0.05, 0.4 * math.sin(0.4 * i) - 0.03,
                    1.8 + 0.02 * math.sin(i)]) + rng.normal(0.0, 0.003, size=3)
    q = quat_from_rpy(0.011 + 0.002 * i, -0.007, 0.13 * i + 0.017)
    return render(lidar_prims(), dirs, q, pos, min_range=0.3, max_range=9.5)


def c5_lidar_sequence(n_scans: int = 16, n_azimuth: int = 2048, n_rings: int = 128,
                      seed: int = 0) -> List[Scan]:
    return [c5_lidar_scan(i, n_azimuth, n_rings, seed) for i in range(n_scans)]


def _call(args):
    fn, a, kw = args
    return fn(*a, **kw)


def generate_parallel(fn, indices: Sequence[int], workers: Optional[int] = None, **kw) -> List[Scan]:
    """[fn(i, **kw) for i in indices] on a process pool (scan generation is host-side
    numpy and embarrassingly parallel; it is never inside a timed region)."""
    import multiprocessing as mp
    import os

    indices = list(indices)
    workers = max(1, min(workers or (os.cpu_count() or 1), len(indices), 32))
    if workers == 1:
        return [fn(i, **kw) for i in indices]
    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(_call, [(fn, (i,), kw) for i in indices], chunksize=1)
