"""Deterministic synthetic depth + label + pose sequences.

Scene: the reference's simulation world (kimera_semantics_ros/src/semantic_simulation_eval.cpp:16-34:
sphere c=(0,0,2) r=2; plane through (-2,-4,2) n=(0,1,0); plane through (4,0,0) n=(-1,0,0);
cube c=(-4,4,2) side 4; ground z=0.03) enclosed by a 12 x 12 x 5 m box so that every pixel hits.
Rays are traced analytically in float64 and rounded to float32.  Labels: object id -> base label, then
a 0.5 m checkerboard over all C classes; 2 % of the pixels get a uniformly random label; depth noise is
multiplicative N(1, 0.005^2).  Randomness is a counter-based hash (splitmix64 of seed, frame, pixel)
so that any frame can be generated independently and identically on every host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np


@dataclass
class Camera:
    width: int
    height: int
    fx: float
    fy: float
    cx: float
    cy: float

    @property
    def K(self) -> np.ndarray:
        return np.array([self.fx, self.fy, self.cx, self.cy], np.float32)


def make_camera(width: int, height: int) -> Camera:
    # pinhole fx = fy = 525 * W / 640, half-pixel principal point (no exactly-zero ray components, A.7)
    f = 525.0 * width / 640.0
    return Camera(width, height, f, f, (width - 1) / 2.0, (height - 1) / 2.0)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _uniform(seed: int, frame: int, n: int, stream: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = np.uint64((seed * 0x100000001B3 + frame * 0x9E3779B1 + stream * 0x85EBCA77) & 0xFFFFFFFFFFFFFFFF)
        idx = np.arange(n, dtype=np.uint64)
        h = _splitmix64(_splitmix64(idx + base) ^ base)
    return ((h >> np.uint64(11)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def quat_from_rpy(roll: float, pitch: float, yaw: float) -> np.ndarray:
    cr, sr = np.cos(roll / 2), np.sin(roll / 2)
    cp, sp = np.cos(pitch / 2), np.sin(pitch / 2)
    cy, sy = np.cos(yaw / 2), np.sin(yaw / 2)
    return np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                     cr * cp * sy - sr * sp * cy])


def _quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def _quat_rotate(q, v):
    w, x, y, z = q
    qv = np.array([x, y, z])
    uv = np.cross(qv, v)
    uv = uv + uv
    return v + w * uv + np.cross(qv, uv)


def pose(frame: int, radius: float = 3.0, height: float = 1.5, yaw_rate: float = 0.02, pitch: float = 0.1,
         roll: float = 0.03, phase: float = -2.967, look_offset: float = -0.6) -> np.ndarray:
    """T_G_C as (qw qx qy qz tx ty tz) float32.  The camera moves on a 3 m circle about the scene
    origin (outside the r = 2 m sphere, clear of the cube and the two walls for 240 frames), looks
    inward but 0.6 rad off-centre (so the view mixes the near sphere, the ground and far walls beyond
    max_ray_length -> clearing rays), angle advancing 0.02 rad / frame, fixed pitch and roll
    (non axis-aligned).  (SURVEY.md 8d put the camera on a 1 m circle, which is inside the sphere.)"""
    ang = phase + yaw_rate * frame
    yaw = ang + np.pi + look_offset
    # camera optical frame (x right, y down, z forward) from a body frame (x forward, y left, z up)
    q_body = quat_from_rpy(roll, pitch, yaw)
    q_opt = np.array([0.5, -0.5, 0.5, -0.5])  # body <- optical
    q = _quat_mul(q_body, q_opt)
    q = q / np.linalg.norm(q)
    t = np.array([radius * np.cos(ang), radius * np.sin(ang), height])
    return np.concatenate([q, t]).astype(np.float32)


def _trace(origin: np.ndarray, dirs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Nearest hit along unit directions `dirs` (N,3) from `origin`; returns (t, object id)."""
    n = dirs.shape[0]
    best_t = np.full(n, np.inf)
    best_id = np.zeros(n, np.int64)

    def consider(t, oid, mask=None):
        nonlocal best_t, best_id
        ok = np.isfinite(t) & (t > 1e-6) & (t < best_t)
        if mask is not None:
            ok &= mask
        best_t = np.where(ok, t, best_t)
        best_id = np.where(ok, oid, best_id)

    with np.errstate(divide="ignore", invalid="ignore"):
        # sphere c=(0,0,2) r=2
        c = np.array([0.0, 0.0, 2.0])
        oc = origin - c
        b = dirs @ oc
        cc = oc @ oc - 4.0
        disc = b * b - cc
        sq = np.sqrt(np.where(disc >= 0, disc, np.nan))
        t0 = -b - sq
        t1 = -b + sq
        consider(np.where(t0 > 1e-6, t0, t1), 1)
        # plane through (-2,-4,2) normal (0,1,0)  -> y = -4
        consider((-4.0 - origin[1]) / dirs[:, 1], 2)
        # plane through (4,0,0) normal (-1,0,0)   -> x = 4
        consider((4.0 - origin[0]) / dirs[:, 0], 3)
        # ground z = 0.03
        consider((0.03 - origin[2]) / dirs[:, 2], 5)
        # cube c=(-4,4,2) side 4 -> [-6,-2] x [2,6] x [0,4] (slab test)
        lo = np.array([-6.0, 2.0, 0.0])
        hi = np.array([-2.0, 6.0, 4.0])
        ta = (lo - origin) / dirs
        tb = (hi - origin) / dirs
        tmin = np.nanmax(np.minimum(ta, tb), axis=1)
        tmax = np.nanmin(np.maximum(ta, tb), axis=1)
        consider(np.where(tmin > 1e-6, tmin, tmax), 4, tmax >= np.maximum(tmin, 0.0))
        # enclosing box [-6,6] x [-6,6] x [0,5] seen from inside: exit point
        lo = np.array([-6.0, -6.0, 0.0])
        hi = np.array([6.0, 6.0, 5.0])
        ta = (lo - origin) / dirs
        tb = (hi - origin) / dirs
        texit = np.nanmin(np.maximum(ta, tb), axis=1)
        consider(texit, 6)
    return best_t, best_id


def frame(cam: Camera, frame_idx: int, num_labels: int, seed: int = 0, T_G_C: np.ndarray | None = None,
          label_noise: float = 0.02, depth_noise: float = 0.005, invalid_fraction: float = 0.0
          ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns (depth float32 [H,W] metres, label uint8 [H,W], T_G_C float32[7])."""
    if T_G_C is None:
        T_G_C = pose(frame_idx)
    T = T_G_C.astype(np.float64)
    q, t = T[:4], T[4:]
    u = np.arange(cam.width, dtype=np.float64)
    v = np.arange(cam.height, dtype=np.float64)
    uu, vv = np.meshgrid(u, v)
    d_c = np.stack([(uu - cam.cx) / cam.fx, (vv - cam.cy) / cam.fy, np.ones_like(uu)], axis=-1).reshape(-1, 3)
    zscale = 1.0 / np.linalg.norm(d_c, axis=1)        # z-depth = range * zscale
    d_c_unit = d_c * zscale[:, None]
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    d_g = d_c_unit @ R.T
    rng_t, obj = _trace(t, d_g)
    hit = t[None, :] + d_g * rng_t[:, None]
    n = hit.shape[0]
    C = num_labels
    cell = np.floor(hit / 0.5).astype(np.int64)
    lab = 1 + ((cell[:, 0] + cell[:, 1] + cell[:, 2] + obj) % (C - 1))
    r_lab = _uniform(seed, frame_idx, n, 1)
    r_val = _uniform(seed, frame_idx, n, 2)
    lab = np.where(r_lab < label_noise, np.floor(r_val * C).astype(np.int64), lab)
    # Box-Muller from two uniform streams
    u1 = np.maximum(_uniform(seed, frame_idx, n, 3), 1e-12)
    u2 = _uniform(seed, frame_idx, n, 4)
    g = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    depth = rng_t * zscale * (1.0 + depth_noise * g)
    depth = depth.astype(np.float32)
    if invalid_fraction > 0:
        r_inv = _uniform(seed, frame_idx, n, 5)
        depth = np.where(r_inv < invalid_fraction, np.float32(np.nan), depth)
    return depth.reshape(cam.height, cam.width), lab.astype(np.uint8).reshape(cam.height, cam.width), T_G_C.astype(np.float32)


def backproject(depth: np.ndarray, cam: Camera) -> Tuple[np.ndarray, np.ndarray]:
    """float32 restatement of PointCloudFromDepth::convert<float>
    (kimera_semantics_ros/include/kimera_semantics_ros/depth_map_to_pointcloud.h:222-266) followed by the
    finite-point filter of voxblox_ros convertPointcloud. Returns (xyz [n,3] float32, pixel index [n])."""
    h, w = depth.shape
    K = cam.K
    constant_x = np.float32(1.0 / np.float64(K[0]))
    constant_y = np.float32(1.0 / np.float64(K[1]))
    uu = np.arange(w, dtype=np.float32)[None, :] - K[2]
    vv = np.arange(h, dtype=np.float32)[:, None] - K[3]
    d = depth.astype(np.float32)
    x = (uu * d) * constant_x
    y = (vv * d) * constant_y
    xyz = np.stack([x, y, d], axis=-1).reshape(-1, 3).astype(np.float32)
    ok = np.isfinite(d).reshape(-1)
    pix = np.nonzero(ok)[0]
    return xyz[pix], pix
