"""Seeded random test cases shared by the CPU (oracle vs reference sources) and GPU (CUDA path vs oracle) fuzz tests."""
import contextlib
import os
import sys

import numpy as np

from kimera_semantics_b200 import capi
from parity_utils import make_config

C21 = 21


def color_table(cfg):
    pal = np.array([[cfg.label_color[l][k] for k in range(3)] for l in range(C21)], np.uint8)
    return pal, np.arange(C21, dtype=np.uint8)


@contextlib.contextmanager
def quiet_stderr():
    sys.stderr.flush()
    saved, devnull = os.dup(2), os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 2)
    try:
        yield
    finally:
        os.dup2(saved, 2)
        os.close(saved)
        os.close(devnull)


def make_case(seed):
    """-> (ksg_config, [(T_G_C, points_C, rgba, freespace), ...])"""
    rng = np.random.default_rng(seed)
    itype = capi.KSG_INTEGRATOR_FAST if rng.random() < 0.5 else capi.KSG_INTEGRATOR_MERGED
    vs = float(rng.choice([0.05, 0.1, 0.2, 0.13]))
    kw = dict(
        voxels_per_side=int(rng.choice([8, 16])),
        use_const_weight=int(rng.random() < 0.3), use_weight_dropoff=int(rng.random() < 0.7),
        allow_clear=int(rng.random() < 0.7), voxel_carving_enabled=int(rng.random() < 0.8),
        use_sparsity_compensation_factor=int(rng.random() < 0.2), sparsity_compensation_factor=float(rng.choice([1.0, 5.0])),
        enable_anti_grazing=int(rng.random() < 0.4), max_consecutive_ray_collisions=int(rng.integers(0, 4)),
        start_voxel_subsampling_factor=float(rng.choice([1.0, 2.0, 3.0])), integration_order_mode=int(rng.random() < 0.3),
        min_ray_length_m=float(rng.choice([0.1, 0.5])), max_ray_length_m=float(rng.choice([2.0, 5.0])),
        color_mode=int(rng.integers(0, 3)), semantic_measurement_probability=float(rng.choice([0.9, 0.8, 0.6])),
        max_weight=float(rng.choice([1e4, 50.0])), clear_checks_every_n_frames=1,
    )
    n = int(rng.integers(50, 1500))
    cfg = make_config(itype, vs, C21, max_points=4096, **kw)
    cfg.default_truncation_distance = float(np.float32(rng.choice([2.0, 4.0])) * np.float32(vs))
    pal = np.array([[cfg.label_color[l][k] for k in range(4)] for l in range(256)], np.uint8)
    frames = []
    for _ in range(int(rng.integers(1, 4))):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        T = np.concatenate([q, rng.uniform(-3, 3, 3)]).astype(np.float32)
        # a blob in front of the camera, plus far points (clearing / rejected), near points (below the minimum range) and |z| ~ 0
        pts = rng.normal(size=(n, 3)) * rng.choice([0.3, 1.0, 3.0]) + rng.uniform(-1, 1, 3) + np.array([0, 0, rng.uniform(0.5, 4)])
        k = n // 10
        pts[:k] *= 4.0
        pts[k:2 * k] *= 0.05
        pts[2 * k:2 * k + 5, 2] = rng.uniform(-1e-7, 1e-7, 5)
        pts = pts.astype(np.float32)
        pts[np.abs(pts) < 1e-4] = 1e-3       # no exactly axis-aligned rays (0/0 in the RayCaster)
        lab = rng.integers(0, C21, n).astype(np.uint8)
        rgba = np.ascontiguousarray(pal[lab])
        rgba[rng.random(n) < 0.02] = (3, 1, 4, 255)          # colours that are not in the table -> label 0
        frames.append((T, pts, rgba, bool(rng.random() < 0.15)))
    return cfg, frames
