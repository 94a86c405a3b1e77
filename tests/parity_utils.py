"""Shared helpers for the parity tests: run the same frames through the CUDA path (C-ABI) and the CPU
oracle and compare the exported maps.  The oracle is the checker, never the thing under test."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from kimera_semantics_b200 import synth
from kimera_semantics_b200.capi import KsgConfig, default_config


def make_config(integrator_type: int, voxel_size: float, num_labels: int, vps: int = 16, max_points: int = 640 * 480,
                **kw) -> KsgConfig:
    cfg = default_config(integrator_type, voxel_size, vps, num_labels)
    cfg.dynamic_label[num_labels - 1] = 1      # mirrors dynamic_semantic_labels [20] at C = 21
    cfg.max_points = max_points
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def frames(width: int, height: int, num_labels: int, n: int, start: int = 0, seed: int = 0, **kw):
    cam = synth.make_camera(width, height)
    for f in range(start, start + n):
        depth, label, T = synth.frame(cam, f, num_labels, seed=seed, **kw)
        yield cam, depth, label, T


def compare_maps(a: Dict[str, np.ndarray], b: Dict[str, np.ndarray], rtol: float = 1e-5) -> Dict[str, float]:
    """a = CUDA path, b = oracle. Block allocation and labels must be bit-exact; TSDF distance / weight within
    1e-5 relative (north_star)."""
    rep: Dict[str, float] = {}
    rep["blocks_a"] = len(a["block_index"])
    rep["blocks_b"] = len(b["block_index"])
    same_blocks = a["block_index"].shape == b["block_index"].shape and np.array_equal(a["block_index"], b["block_index"])
    rep["same_blocks"] = float(same_blocks)
    if not same_blocks:
        return rep
    rep["label_mismatch"] = float((a["sem_label"] != b["sem_label"]).sum())
    rep["sem_rgba_mismatch"] = float((a["sem_rgba"] != b["sem_rgba"]).any(axis=-1).sum())
    rep["tsdf_rgba_mismatch"] = float((a["tsdf_rgba"] != b["tsdf_rgba"]).any(axis=-1).sum())
    for k in ("tsdf_distance", "tsdf_weight", "sem_priors"):
        x, y = a[k].astype(np.float64), b[k].astype(np.float64)
        denom = np.maximum(np.abs(y), 1e-12)
        rel = np.abs(x - y) / denom
        rel[(x == y)] = 0.0
        rep[k + "_max_rel"] = float(rel.max()) if rel.size else 0.0
        rep[k + "_bit_mismatch"] = float((a[k].view(np.uint32) != b[k].view(np.uint32)).sum())
    rep["observed_voxels"] = float((b["tsdf_weight"] > 0).sum())
    return rep


def assert_parity(rep: Dict[str, float], rtol: float = 1e-5):
    assert rep["same_blocks"] == 1.0, f"block allocation differs: {rep}"
    assert rep["label_mismatch"] == 0, f"labels differ: {rep}"
    assert rep["sem_rgba_mismatch"] == 0, f"semantic colours differ: {rep}"
    assert rep["tsdf_rgba_mismatch"] == 0, f"tsdf colours differ: {rep}"
    assert rep["tsdf_distance_max_rel"] <= rtol, f"distance differs: {rep}"
    assert rep["tsdf_weight_max_rel"] <= rtol, f"weight differs: {rep}"
    assert rep["sem_priors_max_rel"] <= rtol, f"log-probabilities differ: {rep}"


STAT_KEYS = ("points_in", "points_valid", "rays_cast", "voxel_updates", "blocks_allocated", "blocks_touched")


def stats_equal(sa, sb) -> Tuple[bool, str]:
    da, db = sa.as_dict(), sb.as_dict()
    bad = [f"{k}: gpu {da[k]} oracle {db[k]}" for k in STAT_KEYS if da[k] != db[k]]
    return (not bad), "; ".join(bad)
