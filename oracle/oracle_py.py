"""ctypes binding of the CPU oracle (oracle/ks_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from kimera_semantics_b200.capi import KsgConfig, KsgFrameStats, export_arrays, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: Dict[str, C.CDLL] = {}


def build(force: bool = False) -> None:
    """Compile the oracle with gcc (seconds)."""
    out = os.path.join(_HERE, "_build", "libks_oracle.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(os.path.join(_HERE, "ks_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])


def load(fast_build: bool = False) -> C.CDLL:
    name = "libks_oracle_fast.so" if fast_build else "libks_oracle.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, "_build", name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    H = C.c_void_p
    fp, u8p, i32p, i64p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    sp = C.POINTER(KsgFrameStats)
    lib.kso_create.argtypes = [C.POINTER(KsgConfig), C.c_int]
    lib.kso_create.restype = H
    lib.kso_destroy.argtypes = [H]
    lib.kso_set_color_to_label.argtypes = [H, u8p, u8p, C.c_int]
    lib.kso_integrate_points.argtypes = [H, fp, fp, u8p, u8p, C.c_int64, C.c_int, sp]
    lib.kso_integrate_points.restype = C.c_int
    lib.kso_integrate_depth.argtypes = [H, fp, fp, u8p, C.c_int, C.c_int, fp, sp]
    lib.kso_integrate_depth.restype = C.c_int
    lib.kso_integrate_depth_k64.argtypes = [H, fp, fp, u8p, C.c_int, C.c_int, C.POINTER(C.c_double), sp]
    lib.kso_integrate_depth_k64.restype = C.c_int
    lib.kso_backproject_k64.argtypes = [fp, C.c_int, C.c_int, C.POINTER(C.c_double), fp, i32p]
    lib.kso_backproject_k64.restype = C.c_int64
    lib.kso_backproject.argtypes = [fp, C.c_int, C.c_int, fp, fp, i32p]
    lib.kso_backproject.restype = C.c_int64
    lib.kso_clear_map.argtypes = [H]
    lib.kso_clear_map.restype = None
    lib.kso_num_blocks.argtypes = [H]
    lib.kso_num_blocks.restype = C.c_int64
    lib.kso_export_blocks.argtypes = [H, C.c_int64, i32p, fp, fp, u8p, u8p, fp, u8p]
    lib.kso_export_blocks.restype = C.c_int
    lib.kso_last_updated_blocks.argtypes = [H, C.c_int64, i32p]
    lib.kso_last_updated_blocks.restype = C.c_int64
    lib.kso_last_integrate_seconds.argtypes = [H]
    lib.kso_last_integrate_seconds.restype = C.c_double
    lib.kso_index_hash.argtypes = [C.c_int64] * 3
    lib.kso_index_hash.restype = C.c_uint64
    lib.kso_mixed_index.argtypes = [C.c_uint64, C.c_uint64]
    lib.kso_mixed_index.restype = C.c_uint64
    lib.kso_transform.argtypes = [fp, fp, fp]
    lib.kso_grid_index.argtypes = [fp, C.c_float, i64p]
    lib.kso_block_and_local.argtypes = [i64p, C.c_int, i32p, i32p]
    lib.kso_raycast.argtypes = [fp, fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, i64p, C.c_int64]
    lib.kso_raycast.restype = C.c_int64
    lib.kso_tsdf_update_sequence.argtypes = [C.POINTER(KsgConfig), fp, fp, fp, u8p, C.c_int, i64p, fp, u8p]
    lib.kso_log_likelihood.argtypes = [C.POINTER(KsgConfig), fp, fp]
    lib.kso_semantic_update_sequence.argtypes = [C.POINTER(KsgConfig), fp, C.c_int, fp, u8p, u8p, u8p]
    lib.kso_approx_set_script.argtypes = [C.POINTER(C.c_uint64), C.c_int, u8p]
    lib.kso_blend.argtypes = [u8p, C.c_float, u8p, C.c_float, u8p]
    lib.kso_rainbow.argtypes = [C.c_double, u8p]
    lib.kso_unordered_map_order.argtypes = [i64p, C.c_int64, i64p, i64p]
    lib.kso_unordered_map_order.restype = C.c_int64
    _LIBS[name] = lib
    return lib


class OracleIntegrator:
    """CPU oracle with the same call surface as kimera_semantics_b200.Integrator."""

    def __init__(self, cfg: KsgConfig, canonical_merged=None, fast_build: bool = False):
        """canonical_merged: None = follow cfg.merged_bundle_order (0 = first-insertion order, 1 = the reference's unordered_map order)."""
        if canonical_merged is None:
            canonical_merged = int(cfg.merged_bundle_order) == 0
        self.lib = load(fast_build)
        self.cfg = cfg
        self.handle = self.lib.kso_create(C.byref(cfg), int(canonical_merged))
        if not self.handle:
            raise ValueError("kso_create rejected the config")

    def close(self):
        if self.handle:
            self.lib.kso_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_color_to_label(self, rgb, labels):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        labels = np.ascontiguousarray(labels, np.uint8)
        self.lib.kso_set_color_to_label(self.handle, _ptr(rgb, C.c_uint8), _ptr(labels, C.c_uint8), len(labels))

    def integrate_points(self, T_G_C, xyz, rgba=None, labels=None, freespace=False) -> KsgFrameStats:
        T = np.ascontiguousarray(T_G_C, np.float32)
        xyz = np.ascontiguousarray(xyz, np.float32)
        rgba = None if rgba is None else np.ascontiguousarray(rgba, np.uint8)
        labels = None if labels is None else np.ascontiguousarray(labels, np.uint8)
        st = KsgFrameStats()
        rc = self.lib.kso_integrate_points(self.handle, _ptr(T, C.c_float), _ptr(xyz, C.c_float), _ptr(rgba, C.c_uint8),
                                           _ptr(labels, C.c_uint8), xyz.shape[0], int(freespace), C.byref(st))
        if rc != 0:
            raise ValueError(f"kso_integrate_points: {rc}")
        return st

    def integrate_depth(self, T_G_C, depth, label, K) -> KsgFrameStats:
        T = np.ascontiguousarray(T_G_C, np.float32)
        depth = np.ascontiguousarray(depth, np.float32)
        label = np.ascontiguousarray(label, np.uint8)
        K = np.ascontiguousarray(K, np.float32)
        st = KsgFrameStats()
        h, w = depth.shape
        rc = self.lib.kso_integrate_depth(self.handle, _ptr(T, C.c_float), _ptr(depth, C.c_float), _ptr(label, C.c_uint8), w, h,
                                          _ptr(K, C.c_float), C.byref(st))
        if rc != 0:
            raise ValueError(f"kso_integrate_depth: {rc}")
        return st

    def integrate_depth_k64(self, T_G_C, depth, label, K64) -> KsgFrameStats:
        T = np.ascontiguousarray(T_G_C, np.float32)
        depth = np.ascontiguousarray(depth, np.float32)
        label = np.ascontiguousarray(label, np.uint8)
        K = np.ascontiguousarray(K64, np.float64)
        st = KsgFrameStats()
        h, w = depth.shape
        rc = self.lib.kso_integrate_depth_k64(self.handle, _ptr(T, C.c_float), _ptr(depth, C.c_float), _ptr(label, C.c_uint8), w, h,
                                              _ptr(K, C.c_double), C.byref(st))
        if rc != 0:
            raise ValueError(f"kso_integrate_depth_k64: {rc}")
        return st

    def last_integrate_seconds(self) -> float:
        return float(self.lib.kso_last_integrate_seconds(self.handle))

    def num_blocks(self) -> int:
        return int(self.lib.kso_num_blocks(self.handle))

    def export(self) -> Dict[str, np.ndarray]:
        return export_arrays(self.lib, self.handle, "kso", self.cfg.voxels_per_side, self.cfg.num_labels)

    def clear_map(self):
        """Remove every block, keep the integrator (its per-scan approximate sets) - Layer::removeAllBlocks on a live integrator."""
        self.lib.kso_clear_map(self.handle)

    def last_updated_blocks(self) -> np.ndarray:
        n = int(self.lib.kso_last_updated_blocks(self.handle, 0, None))
        out = np.zeros((n, 3), np.int32)
        if n:
            self.lib.kso_last_updated_blocks(self.handle, n, _ptr(out, C.c_int32))
        return out
