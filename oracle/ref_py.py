"""ctypes binding of oracle/_ref/libks_ref_hybrid.so: the reference's own kimera_semantics sources compiled against
the stand-in dependency headers of oracle/ref_stubs/ (see oracle/ref_hybrid.cpp).  TEST INFRASTRUCTURE ONLY: it pins the
oracle; nothing in the product, smoke() or bench.py loads it.

The library is built by `make -C oracle ref` where /root/reference exists and travels to the GPU box as a prebuilt file.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

from kimera_semantics_b200.capi import KsgConfig, export_arrays, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libks_ref_hybrid.so")
FAST_LIB_PATH = os.path.join(_HERE, "_ref", "libks_ref_hybrid_fast.so")   # -O3 timing build, same results
_LIBS: Dict[str, C.CDLL] = {}


def available(fast_build: bool = False) -> bool:
    return os.path.exists(FAST_LIB_PATH if fast_build else LIB_PATH)


def load(fast_build: bool = False) -> C.CDLL:
    path = FAST_LIB_PATH if fast_build else LIB_PATH
    if path not in _LIBS:
        lib = C.CDLL(path)
        H = C.c_void_p
        fp, u8p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
        lib.kref_num_labels.restype = C.c_int
        lib.kref_create.argtypes = [C.POINTER(KsgConfig)]
        lib.kref_create.restype = H
        lib.kref_destroy.argtypes = [H]
        lib.kref_integrate_points.argtypes = [H, fp, fp, u8p, C.c_int64, C.c_int]
        lib.kref_integrate_points.restype = C.c_int
        lib.kref_num_blocks.argtypes = [H]
        lib.kref_num_blocks.restype = C.c_int64
        lib.kref_num_semantic_blocks.argtypes = [H]
        lib.kref_num_semantic_blocks.restype = C.c_int64
        lib.kref_export_blocks.argtypes = [H, C.c_int64, i32p, fp, fp, u8p, u8p, fp, u8p]
        lib.kref_export_blocks.restype = C.c_int
        lib.kref_update_probabilities.argtypes = [H, fp, fp]
        lib.kref_normalize_probabilities.argtypes = [H, fp]
        lib.kref_label_color.argtypes = [H, C.c_int, u8p]
        lib.kref_log_likelihood.argtypes = [H, fp, fp, fp]
        lib.kref_csv_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_int64]
        lib.kref_csv_dump.restype = C.c_int64
        lib.kref_last_integrate_seconds.argtypes = [H]
        lib.kref_last_integrate_seconds.restype = C.c_double
        _LIBS[path] = lib
    return _LIBS[path]


def csv_dump(path: str) -> str:
    """Both SemanticLabel2Color tables of a CSV file as parsed by the reference's own reader."""
    lib = load()
    n = lib.kref_csv_dump(path.encode(), None, 0)
    buf = C.create_string_buffer(int(n))
    lib.kref_csv_dump(path.encode(), buf, n)
    return buf.raw.decode()


def ros_params(text: str) -> str:
    """The reference's ros_params.cpp (method / csv path / SemanticConfig) evaluated on "key: value" lines; aborts where it aborts."""
    lib = load()
    lib.kref_ros_params.argtypes = [C.c_char_p, C.c_char_p, C.c_int64]
    lib.kref_ros_params.restype = C.c_int64
    n = lib.kref_ros_params(text.encode(), None, 0)
    buf = C.create_string_buffer(int(n))
    lib.kref_ros_params(text.encode(), buf, n)
    return buf.raw.decode()


class RefHybridIntegrator:
    """kimera::FastSemanticTsdfIntegrator / MergedSemanticTsdfIntegrator (the reference's classes) behind the export
    layout of the oracle.  num_labels is the reference's compile-time 21 (common.h:27)."""

    def __init__(self, cfg: KsgConfig, fast_build: bool = False):
        self.lib = load(fast_build)
        self.cfg = cfg
        self.handle = self.lib.kref_create(C.byref(cfg))
        if not self.handle:
            raise ValueError(f"kref_create rejected the config (num_labels must be {self.lib.kref_num_labels()})")

    def close(self):
        if self.handle:
            self.lib.kref_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def integrate_points(self, T_G_C, xyz, rgba=None, freespace: bool = False) -> None:
        T = np.ascontiguousarray(T_G_C, np.float32)
        xyz = np.ascontiguousarray(xyz, np.float32)
        rgba = np.ascontiguousarray(rgba, np.uint8)
        assert rgba.shape == (xyz.shape[0], 4)
        rc = self.lib.kref_integrate_points(self.handle, _ptr(T, C.c_float), _ptr(xyz, C.c_float), _ptr(rgba, C.c_uint8),
                                            xyz.shape[0], int(freespace))
        if rc != 0:
            raise ValueError(f"kref_integrate_points: {rc}")

    # the reference's public per-vector helpers (SemanticIntegratorBase), on plain arrays
    def update_probabilities(self, frequencies, prior) -> np.ndarray:
        f = np.ascontiguousarray(frequencies, np.float32)
        p = np.array(prior, np.float32, copy=True)
        self.lib.kref_update_probabilities(self.handle, _ptr(f, C.c_float), _ptr(p, C.c_float))
        return p

    def normalize_probabilities(self, probs) -> np.ndarray:
        p = np.array(probs, np.float32, copy=True)
        self.lib.kref_normalize_probabilities(self.handle, _ptr(p, C.c_float))
        return p

    def label_color(self, label: int) -> np.ndarray:
        out = np.zeros(4, np.uint8)
        self.lib.kref_label_color(self.handle, int(label), _ptr(out, C.c_uint8))
        return out

    def log_likelihood(self):
        n = self.lib.kref_num_labels()
        m = np.zeros((n, n), np.float32)
        a, b = C.c_float(), C.c_float()
        self.lib.kref_log_likelihood(self.handle, _ptr(m, C.c_float), C.byref(a), C.byref(b))
        return m, a.value, b.value

    def last_integrate_seconds(self) -> float:
        return float(self.lib.kref_last_integrate_seconds(self.handle))

    def num_blocks(self) -> int:
        return int(self.lib.kref_num_blocks(self.handle))

    def num_semantic_blocks(self) -> int:
        return int(self.lib.kref_num_semantic_blocks(self.handle))

    def export(self) -> Dict[str, np.ndarray]:
        return export_arrays(self.lib, self.handle, "kref", self.cfg.voxels_per_side, self.cfg.num_labels)
