"""tools/libstdcxx_order.py (the data-parallel closed form of libstdc++'s unordered_map iteration order - the bundle order of
the reference's `merged` integrator, merged.cpp:210-231) against the real container, probed through the oracle library."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from kimera_semantics_b200.capi import _ptr
from oracle import oracle_py

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import libstdcxx_order  # noqa: E402

lib = oracle_py.load()


def probe(keys):
    keys = np.ascontiguousarray(keys, np.int64)
    n = len(keys)
    order, bc = np.zeros(n, np.int64), np.zeros(n, np.int64)
    assert lib.kso_unordered_map_order(_ptr(keys, C.c_int64), n, _ptr(order, C.c_int64), _ptr(bc, C.c_int64)) == n
    return order, bc


def hashes_of(keys):
    return np.array([lib.kso_index_hash(int(x), int(y), int(z)) for x, y, z in keys], dtype=np.uint64)


def distinct_keys(rng, n, span):
    keys = np.unique(rng.integers(-span, span, (3 * n, 3)), axis=0)
    rng.shuffle(keys)
    return keys[:n].astype(np.int64)


@pytest.mark.parametrize("n", [1, 2, 11, 12, 13, 14, 29, 30, 59, 60, 127, 128, 257, 1000, 5000])
def test_closed_form_equals_real_container_order(n):
    rng = np.random.default_rng(n)
    keys = distinct_keys(rng, n, 40)
    order, bc = probe(keys)
    got = libstdcxx_order.iteration_order(hashes_of(keys), bc)
    assert np.array_equal(got, order)


def test_bundle_like_keys_of_a_frame():
    """Voxel keys as bundleRays produces them: a thin surface shell, heavily clustered hashes."""
    from kimera_semantics_b200 import synth
    cam = synth.make_camera(320, 240)
    depth, label, T = synth.frame(cam, 0, 21)
    xyz, _ = synth.backproject(depth, cam)
    vox = np.floor(xyz / np.float32(0.05)).astype(np.int64)
    _, first = np.unique(vox, axis=0, return_index=True)
    keys = vox[np.sort(first)]                      # distinct, first-occurrence order
    order, bc = probe(keys)
    assert len(keys) > 3000 and len(np.unique(bc)) >= 8
    assert np.array_equal(libstdcxx_order.iteration_order(hashes_of(keys), bc), order)
    ph = libstdcxx_order.phases(bc)
    assert sum(e for _, e, _ in ph) < 3.0 * len(keys)            # total keys sorted over all phases


def test_rehash_schedule_depends_only_on_the_size():
    """bucket_count() after the i-th insertion does not depend on the keys: one table per process serves every frame."""
    rng = np.random.default_rng(1)
    _, a = probe(distinct_keys(rng, 3000, 100))
    _, b = probe(distinct_keys(rng, 3000, 12))
    assert np.array_equal(a, b)


def test_product_library_probes_the_same_rehash_schedule_as_the_reference_container():
    """KSG_BUNDLE_ORDER_LIBSTDCXX sizes its phases from ksg_unordered_map_schedule (host only, callable without a GPU)."""
    from kimera_semantics_b200 import capi
    rng = np.random.default_rng(2)
    _, bc = probe(distinct_keys(rng, 20000, 60))
    assert np.array_equal(capi.unordered_map_schedule(20000), bc)
    assert capi.unordered_map_schedule(0).shape == (0,)
