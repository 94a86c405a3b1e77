"""NEXT-4: semantic mesh extraction on the device (ksg_extract_mesh, csrc/ksg_mesh.cuh) equals its numpy twin
(tests/mesh_ref.py) bit for bit - vertex positions, TsdfVoxel.color and semantic label per vertex, per-block ranges."""
import numpy as np
import pytest

from kimera_semantics_b200.capi import Integrator, KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED
from parity_utils import frames, make_config
import mesh_ref as mr

pytestmark = pytest.mark.gpu


def _same(got, want):
    assert np.array_equal(got["block_first"], want["block_first"])
    assert np.array_equal(got["vertices"].view(np.uint32), want["vertices"].view(np.uint32))
    assert np.array_equal(got["rgba"], want["rgba"])
    assert np.array_equal(got["labels"], want["labels"])


@pytest.mark.parametrize("itype", [KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED])
def test_mesh_of_the_integrated_scene(itype):
    W, H, C, vs = 320, 240, 21, 0.05
    cfg = make_config(itype, vs, C, max_points=W * H, max_updates=16 << 20)
    gpu = Integrator(cfg)
    for cam, depth, label, T in frames(W, H, C, 6):
        gpu.integrate_depth(T, depth, label, cam.K)
    got = gpu.extract_mesh()
    exp = gpu.export()
    assert np.array_equal(got["block_index"], exp["block_index"])
    want = mr.extract(exp, vs, cfg.voxels_per_side)
    _same(got, want)
    n = len(got["vertices"])
    assert n > 3000 and n % 3 == 0
    # the vertex colour is the colour the semantic integrator handed to the TSDF voxel = the colour of the voxel's label
    # (semantic_integrator_base.cpp:172-191): where a label is set, the two agree
    labelled = got["labels"] > 0
    assert labelled.mean() > 0.9
    lut = {}
    for b in range(len(exp["block_index"])):
        for lab in np.unique(exp["sem_label"][b]):
            if lab and lab not in lut:
                k = np.nonzero(exp["sem_label"][b] == lab)[0][0]
                lut[int(lab)] = exp["sem_rgba"][b].reshape(-1, 4)[k]
    for lab, col in lut.items():
        m = got["labels"] == lab
        if m.any():
            assert np.all(got["rgba"][m] == col[None, :]), lab
    # triangles of the sphere (centre (0, 0, 2), radius 2) face the free space in front of it
    tri = got["vertices"].astype(np.float64).reshape(-1, 3, 3)
    cen = tri.mean(axis=1)
    near = np.abs(np.linalg.norm(cen - np.array([0.0, 0.0, 2.0]), axis=1) - 2.0) < 2 * vs
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    out = ((cen - np.array([0.0, 0.0, 2.0])) * nrm).sum(axis=1)
    assert near.sum() > 100 and (out[near] > 0).mean() > 0.9
    gpu.close()


def test_mesh_of_an_imported_analytic_field_and_the_capacity_contract():
    vs, vps, C = 0.1, 16, 8
    cfg = make_config(KSG_INTEGRATOR_FAST, vs, C, max_points=1024, max_updates=1 << 16)
    exp = mr.sdf_export(lambda x, y, z: np.sqrt(x * x + y * y + z * z) - 1.27, vs, vps, -1, 1)
    V = vps ** 3
    exp["sem_priors"] = np.zeros((len(exp["block_index"]), V, C), np.float32)
    exp["sem_rgba"] = np.zeros((len(exp["block_index"]), V, 4), np.uint8)
    gpu = Integrator(cfg)
    gpu.import_blocks(exp)
    got = gpu.extract_mesh()
    back = gpu.export()
    _same(got, mr.extract(back, vs, vps))
    assert len(got["vertices"]) > 1000
    rad = np.linalg.norm(got["vertices"].astype(np.float64), axis=1)
    assert np.abs(rad - 1.27).max() < 0.15 * vs
    # min_weight above the stored weights: nothing is observed
    assert len(gpu.extract_mesh(min_weight=2.0)["vertices"]) == 0
    # a vertex buffer that is too small is refused and the need is reported
    import ctypes as Ct
    nv, nb = Ct.c_int64(), Ct.c_int64()
    small = np.zeros((10, 3), np.float32)
    rc = gpu.lib.ksg_extract_mesh(gpu.handle, 1e-4, 10, small.ctypes.data_as(Ct.c_void_p), None, None, 0, None, None, Ct.byref(nv), Ct.byref(nb))
    assert rc != 0 and nv.value == len(got["vertices"]) and nb.value == 8
    gpu.close()
