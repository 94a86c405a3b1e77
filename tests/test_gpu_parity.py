"""GPU parity tests: CUDA path (through the C-ABI, include/ksg.h) vs the CPU oracle on identical seeded
depth + label + pose sequences.  Bar (north_star): block allocation and per-voxel arg-max label bit-exact,
TSDF distance / weight within 1e-5 relative (the implementation is designed to be bit-exact; the report
also counts bit mismatches)."""
import numpy as np
import pytest

from kimera_semantics_b200 import synth
from kimera_semantics_b200.capi import (Integrator, KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED, KSG_COLOR_MODE_COLOR,
                                        KSG_COLOR_MODE_SEMANTIC_PROBABILITY, KSG_ORDER_SORTED)
from oracle.oracle_py import OracleIntegrator
from parity_utils import assert_parity, compare_maps, frames, make_config, stats_equal

pytestmark = pytest.mark.gpu


def run_depth_sequence(cfg, width, height, n_frames, check_each=False, **fkw):
    gpu = Integrator(cfg)
    ora = OracleIntegrator(cfg)
    reps = []
    for i, (cam, depth, label, T) in enumerate(frames(width, height, cfg.num_labels, n_frames, **fkw)):
        sg = gpu.integrate_depth(T, depth, label, cam.K)
        so = ora.integrate_depth(T, depth, label, cam.K)
        ok, why = stats_equal(sg, so)
        assert ok, f"frame {i}: {why}"
        if check_each or i == n_frames - 1:
            rep = compare_maps(gpu.export(), ora.export())
            reps.append(rep)
            assert_parity(rep)
            assert np.array_equal(gpu.last_updated_blocks(), ora.last_updated_blocks()), f"frame {i}: updated() block sets differ"
    gpu.close()
    return reps


# BASELINE.json configs[0]: 320x240, 10 cm, 5 classes, fast
def test_fast_config0_320x240_10cm_c5():
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.10, 5, max_points=320 * 240)
    reps = run_depth_sequence(cfg, 320, 240, 3, check_each=True)
    assert reps[-1]["observed_voxels"] > 1000


# BASELINE.json configs[1] geometry: 640x480, 5 cm, 21 classes, fast; several frames exercise the
# cross-frame persistence of the two approximate hash sets
def test_fast_config1_640x480_5cm_c21_sequence():
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.05, 21)
    run_depth_sequence(cfg, 640, 480, 5, check_each=True)


def test_fast_cooperative_copy_apply_mode():
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.05, 21, apply_mode=1)
    run_depth_sequence(cfg, 640, 480, 2)


def test_merged_320x240_5cm_c21():
    cfg = make_config(KSG_INTEGRATOR_MERGED, 0.05, 21, max_points=320 * 240, max_updates=8 << 20)
    run_depth_sequence(cfg, 320, 240, 2, check_each=True)


# BASELINE.json configs[2] at reduced resolution (full size: test_gpu_fullsize.py): 2 cm, merged
def test_merged_160x120_2cm_c21():
    cfg = make_config(KSG_INTEGRATOR_MERGED, 0.02, 21, max_points=160 * 120, max_updates=16 << 20)
    run_depth_sequence(cfg, 160, 120, 2)


def test_merged_640x480_5cm_c21():
    cfg = make_config(KSG_INTEGRATOR_MERGED, 0.05, 21, max_updates=16 << 20)
    run_depth_sequence(cfg, 640, 480, 2)


def test_merged_anti_grazing():
    cfg = make_config(KSG_INTEGRATOR_MERGED, 0.05, 21, max_points=320 * 240, max_updates=8 << 20, enable_anti_grazing=1)
    run_depth_sequence(cfg, 320, 240, 2)


# ADE20K-size label set (configs[3] class count): exercises the class-group loop of the tile kernel
@pytest.mark.parametrize("itype", [KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED])
def test_c150_class_groups(itype):
    cfg = make_config(itype, 0.10, 150, max_points=320 * 240, max_updates=8 << 20, max_blocks=2048)
    run_depth_sequence(cfg, 320, 240, 2)


@pytest.mark.parametrize("vps", [4, 8, 32])
def test_voxels_per_side(vps):
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.10, 5, vps=vps, max_points=320 * 240, max_blocks=65536 if vps == 4 else 8192)
    run_depth_sequence(cfg, 320, 240, 2)
    cfg = make_config(KSG_INTEGRATOR_MERGED, 0.10, 5, vps=vps, max_points=160 * 120, max_updates=8 << 20,
                      max_blocks=65536 if vps == 4 else 8192)
    run_depth_sequence(cfg, 160, 120, 2)


def test_invalid_depth_pixels_are_dropped():
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.05, 21, max_points=320 * 240)
    run_depth_sequence(cfg, 320, 240, 2, invalid_fraction=0.1)
    cfg = make_config(KSG_INTEGRATOR_MERGED, 0.05, 21, max_points=320 * 240, max_updates=8 << 20)
    run_depth_sequence(cfg, 320, 240, 2, invalid_fraction=0.1)


@pytest.mark.parametrize("mode", [KSG_COLOR_MODE_COLOR, KSG_COLOR_MODE_SEMANTIC_PROBABILITY])
@pytest.mark.parametrize("itype", [KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED])
def test_color_modes(mode, itype):
    cfg = make_config(itype, 0.10, 5, max_points=320 * 240, max_updates=8 << 20, color_mode=mode)
    gpu, ora = Integrator(cfg), OracleIntegrator(cfg)
    for cam, depth, label, T in frames(320, 240, 5, 2):
        gpu.integrate_depth(T, depth, label, cam.K)
        ora.integrate_depth(T, depth, label, cam.K)
    a, b = gpu.export(), ora.export()
    rep = compare_maps(a, b)
    if mode == KSG_COLOR_MODE_SEMANTIC_PROBABILITY:
        # colour = rainbow(exp(prior)): expf differs by <= 2 ulp between libm and CUDA -> allow +-1 per channel
        d = np.abs(a["tsdf_rgba"].astype(np.int32) - b["tsdf_rgba"].astype(np.int32)).max()
        assert d <= 1, rep
        rep["tsdf_rgba_mismatch"] = 0.0
    assert_parity(rep)


def test_points_entry_with_colour_coded_labels():
    """integratePointCloud(T_G_C, points_C, colors): labels arrive encoded as colours (fast.cpp:152-158)."""
    C = 21
    for itype in (KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED):
        cfg = make_config(itype, 0.05, C, max_points=320 * 240, max_updates=8 << 20)
        gpu, ora = Integrator(cfg), OracleIntegrator(cfg)
        pal = np.array([[cfg.label_color[l][k] for k in range(4)] for l in range(C)], np.uint8)
        for obj in (gpu, ora):
            obj.set_color_to_label(pal[:, :3], np.arange(C, dtype=np.uint8))
        for cam, depth, label, T in frames(320, 240, C, 2):
            xyz, pix = synth.backproject(depth, cam)
            rgba = pal[label.reshape(-1)[pix]].copy()
            rgba[::97] = (1, 2, 3, 255)          # unknown colours -> label 0 (color.cpp:80)
            sg = gpu.integrate_points(T, xyz, rgba=rgba)
            so = ora.integrate_points(T, xyz, rgba=rgba)
            ok, why = stats_equal(sg, so)
            assert ok, why
        assert_parity(compare_maps(gpu.export(), ora.export()))


def test_points_entry_explicit_labels_and_freespace():
    C = 5
    for itype in (KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED):
        cfg = make_config(itype, 0.10, C, max_points=320 * 240, max_updates=8 << 20)
        gpu, ora = Integrator(cfg), OracleIntegrator(cfg)
        for i, (cam, depth, label, T) in enumerate(frames(320, 240, C, 2)):
            xyz, pix = synth.backproject(depth, cam)
            lab = label.reshape(-1)[pix]
            sg = gpu.integrate_points(T, xyz, labels=lab, freespace=(i == 1))
            so = ora.integrate_points(T, xyz, labels=lab, freespace=(i == 1))
            ok, why = stats_equal(sg, so)
            assert ok, why
        assert_parity(compare_maps(gpu.export(), ora.export()))


def test_sorted_integration_order():
    for itype in (KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED):
        cfg = make_config(itype, 0.10, 5, max_points=160 * 120, max_updates=8 << 20, integration_order_mode=KSG_ORDER_SORTED)
        run_depth_sequence(cfg, 160, 120, 2)


def test_empty_and_tiny_clouds():
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.10, 5, max_points=1024)
    gpu, ora = Integrator(cfg), OracleIntegrator(cfg)
    T = synth.pose(0)
    for n in (0, 1, 7, 1023, 1024):
        rng = np.random.default_rng(n)
        xyz = (rng.random((n, 3), dtype=np.float32) * 2 + 0.5).astype(np.float32)
        lab = rng.integers(0, 4, n).astype(np.uint8)
        sg = gpu.integrate_points(T, xyz, labels=lab)
        so = ora.integrate_points(T, xyz, labels=lab)
        ok, why = stats_equal(sg, so)
        assert ok, f"n={n}: {why}"
    assert_parity(compare_maps(gpu.export(), ora.export()))


def test_reset_and_reuse():
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.10, 5, max_points=320 * 240)
    gpu = Integrator(cfg)
    for cam, depth, label, T in frames(320, 240, 5, 2):
        gpu.integrate_depth(T, depth, label, cam.K)
    gpu.reset()
    assert gpu.num_blocks() == 0
    ora = OracleIntegrator(cfg)
    for cam, depth, label, T in frames(320, 240, 5, 2):
        gpu.integrate_depth(T, depth, label, cam.K)
        ora.integrate_depth(T, depth, label, cam.K)
    assert_parity(compare_maps(gpu.export(), ora.export()))
