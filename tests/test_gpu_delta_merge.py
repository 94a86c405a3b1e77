"""Frame-per-GPU batch mode on ONE device: every frame of a batch is integrated into an empty map (a "delta",
what each GPU of the batch does) and the deltas are merged into the base map in frame order with ksg_merge_blocks_device.  The oracle runs
the same schedule - one fresh oracle integrator per frame, merged with the numpy twin of the merge kernel (tests/delta_merge_ref.py) - and
the maps must agree bit for bit."""
import numpy as np
import pytest

from kimera_semantics_b200.capi import Integrator, KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED, KSG_COLOR_MODE_COLOR
from oracle.oracle_py import OracleIntegrator
from parity_utils import assert_parity, compare_maps, frames, make_config
import delta_merge_ref as dm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("itype,color_mode", [(KSG_INTEGRATOR_FAST, 1), (KSG_INTEGRATOR_MERGED, 1), (KSG_INTEGRATOR_FAST, KSG_COLOR_MODE_COLOR)])
def test_batch_of_frames_then_merge_equals_the_same_schedule_on_the_oracle(itype, color_mode):
    W, H, C, G = 320, 240, 21, 4
    cfg = make_config(itype, 0.05, C, max_points=W * H, max_updates=16 << 20, color_mode=color_mode)
    pal = np.array([[cfg.label_color[l][k] if cfg.label_color_known[l] else 0 for k in range(4)] for l in range(256)], np.uint8)
    base = Integrator(cfg)
    ref = dm.empty_map(cfg.voxels_per_side, C)
    for batch in range(2):
        for cam, depth, label, T in frames(W, H, C, G, start=batch * G):
            d = Integrator(cfg)
            d.integrate_depth(T, depth, label, cam.K)
            nb, stride, pool, keys = d.device_map_view()
            base.merge_blocks_device(nb, keys, pool)
            base.sync()
            d.close()
            o = OracleIntegrator(cfg)
            o.integrate_depth(T, depth, label, cam.K)
            ref = dm.merge(ref, o.export(), pal, cfg.max_weight, color_mode)
            o.close()
    rep = compare_maps(base.export(), ref)
    assert_parity(rep)
    assert rep["tsdf_distance_bit_mismatch"] == 0 and rep["tsdf_weight_bit_mismatch"] == 0 and rep["sem_priors_bit_mismatch"] == 0, rep
    base.close()


def test_round_robin_integrators_with_emptied_layers_equal_the_same_schedule_on_the_oracle():
    """What bench.py --sharding frames runs: integrator r lives for the whole run, sees frames r, r + G, ..., and its layers are emptied
    between its frames (ksg_clear_map / Layer::removeAllBlocks) - the fast integrator's approximate sets carry over, which changes which
    rays are cast (the sets are only re-offset per scan, fast.cpp:165-171), so this is NOT the fresh-integrator schedule of the test above."""
    W, H, C, G, rounds = 320, 240, 21, 2, 3
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.05, C, max_points=W * H, max_updates=16 << 20)
    pal = np.array([[cfg.label_color[l][k] if cfg.label_color_known[l] else 0 for k in range(4)] for l in range(256)], np.uint8)
    base = Integrator(cfg)
    gpus = [Integrator(cfg) for _ in range(G)]
    oracles = [OracleIntegrator(cfg) for _ in range(G)]
    ref = dm.empty_map(cfg.voxels_per_side, C)
    fr = list(frames(W, H, C, G * rounds))
    fresh_differs = False
    for k, (cam, depth, label, T) in enumerate(fr):
        r = k % G
        gpus[r].clear_map()
        oracles[r].clear_map()
        gpus[r].integrate_depth(T, depth, label, cam.K)
        oracles[r].integrate_depth(T, depth, label, cam.K)
        delta, want = gpus[r].export(), oracles[r].export()
        rep = compare_maps(delta, want)
        assert_parity(rep)
        assert rep["tsdf_distance_bit_mismatch"] == 0 and rep["tsdf_weight_bit_mismatch"] == 0 and rep["sem_priors_bit_mismatch"] == 0, (k, rep)
        if k >= G:      # a live integrator's second scan differs from a fresh integrator's: the emptied-layers semantics is observable
            f = OracleIntegrator(cfg)
            f.integrate_depth(T, depth, label, cam.K)
            e = f.export()
            fresh_differs |= (len(e["block_index"]) != len(want["block_index"])) or not np.array_equal(e["tsdf_weight"], want["tsdf_weight"])
            f.close()
        nb, stride, pool, keys = gpus[r].device_map_view()
        base.merge_blocks_device(nb, keys, pool)
        base.sync()
        ref = dm.merge(ref, want, pal, cfg.max_weight, 1)
    assert fresh_differs
    rep = compare_maps(base.export(), ref)
    assert_parity(rep)
    assert rep["tsdf_distance_bit_mismatch"] == 0 and rep["tsdf_weight_bit_mismatch"] == 0 and rep["sem_priors_bit_mismatch"] == 0, rep
    for x in gpus + oracles + [base]:
        x.close()


def test_voxel_granular_deltas_merge_to_the_same_map_as_whole_block_deltas():
    """ksg_copy_update_log_device + ksg_merge_voxels_device (what bench.py --sharding frames exchanges for `fast`): per batch, the update
    logs of the G live integrators are stacked into one buffer and merged with ONE call; the result equals merging the same deltas block
    by block (ksg_merge_blocks_device), bit for bit, including the updated() list."""
    import torch
    W, H, C, G, rounds = 320, 240, 21, 3, 2
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.05, C, max_points=W * H, max_updates=16 << 20)
    by_blocks, by_voxels = Integrator(cfg), Integrator(cfg)
    gpus = [Integrator(cfg) for _ in range(G)]
    for g in gpus:
        g.set_update_log(1 << 20)
    fr = list(frames(W, H, C, G * rounds))
    for b in range(rounds):
        sizes = []
        for r in range(G):
            cam, depth, label, T = fr[b * G + r]
            gpus[r].clear_map()
            gpus[r].integrate_depth(T, depth, label, cam.K)
            sizes.append(gpus[r].update_log_size())
        assert min(sizes) > 1000
        stride = max(sizes) + 5
        upd = torch.zeros(G * stride * 32, dtype=torch.uint8, device="cuda")
        pri = torch.zeros(G * stride * C, dtype=torch.float32, device="cuda")
        for r in range(G):
            n = gpus[r].copy_update_log_device(upd[r * stride * 32:].data_ptr(), pri[r * stride * C:].data_ptr(), stride)
            assert n == sizes[r]
        torch.cuda.synchronize()
        by_voxels.merge_voxels_device(sizes, stride, upd.data_ptr(), pri.data_ptr())
        touched_v = by_voxels.last_updated_blocks()
        touched_b = []
        for r in range(G):
            nb, _, pool, keys = gpus[r].device_map_view()
            by_blocks.merge_blocks_device(nb, keys, pool)
            by_blocks.sync()
            touched_b.append(by_blocks.last_updated_blocks())
        want = np.unique(np.concatenate(touched_b), axis=0)
        got = np.unique(touched_v, axis=0)
        assert np.array_equal(got, want)
    a, b = by_voxels.export(), by_blocks.export()
    rep = compare_maps(a, b)
    assert_parity(rep)
    assert rep["tsdf_distance_bit_mismatch"] == 0 and rep["tsdf_weight_bit_mismatch"] == 0 and rep["sem_priors_bit_mismatch"] == 0, rep
    assert np.array_equal(a["tsdf_rgba"], b["tsdf_rgba"]) and np.array_equal(a["sem_label"], b["sem_label"]) and np.array_equal(a["sem_rgba"], b["sem_rgba"])
    for x in gpus + [by_blocks, by_voxels]:
        x.close()
