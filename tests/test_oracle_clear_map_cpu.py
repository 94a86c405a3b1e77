"""kso_clear_map (twin of ksg_clear_map): Layer::removeAllBlocks() on a live integrator keeps the fast integrator's two approximate sets -
a scan only re-offsets them (fast.cpp:165-171) - so the next scan into the emptied layers is NOT what a fresh integrator
would produce.  This is the semantics of the frame-per-GPU batch mode.  CPU only."""
import numpy as np

from kimera_semantics_b200.capi import KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED
from oracle.oracle_py import OracleIntegrator
from parity_utils import frames, make_config


def _run(o, fr):
    cam, depth, label, T = fr
    st = o.integrate_depth(T, depth, label, cam.K)
    return st, o.export()


def test_emptied_layers_of_a_live_fast_integrator_differ_from_a_fresh_one():
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.05, 21, max_points=160 * 120)
    fr = list(frames(160, 120, 21, 2))
    live = OracleIntegrator(cfg)
    _run(live, fr[0])
    live.clear_map()
    assert live.num_blocks() == 0
    st_live, e_live = _run(live, fr[1])
    fresh = OracleIntegrator(cfg)
    st_fresh, e_fresh = _run(fresh, fr[1])
    # stale set entries of the previous offset answer "already seen": far fewer rays are cast and far fewer voxels updated
    assert 0 < st_live.voxel_updates < st_fresh.voxel_updates // 2
    assert float(e_live["tsdf_weight"].sum()) < 0.5 * float(e_fresh["tsdf_weight"].sum())
    # the map itself is really empty before the scan: every block of the live result was touched by the second scan alone
    assert len(e_live["block_index"]) <= len(e_fresh["block_index"])
    # a second clear + the same frame again: the sets moved on once more, the result changes again (the state is the integrator's)
    live.clear_map()
    st_again, _ = _run(live, fr[1])
    assert st_again.voxel_updates > 0      # value is data dependent; the call must simply work
    live.close(); fresh.close()


def test_merged_has_no_state_outside_its_layers():
    cfg = make_config(KSG_INTEGRATOR_MERGED, 0.10, 5, max_points=160 * 120)
    fr = list(frames(160, 120, 5, 2))
    live = OracleIntegrator(cfg)
    _run(live, fr[0])
    live.clear_map()
    _, e_live = _run(live, fr[1])
    fresh = OracleIntegrator(cfg)
    _, e_fresh = _run(fresh, fr[1])
    for k in e_fresh:
        assert np.array_equal(e_live[k], e_fresh[k]), k
    live.close(); fresh.close()
