"""Helper run in its own process by tests/test_gpu_hot_voxels.py: `merged` with ksg_config.hot_voxel_mode = 1 (the parallel pre-pass for
the voxels that receive thousands of updates per frame, csrc/ksg_hot.cuh) against the oracle, bit for bit, on a 2 cm workload
where such voxels exist.  Prints REPORT {...}."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main(make_integrator=None):
    from kimera_semantics_b200 import capi
    from oracle.oracle_py import OracleIntegrator
    from parity_utils import compare_maps, frames, make_config
    if make_integrator is None:
        make_integrator = capi.Integrator
    report = {}
    for order, mode in ((0, 1), (1, 1), (0, 2), (1, 2)):
        cfg = make_config(capi.KSG_INTEGRATOR_MERGED, 0.02, 21, max_points=320 * 240, max_updates=48 << 20, max_blocks=4096)
        cfg.merged_bundle_order = order
        ora = OracleIntegrator(cfg, canonical_merged=(order == 0))
        cfg.hot_voxel_mode = mode
        gpu = make_integrator(cfg)
        hot, fallback, stats_ok = 0, 0, True
        for cam, depth, label, T in frames(320, 240, 21, 3):
            sg = gpu.integrate_depth(T, depth, label, cam.K)
            so = ora.integrate_depth(T, depth, label, cam.K)
            stats_ok &= (sg.voxel_updates, sg.rays_cast) == (so.voxel_updates, so.rays_cast)
            hot += int(getattr(sg, "hot_voxels", 0))
            fallback = int(getattr(sg, "hot_fallback_chunks", 0))
        rep = compare_maps(gpu.export(), ora.export())
        entry = {k: v for k, v in rep.items() if k.endswith("mismatch") or k == "same_blocks"}
        entry.update(stats_ok=bool(stats_ok), hot_voxels=hot, hot_fallback_chunks=fallback)
        report[("libstdcxx" if order else "canonical") + f"/mode{mode}"] = entry
        gpu.close()
    print("REPORT " + json.dumps(report), flush=True)
    return report


if __name__ == "__main__":
    main()
