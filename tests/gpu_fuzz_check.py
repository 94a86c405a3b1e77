"""Helper run in its own process by tests/test_gpu_fuzz.py: the CUDA path (C-ABI) against the oracle on the seeded random cases of
tests/fuzz_cases.py.  Prints REPORT {seed: {...mismatch counters...}}.  `merged` runs in the canonical bundle order on both sides
for even seeds and in the reference's libstdc++ order for odd seeds."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main(seeds, make_integrator=None):
    import fuzz_cases
    from kimera_semantics_b200 import capi
    from oracle.oracle_py import OracleIntegrator
    from parity_utils import compare_maps
    if make_integrator is None:
        make_integrator = capi.Integrator
    report = {}
    for seed in seeds:
        cfg, frames = fuzz_cases.make_case(seed)
        cfg.merged_bundle_order = seed & 1
        ora = OracleIntegrator(cfg, canonical_merged=(cfg.merged_bundle_order == 0))
        ora.set_color_to_label(*fuzz_cases.color_table(cfg))
        try:
            gpu = make_integrator(cfg)
            gpu.set_color_to_label(*fuzz_cases.color_table(cfg))
            stats_ok = True
            for T, pts, rgba, freespace in frames:
                sg = gpu.integrate_points(T, pts, rgba=rgba, freespace=freespace)
                so = ora.integrate_points(T, pts, rgba=rgba, freespace=freespace)
                stats_ok &= (sg.voxel_updates, sg.rays_cast, sg.points_valid) == (so.voxel_updates, so.rays_cast, so.points_valid)
            rep = compare_maps(gpu.export(), ora.export())
            gpu.close()
            entry = {k: v for k, v in rep.items() if k.endswith("mismatch") or k == "same_blocks"}
            entry["stats_ok"] = bool(stats_ok)
            # ColorMode::kSemanticProbability paints the TSDF colour from expf(): CUDA and libm may differ by one count
            if cfg.color_mode == capi.KSG_COLOR_MODE_SEMANTIC_PROBABILITY:
                entry.pop("tsdf_rgba_mismatch", None)
        except Exception as e:   # a status code from the C-ABI
            entry = {"error": str(e)[:300]}
        report[str(seed)] = entry
    print("REPORT " + json.dumps(report), flush=True)
    return report


if __name__ == "__main__":
    lo, hi = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 24)
    main(range(lo, hi))
