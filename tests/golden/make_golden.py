"""Generates tests/golden/oracle_golden.json: per-frame counters and SHA-256 digests of the exported map for a set of
small seeded sequences, produced by the CPU oracle (parity build).  The reference cannot be compiled or imported in this
container, so these fixtures pin the ORACLE (against accidental change) and give the GPU tests a
second, oracle-free anchor.  Re-run only when the oracle's definition changes on purpose:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from kimera_semantics_b200.capi import KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED  # noqa: E402
from oracle.oracle_py import OracleIntegrator  # noqa: E402
from parity_utils import frames, make_config  # noqa: E402

CASES = {
    # name: (integrator, width, height, voxel, classes, frames, extra config)
    "fast_160x120_10cm_c5": (KSG_INTEGRATOR_FAST, 160, 120, 0.10, 5, 3, {}),
    "fast_320x240_5cm_c21": (KSG_INTEGRATOR_FAST, 320, 240, 0.05, 21, 3, {}),
    # first-insertion ("canonical") bundle order, opt-in since round 2
    "merged_160x120_10cm_c5": (KSG_INTEGRATOR_MERGED, 160, 120, 0.10, 5, 2, {"merged_bundle_order": 0}),
    "merged_160x120_5cm_c21": (KSG_INTEGRATOR_MERGED, 160, 120, 0.05, 21, 2, {"merged_bundle_order": 0}),
    "merged_antigrazing_160x120_10cm_c5": (KSG_INTEGRATOR_MERGED, 160, 120, 0.10, 5, 2, {"enable_anti_grazing": 1, "merged_bundle_order": 0}),
    # default = the reference's std::unordered_map bundle order (merged.cpp:210-231)
    "merged_reforder_160x120_10cm_c5": (KSG_INTEGRATOR_MERGED, 160, 120, 0.10, 5, 2, {}),
    "merged_reforder_160x120_5cm_c21": (KSG_INTEGRATOR_MERGED, 160, 120, 0.05, 21, 2, {}),
    "merged_reforder_antigrazing_160x120_10cm_c5": (KSG_INTEGRATOR_MERGED, 160, 120, 0.10, 5, 2, {"enable_anti_grazing": 1}),
}
KEYS = ("block_index", "tsdf_distance", "tsdf_weight", "tsdf_rgba", "sem_label", "sem_priors", "sem_rgba")


def digest(exp):
    return {k: hashlib.sha256(np.ascontiguousarray(exp[k]).tobytes()).hexdigest() for k in KEYS}


def run_case(name, make_integrator):
    itype, w, h, vs, C, nf, kw = CASES[name]
    cfg = make_config(itype, vs, C, max_points=w * h, max_updates=8 << 20, **kw)
    integ = make_integrator(cfg)
    stats = []
    for cam, depth, label, T in frames(w, h, C, nf):
        st = integ.integrate_depth(T, depth, label, cam.K).as_dict()
        stats.append({k: st[k] for k in ("points_in", "points_valid", "rays_cast", "voxel_updates", "blocks_allocated", "blocks_touched")})
    return {"stats": stats, "digest": digest(integ.export())}


if __name__ == "__main__":
    out = {name: run_case(name, OracleIntegrator) for name in CASES}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)
