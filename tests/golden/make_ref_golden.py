"""Generates tests/golden/ref_hybrid_golden.json from the REFERENCE'S OWN kimera_semantics sources
(semantic_tsdf_integrator_{fast,merged}.cpp, semantic_integrator_base.cpp, color.cpp, csv_iterator.cpp), compiled in the
build container by `make -C oracle ref` against the stand-in Eigen/glog/voxblox headers of oracle/ref_stubs/ (see
oracle/ref_hybrid.cpp for exactly which half is the real reference).  The fixtures are SHA-256 digests of the exported map
after small seeded sequences pushed through the reference boundary integratePointCloud(T_G_C, points_C, colors, freespace)
(fast.cpp:145-149, merged.cpp:65-69), at the reference's compile-time 21 labels.

They travel to the GPU box (where neither /root/reference nor, necessarily, the hybrid library exists) and anchor
  * the oracle            (tests/test_oracle_vs_ref_hybrid.py, CPU)  - every digest, bit for bit;
  * the CUDA path          (tests/test_gpu_ref_golden.py, GPU)        - `fast` and `merged`: every digest, bit for bit (the product's
    default bundle order for `merged` is the reference's libstdc++ hash-map order).

Regenerate (only in a container that has /root/reference):   make -C oracle ref && python tests/golden/make_ref_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from kimera_semantics_b200 import capi, synth  # noqa: E402
from kimera_semantics_b200.capi import KSG_INTEGRATOR_FAST as FAST, KSG_INTEGRATOR_MERGED as MERGED  # noqa: E402
from parity_utils import frames, make_config  # noqa: E402

C21 = 21
# name: (integrator, width, height, voxel size, frames, config overrides, scenario)
#   scenario: freespace -> the cloud is integrated as freespace points; unknown_colors -> every 97th point carries a colour
#   that is not in the label table (color.cpp:75-80 maps it to label 0)
CASES = {
    "fast_default_3f": (FAST, 160, 120, 0.10, 3, {}, {}),
    "fast_5cm_2f": (FAST, 160, 120, 0.05, 2, {}, {}),
    "fast_vps8": (FAST, 128, 96, 0.10, 2, {"voxels_per_side": 8}, {}),
    "fast_sorted_order": (FAST, 128, 96, 0.10, 2, {"integration_order_mode": capi.KSG_ORDER_SORTED}, {}),
    "fast_color_mode_color": (FAST, 128, 96, 0.10, 2, {"color_mode": capi.KSG_COLOR_MODE_COLOR}, {}),
    "fast_color_mode_probability": (FAST, 128, 96, 0.10, 2, {"color_mode": capi.KSG_COLOR_MODE_SEMANTIC_PROBABILITY}, {}),
    "fast_const_weight_no_dropoff": (FAST, 128, 96, 0.10, 2, {"use_const_weight": 1, "use_weight_dropoff": 0}, {}),
    "fast_sparsity_compensation": (FAST, 128, 96, 0.10, 2, {"use_sparsity_compensation_factor": 1, "sparsity_compensation_factor": 10.0}, {}),
    "fast_no_collision_budget": (FAST, 128, 96, 0.10, 2, {"max_consecutive_ray_collisions": 0}, {}),
    "fast_subsampling_1": (FAST, 128, 96, 0.10, 2, {"start_voxel_subsampling_factor": 1.0}, {}),
    "fast_p08": (FAST, 128, 96, 0.10, 2, {"semantic_measurement_probability": 0.8}, {}),
    "fast_clearing_rays": (FAST, 128, 96, 0.10, 2, {"max_ray_length_m": 2.5}, {}),
    "fast_no_clearing": (FAST, 128, 96, 0.10, 2, {"max_ray_length_m": 2.5, "allow_clear": 0}, {}),
    "fast_no_carving": (FAST, 128, 96, 0.10, 2, {"voxel_carving_enabled": 0, "max_ray_length_m": 2.5}, {}),
    "fast_freespace_cloud": (FAST, 128, 96, 0.10, 2, {}, {"freespace": True}),
    "fast_freespace_no_carving": (FAST, 128, 96, 0.10, 2, {"voxel_carving_enabled": 0}, {"freespace": True}),
    "fast_unknown_colors": (FAST, 128, 96, 0.10, 2, {}, {"unknown_colors": True}),
    "fast_clear_sets_every_2nd_frame": (FAST, 128, 96, 0.10, 4, {"clear_checks_every_n_frames": 2}, {}),
    "fast_min_range_gate": (FAST, 128, 96, 0.10, 2, {"min_ray_length_m": 2.0}, {}),
    # kimera_semantics_ros/launch/kimera_semantics.launch:98-122: 5 cm, 32 voxels per side, p = 0.8, semantic colours, fast
    "fast_launch_file_config": (FAST, 320, 240, 0.05, 3, {"voxels_per_side": 32, "semantic_measurement_probability": 0.8, "max_blocks": 1024}, {}),
    "fast_fullsize_640x480_5cm_4f": (FAST, 640, 480, 0.05, 4, {}, {}),             # BASELINE.json configs[1] geometry
    "merged_default_2f": (MERGED, 160, 120, 0.10, 2, {}, {}),
    "merged_fullsize_640x480_5cm_1f": (MERGED, 640, 480, 0.05, 1, {}, {}),
    # BASELINE.json configs[2] at full size: 640x480, 2 cm, `merged` (31 M voxel updates in the frame)
    "merged_fullsize_640x480_2cm_1f": (MERGED, 640, 480, 0.02, 1, {"max_updates": 80 << 20, "max_blocks": 32768}, {}),
    "merged_5cm": (MERGED, 128, 96, 0.05, 2, {}, {}),
    "merged_antigrazing": (MERGED, 128, 96, 0.10, 2, {"enable_anti_grazing": 1}, {}),
    "merged_clearing_rays": (MERGED, 128, 96, 0.10, 2, {"max_ray_length_m": 2.5}, {}),
    "merged_clearing_antigrazing": (MERGED, 128, 96, 0.10, 2, {"max_ray_length_m": 2.5, "enable_anti_grazing": 1}, {}),
    "merged_no_carving": (MERGED, 128, 96, 0.10, 2, {"voxel_carving_enabled": 0, "max_ray_length_m": 2.5}, {}),
    "merged_const_weight": (MERGED, 128, 96, 0.10, 2, {"use_const_weight": 1}, {}),
    "merged_color_mode_probability": (MERGED, 128, 96, 0.10, 2, {"color_mode": capi.KSG_COLOR_MODE_SEMANTIC_PROBABILITY}, {}),
    "merged_freespace_cloud": (MERGED, 128, 96, 0.10, 2, {}, {"freespace": True}),
    "merged_unknown_colors": (MERGED, 128, 96, 0.10, 2, {}, {"unknown_colors": True}),
}
KEYS = ("block_index", "tsdf_distance", "tsdf_weight", "tsdf_rgba", "sem_label", "sem_priors", "sem_rgba")


def case_config(name):
    itype, w, h, vs, nf, kw, sc = CASES[name]
    return make_config(itype, vs, C21, max_points=w * h, **{"max_updates": 8 << 20, **kw})


def case_frames(name, cfg):
    """Yields (T_G_C, points_C, rgba, freespace) exactly as every arm of the comparison receives them."""
    itype, w, h, vs, nf, kw, sc = CASES[name]
    pal = np.array([[cfg.label_color[l][k] for k in range(4)] for l in range(256)], np.uint8)
    for cam, depth, label, T in frames(w, h, C21, nf):
        xyz, pix = synth.backproject(depth, cam)
        rgba = np.ascontiguousarray(pal[label.reshape(-1)[pix]])
        if sc.get("unknown_colors"):
            rgba[::97] = (1, 2, 3, 255)
        yield T, xyz, rgba, bool(sc.get("freespace", False))


def color_table(cfg):
    pal = np.array([[cfg.label_color[l][k] for k in range(3)] for l in range(C21)], np.uint8)
    return pal, np.arange(C21, dtype=np.uint8)


def order_insensitive(exp):
    """What survives a change of the per-voxel update order (bundle order in `merged`)."""
    w = exp["tsdf_weight"].astype(np.float64)
    touched = (exp["tsdf_weight"] > 0) | (exp["sem_priors"] != np.float32(-0.60205999132)).any(axis=-1)
    return {"block_index": hashlib.sha256(np.ascontiguousarray(exp["block_index"]).tobytes()).hexdigest(),
            "observed_mask": hashlib.sha256(np.packbits(exp["tsdf_weight"] > 0).tobytes()).hexdigest(),
            "touched_mask": hashlib.sha256(np.packbits(touched).tobytes()).hexdigest(),
            "observed_voxels": int((exp["tsdf_weight"] > 0).sum()), "touched_voxels": int(touched.sum()),
            "weight_sum": float(w.sum())}


def digest(exp):
    d = {k: hashlib.sha256(np.ascontiguousarray(exp[k]).tobytes()).hexdigest() for k in KEYS}
    d["order_insensitive"] = order_insensitive(exp)
    return d


def run_case(name, make_integrator):
    """make_integrator(cfg) -> object with set_color_to_label (optional), integrate_points(T, xyz, rgba=, freespace=), export()."""
    cfg = case_config(name)
    integ = make_integrator(cfg)
    if hasattr(integ, "set_color_to_label"):
        integ.set_color_to_label(*color_table(cfg))
    for T, xyz, rgba, freespace in case_frames(name, cfg):
        integ.integrate_points(T, xyz, rgba=rgba, freespace=freespace)
    return integ.export()


if __name__ == "__main__":
    from oracle.ref_py import RefHybridIntegrator, available
    assert available(), "build oracle/_ref first: make -C oracle ref"
    out = {name: digest(run_case(name, RefHybridIntegrator)) for name in CASES}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_hybrid_golden.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, len(out), "cases")
