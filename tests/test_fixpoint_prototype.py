"""DESIGN.md section 4 claims that `fast`'s sequential, order-dependent observed-voxel set can be solved in parallel because the
dependence is triangular (unique fixpoint, reached from any start).  tools/observed_set_fixpoint.py states that model in ~60
lines; here it is checked against the sequential definition on real frame geometry (rays from the numpy float32 restatement in
test_oracle_crosscheck.py, which itself matches the oracle and, through it, the reference's own sources)."""
import os
import sys

import numpy as np
import pytest

from kimera_semantics_b200 import synth
from kimera_semantics_b200.capi import KSG_INTEGRATOR_FAST
from parity_utils import frames, make_config
from test_oracle_crosscheck import ApproxSet, f32, grid_index, index_hash, mixed_order, norm, raycast, transform

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import observed_set_fixpoint as fx  # noqa: E402


def cast_rays_of_frame(cfg, T, xyz, labels, start_set, offset):
    """Values (hash + offset) along every ray that survives the start-voxel set, in rank (ThreadSafeIndex) order."""
    vsi = f32(1.0 / f32(cfg.voxel_size))
    start_inv = f32(f32(cfg.start_voxel_subsampling_factor) * vsi)
    origin = T[4:].astype(np.float32)
    rays = []
    for i in mixed_order(len(xyz)):
        p = xyz[i]
        rng = norm(p)
        if rng < f32(cfg.min_ray_length_m) or cfg.dynamic_label[int(labels[i])]:
            continue
        clearing = bool(rng > f32(cfg.max_ray_length_m))
        pG = transform(T, p)
        if not start_set.replace(index_hash(grid_index(pG, start_inv))):
            continue
        vox = raycast(origin, pG, clearing, f32(cfg.max_ray_length_m), vsi, f32(cfg.default_truncation_distance), False)
        rays.append([index_hash(g) + offset for g in vox])
    return rays


@pytest.mark.parametrize("max_collisions", [0, 2])
def test_jacobi_sweeps_reach_the_sequential_result_from_any_start(max_collisions):
    C_, w, h = 5, 64, 48
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.10, C_, max_points=w * h, max_consecutive_ray_collisions=max_collisions)
    start_set = ApproxSet()
    table = {0: (1 << 64) - 1}     # persistent observed-set table (A.4 initial state)
    offset = 0
    rng = np.random.default_rng(3)
    for cam, depth, label, T in frames(w, h, C_, 3):
        offset += 1                                  # resetApproxSet() of both sets at the start of every frame
        start_set.reset()
        xyz, pix = synth.backproject(depth, cam)
        rays = cast_rays_of_frame(cfg, T, xyz, label.reshape(-1)[pix], start_set, offset)
        assert len(rays) > 300
        U_seq, table_after = fx.sequential(rays, table, max_collisions)
        lengths = [len(r) for r in rays]
        starts = {
            "lower bound (what the device uses)": [min(L, max_collisions) for L in lengths],
            "nothing performed": [0] * len(rays),
            "everything performed": lengths,
            "random": [int(rng.integers(0, L + 1)) for L in lengths],
        }
        sweeps = {}
        for name, U0 in starts.items():
            U, n = fx.solve(rays, table, max_collisions, U0)
            assert U == U_seq, name                  # unique fixpoint = the sequential answer
            sweeps[name] = n
        assert max(sweeps.values()) < 40 < len(rays)     # far fewer sweeps than the R + 1 bound
        assert sum(U_seq) < sum(lengths)                 # early termination really happens in these frames
        table = table_after                              # cross-frame persistence (stale entries can alias, A.4)


def test_first_sweep_is_exact_for_a_prefix_and_the_solution_is_a_fixpoint():
    """Triangularity: after k sweeps the first k rays are final (ray r depends only on ranks < r)."""
    rays = [[5, 6, 7, 8], [5, 6, 7, 9], [5, 6, 7, 8, 10], [11, 5, 6, 7, 12]]
    U_seq, _ = fx.sequential(rays, {}, 1)
    U = [0, 0, 0, 0]
    for k in range(1, len(rays) + 1):
        U = fx.jacobi_sweep(rays, {}, 1, U)
        assert U[:k] == U_seq[:k]
    assert fx.jacobi_sweep(rays, {}, 1, U_seq) == U_seq
