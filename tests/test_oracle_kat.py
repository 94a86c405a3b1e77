"""Known-answer tests that pin the CPU oracle (oracle/ks_oracle.cpp).  The reference ships no tests or golden
vectors for this path ("parity unpinned"), so every expected value below is derived
independently of the oracle: by hand, in float64 numpy, or from the closed forms in SURVEY.md Appendix A."""
import ctypes as C

import numpy as np
import pytest

from kimera_semantics_b200.capi import KSG_INTEGRATOR_FAST, _ptr, default_config
from oracle import oracle_py

lib = oracle_py.load()
F = lambda a: np.ascontiguousarray(a, np.float32)


def test_index_hash_values():
    # LongIndexHash: (x + y*17191 + z*17191^2) mod 2^64 truncated to 32 bits (A.2)
    sl = 17191
    for x, y, z in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (-1, 0, 0), (3, -7, 11), (-100, 250, -3), (2**31 - 1, -2**31, 5)]:
        want = (x + y * sl + z * sl * sl) % (1 << 64) % (1 << 32)
        assert lib.kso_index_hash(x, y, z) == want
    # the 2^20-slot alias pair derived in DESIGN.md: (x, y, z) and (x - 75, y + 61, z) share a slot
    a, b = lib.kso_index_hash(10, 20, 30), lib.kso_index_hash(10 - 75, 20 + 61, 30)
    assert (a & 0xFFFFF) == (b & 0xFFFFF) and a != b


def test_mixed_thread_safe_index():
    n = 307200  # 640x480 -> 300 groups of 1024 (A.3)
    first = [lib.kso_mixed_index(n, s) for s in range(4)]
    assert first == [0, 1024, 2048, 3072]
    assert lib.kso_mixed_index(n, 300) == 1 and lib.kso_mixed_index(n, 301) == 1025
    n = 5000  # 4 groups, tail 4096..4999 maps to itself
    perm = [lib.kso_mixed_index(n, s) for s in range(n)]
    assert sorted(perm) == list(range(n))
    assert perm[:5] == [0, 1024, 2048, 3072, 1] and perm[4096:4100] == [4096, 4097, 4098, 4099]
    assert [lib.kso_mixed_index(100, s) for s in range(100)] == list(range(100))  # fewer than 1024 points: identity


def test_transformation_matches_float64_rotation():
    rng = np.random.default_rng(0)
    for _ in range(50):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        t = rng.normal(size=3) * 3
        p = rng.normal(size=3) * 4
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        out = np.zeros(3, np.float32)
        lib.kso_transform(_ptr(F(np.concatenate([q, t])), C.c_float), _ptr(F(p), C.c_float), _ptr(out, C.c_float))
        assert np.allclose(out, R @ p + t, atol=2e-5)
    # identity rotation is exact
    out = np.zeros(3, np.float32)
    lib.kso_transform(_ptr(F([1, 0, 0, 0, 1, 2, 3]), C.c_float), _ptr(F([0.5, 0.25, -4]), C.c_float), _ptr(out, C.c_float))
    assert out.tolist() == [1.5, 2.25, -1.0]


def test_grid_index_epsilon_and_negative_coordinates():
    out = np.zeros(3, np.int64)
    inv = np.float32(20.0)  # 5 cm voxels
    lib.kso_grid_index(_ptr(F([0.05, -0.05, 0.0499]), C.c_float), inv, _ptr(out, C.c_int64))
    # 0.05*20 = 1.0000000149 (float32 0.05 > 0.05) -> 1 ; -0.05*20+1e-6 -> floor(-0.999999) = -1 ; 0.998+1e-6 -> 0
    assert out.tolist() == [1, -1, 0]
    lib.kso_grid_index(_ptr(F([-1e-7, 2.0, -3.0]), C.c_float), np.float32(1.0), _ptr(out, C.c_int64))
    assert out.tolist() == [0, 2, -3]  # the +1e-6 epsilon pulls -1e-7 into cell 0 (A.2)


def test_block_and_local_index():
    blk, loc = np.zeros(3, np.int32), np.zeros(3, np.int32)
    g = np.array([-1, 16, 35], np.int64)
    lib.kso_block_and_local(_ptr(g, C.c_int64), 16, _ptr(blk, C.c_int32), _ptr(loc, C.c_int32))
    assert blk.tolist() == [-1, 1, 2] and loc.tolist() == [15, 0, 3]
    g = np.array([-16, -17, 15], np.int64)
    lib.kso_block_and_local(_ptr(g, C.c_int64), 16, _ptr(blk, C.c_int32), _ptr(loc, C.c_int32))
    assert blk.tolist() == [-1, -2, 0] and loc.tolist() == [0, 15, 15]


def _raycast(origin, point, clearing=0, carving=1, max_len=5.0, vsi=1.0, trunc=0.0, from_origin=1, cap=4096):
    buf = np.zeros((cap, 3), np.int64)
    n = lib.kso_raycast(_ptr(F(origin), C.c_float), _ptr(F(point), C.c_float), clearing, carving, np.float32(max_len),
                        np.float32(vsi), np.float32(trunc), from_origin, _ptr(buf, C.c_int64), cap)
    return buf[:n]


def test_raycaster_hand_checked_sequences():
    # unit voxels, ray from (0.5,0.5,0.5) to (3.5,1.5,0.6): dx=3, dy=1, dz=0.1
    # x crossings at t=1/6,1/2,5/6 ; y crossing at t=1/2 (tie with x at t=1/2 -> x first: first minimum wins)
    seq = _raycast([0.5, 0.5, 0.5], [3.5, 1.5, 0.6]).tolist()
    assert seq == [[0, 0, 0], [1, 0, 0], [2, 0, 0], [2, 1, 0], [3, 1, 0]]
    # reversed traversal (fast integrator: cast_from_origin = false) starts at the far end
    seq = _raycast([0.5, 0.5, 0.5], [3.5, 1.5, 0.6], from_origin=0).tolist()
    assert seq[0] == [3, 1, 0] and seq[-1] == [0, 0, 0] and len(seq) == 5
    # truncation extends the ray beyond the point by trunc along the ray
    # (exactly axis-aligned rays are degenerate upstream: 0/0 in t_step, A.7 - every component is kept non-zero)
    seq = _raycast([0.5, 0.5, 0.5], [0.6, 0.7, 2.5], trunc=1.0, vsi=1.0)
    assert seq[-1].tolist() == [0, 0, 3] and len(seq) == 4
    # clearing ray stops trunc short of the point and is limited by max_len
    seq = _raycast([0.25, 0.25, 0.25], [0.3, 0.35, 9.25], clearing=1, trunc=1.0, max_len=5.0)
    assert seq[-1].tolist() == [0, 0, 5] and len(seq) == 6
    # no carving: only the truncation band around the surface
    seq = _raycast([0.5, 0.5, 0.5], [0.6, 0.7, 5.5], carving=0, trunc=1.0)
    assert seq[:, 2].tolist() == [4, 5, 6]


def test_raycaster_invariants_random():
    rng = np.random.default_rng(1)
    for _ in range(200):
        o = rng.uniform(-3, 3, 3)
        p = o + rng.normal(size=3) * rng.uniform(0.3, 6)
        seq = _raycast(o, p, vsi=20.0, trunc=0.2, max_len=50.0)
        d = np.abs(np.diff(seq, axis=0)).sum(axis=1)
        assert (d == 1).all()                      # 6-connected path
        assert len(seq) == np.abs(seq[-1] - seq[0]).sum() + 1   # ray_length_in_steps + 1 indices
        assert seq[0].tolist() == np.floor(np.float32(o) * np.float32(20.0) + np.float32(1e-6)).astype(int).tolist()


def _tsdf_seq(cfg, sdfs, weights, colors=None):
    """Voxel (0,0,0) of a 1 m grid (centre 0.5^3), origin at x=-10 on the voxel's axis: a point at distance d along +x
    gives sdf = d - 10.5 exactly representable for the values used."""
    n = len(sdfs)
    origin = F([-10.0, 0.5, 0.5])
    pts = np.zeros((n, 3), np.float32)
    pts[:, 0] = np.float32(-10.0) + np.float32(10.5) + F(sdfs)
    pts[:, 1:] = 0.5
    rgba = np.zeros((n, 4), np.uint8) if colors is None else np.ascontiguousarray(colors, np.uint8)
    out, oc = np.zeros(2, np.float32), np.zeros(4, np.uint8)
    g = np.zeros(3, np.int64)
    lib.kso_tsdf_update_sequence(C.byref(cfg), _ptr(origin, C.c_float), _ptr(pts, C.c_float), _ptr(F(weights), C.c_float),
                                 _ptr(rgba, C.c_uint8), n, _ptr(g, C.c_int64), _ptr(out, C.c_float), _ptr(oc, C.c_uint8))
    return out, oc


def test_tsdf_update_is_order_dependent_clamp_after_average():
    # equal weights, truncation 0.2: (1.0 then 0.1) -> 0.15 ; (0.1 then 1.0) -> 0.2
    cfg = default_config(KSG_INTEGRATOR_FAST, 1.0, 16, 5)
    cfg.default_truncation_distance = 0.2
    cfg.use_weight_dropoff = 0
    a, _ = _tsdf_seq(cfg, [1.0, 0.125], [1.0, 1.0])
    b, _ = _tsdf_seq(cfg, [0.125, 1.0], [1.0, 1.0])
    assert a[0] == pytest.approx((0.125 + 0.2) / 2, abs=1e-7) and b[0] == pytest.approx(0.2, abs=1e-7)
    assert a[1] == 2.0 and b[1] == 2.0


def test_tsdf_weight_dropoff_cap_and_tiny_weights():
    cfg = default_config(KSG_INTEGRATOR_FAST, 0.5, 16, 5)   # voxel 0.5 m, trunc 2.0
    cfg.voxel_size = 1.0                                    # centre of voxel 0 stays 0.5
    cfg.default_truncation_distance = 2.0
    # sdf = -1.5 < -voxel_size(1.0): weight * (2.0 - 1.5) / (2.0 - 1.0) = 0.5
    out, _ = _tsdf_seq(cfg, [-1.5], [1.0])
    assert out.tolist() == [-1.5, 0.5]
    # behind the truncation band the drop-off clamps to 0 -> new_weight < 1e-6 -> voxel untouched
    out, _ = _tsdf_seq(cfg, [-2.5], [1.0])
    assert out.tolist() == [0.0, 0.0]
    cfg.max_weight = 3.0
    out, _ = _tsdf_seq(cfg, [0.5] * 5, [1.0] * 5)
    assert out.tolist() == [0.5, 3.0]


def test_colour_blending_rounds_half_away_and_only_near_surface():
    o = np.zeros(4, np.uint8)
    lib.kso_blend(_ptr(np.array([10, 0, 255, 255], np.uint8), C.c_uint8), np.float32(1.0),
                  _ptr(np.array([11, 1, 0, 255], np.uint8), C.c_uint8), np.float32(1.0), _ptr(o, C.c_uint8))
    assert o.tolist() == [11, 1, 128, 255]  # 10.5 -> 11, 0.5 -> 1, 127.5 -> 128
    cfg = default_config(KSG_INTEGRATOR_FAST, 1.0, 16, 5)
    cfg.default_truncation_distance = 0.25
    cfg.use_weight_dropoff = 0
    _, c = _tsdf_seq(cfg, [0.125, 1.0], [1.0, 3.0], colors=[[200, 100, 0, 255], [0, 0, 0, 0]])
    assert c.tolist() == [200, 100, 0, 255]  # |sdf| >= trunc: colour untouched by the second update


@pytest.mark.parametrize("p,lm,ln", [(0.9, -0.105360545, -2.30258489), (0.8, -0.223143533, -1.60943794)])
def test_log_likelihood_matrix(p, lm, ln):
    C_ = 7
    cfg = default_config(KSG_INTEGRATOR_FAST, 0.05, 16, C_)
    cfg.semantic_measurement_probability = p
    L, two = np.zeros((C_, C_), np.float32), np.zeros(2, np.float32)
    lib.kso_log_likelihood(C.byref(cfg), _ptr(L, C.c_float), _ptr(two, C.c_float))
    assert two[0] == pytest.approx(lm, rel=2e-7) and two[1] == pytest.approx(ln, rel=2e-7)
    assert (L[:, 0] == 0).all()                                   # column of the unknown label is zero (base.cpp:127)
    assert (np.diag(L)[1:] == two[0]).all()
    off = L[:, 1:][~np.eye(C_, dtype=bool)[:, 1:]]
    assert (off == two[1]).all()


def test_semantic_update_closed_form_and_first_max_tie_break():
    C_ = 5
    cfg = default_config(KSG_INTEGRATOR_FAST, 0.05, 16, C_)
    lm, ln = np.float32(np.log(np.float32(0.9))), np.float32(np.log(np.float32(1) - np.float32(0.9)))
    init = np.float32(-0.60205999132)

    def run(freqs):
        f = F(freqs)
        pri, lab, sc, tc = np.zeros(C_, np.float32), np.zeros(1, np.uint8), np.zeros(4, np.uint8), np.zeros(4, np.uint8)
        lib.kso_semantic_update_sequence(C.byref(cfg), _ptr(f, C.c_float), len(f), _ptr(pri, C.c_float), _ptr(lab, C.c_uint8),
                                         _ptr(sc, C.c_uint8), _ptr(tc, C.c_uint8))
        return pri, int(lab[0]), sc, tc

    onehot = lambda l: np.eye(C_, dtype=np.float32)[l]
    # observing label 0 changes nothing; arg-max of the uniform prior is label 0 (first maximum)
    pri, lab, sc, tc = run([onehot(0)])
    assert (pri == init).all() and lab == 0
    # one observation of label 3 (A.9): prior[3] += lm, every other entry (also entry 0) += ln
    pri, lab, sc, tc = run([onehot(3)])
    want = np.full(C_, init + ln, np.float32); want[3] = init + lm
    assert (pri == want).all() and lab == 3
    assert sc.tolist() == [cfg.label_color[3][k] for k in range(4)] and tc.tolist() == sc.tolist()   # kSemantic hand-off
    # tie between labels 2 and 4 -> first maximum (2) wins
    pri, lab, _, _ = run([onehot(2), onehot(4)])
    assert pri[2] == pri[4] and lab == 2
    # histogram update: sequential j-ascending sum, one multiply and one add per term
    hist = np.array([3, 0, 2, 0, 1], np.float32)
    pri, lab, _, _ = run([hist])
    want = np.zeros(C_, np.float32)
    for i in range(C_):
        acc = np.float32(0)
        for j in range(C_):
            Lij = np.float32(0) if j == 0 else (lm if i == j else ln)
            acc = np.float32(acc + np.float32(Lij * hist[j]))
        want[i] = np.float32(init + acc)
    assert (pri == want).all() and lab == 2


def test_approx_hash_set_semantics_and_cross_frame_alias():
    RESET = np.uint64(0xFFFFFFFFFFFFFFFF)

    def script(ops):
        ops = np.array(ops, np.uint64)
        res = np.zeros(len(ops), np.uint8)
        lib.kso_approx_set_script(_ptr(ops, C.c_uint64), len(ops), _ptr(res, C.c_uint8))
        return res.tolist()

    h = 123456
    # replaceHash: true the first time, false while the slot still holds the value, true again after an aliasing value
    assert script([h, h, h + (1 << 20), h, h]) == [1, 0, 1, 1, 0]
    # hash 0 is not "present" in a fresh set (table[0] = SIZE_MAX)
    assert script([0, 0]) == [1, 0]
    # reset only bumps the offset: (h) at offset 0 stores h; after one reset, hash h-1 maps to value h -> reads as present
    assert script([h, RESET, h - 1, h]) == [1, 2, 0, 1]


def test_rainbow_colour_map_anchor_points():
    o = np.zeros(4, np.uint8)
    for hval, want in [(0.0, [255, 0, 0, 255]), (1.0 / 6, [255, 255, 0, 255]), (0.5, [0, 255, 255, 255]), (1.0, [255, 0, 0, 255])]:
        lib.kso_rainbow(hval, _ptr(o, C.c_uint8))
        assert o.tolist() == want
