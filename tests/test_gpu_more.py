"""More GPU tests: golden fixtures without the oracle in the loop, the C++ drop-in shim end to end, BASELINE.json
full-size configurations (direct oracle comparison where the oracle finishes in seconds, size-independent properties
beyond that)."""
import importlib.util
import json
import os
import subprocess

import numpy as np
import pytest

from kimera_semantics_b200 import synth
from kimera_semantics_b200.capi import Integrator, KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED
from oracle.oracle_py import OracleIntegrator
from parity_utils import assert_parity, compare_maps, frames, make_config, stats_equal
from test_shim_cpu import CPP, demo, write_frames  # noqa: F401

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
GOLDEN = json.load(open(os.path.join(HERE, "golden", "oracle_golden.json")))


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_cuda_path_matches_committed_golden_digests(name):
    got = make_golden.run_case(name, Integrator)
    assert got["stats"] == GOLDEN[name]["stats"]
    assert got["digest"] == GOLDEN[name]["digest"]      # bit-exact map, no oracle involved at run time


def read_shim_output(path, vps, C):
    V = vps ** 3
    rec = np.dtype([("d", "<f4"), ("w", "<f4"), ("rgba", "u1", 4), ("label", "u1"), ("priors", "<f4", C), ("srgba", "u1", 4)])
    raw = open(path, "rb").read()
    nb = int(np.frombuffer(raw, "<i4", 1)[0])
    off = 4
    out = {"block_index": np.zeros((nb, 3), np.int32), "tsdf_distance": np.zeros((nb, V), np.float32), "tsdf_weight": np.zeros((nb, V), np.float32),
           "tsdf_rgba": np.zeros((nb, V, 4), np.uint8), "sem_label": np.zeros((nb, V), np.uint8), "sem_priors": np.zeros((nb, V, C), np.float32),
           "sem_rgba": np.zeros((nb, V, 4), np.uint8)}
    for b in range(nb):
        out["block_index"][b] = np.frombuffer(raw, "<i4", 3, off); off += 12
        a = np.frombuffer(raw, rec, V, off); off += V * rec.itemsize
        out["tsdf_distance"][b], out["tsdf_weight"][b], out["tsdf_rgba"][b] = a["d"], a["w"], a["rgba"]
        out["sem_label"][b], out["sem_priors"][b], out["sem_rgba"][b] = a["label"], a["priors"], a["srgba"]
    return out


REF_FACTORY_DEMO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "shim_demo_ref_factory")




@pytest.mark.parametrize("method,mode,factory", [("fast", "eager", "shim"), ("merged", "eager", "shim"), ("fast", "lazy", "shim"),
                                                 ("fast", "eager", "reference"),
                                                 ("merged", "eager", "reference")])
def test_cpp_shim_factory_and_integrate_match_oracle(demo, tmp_path, method, mode, factory):
    """SemanticTsdfIntegratorFactory::create(method, ...) + integratePointCloud(T_G_C, points_C, colors) through the C++ shim
    fill the host Layer<TsdfVoxel> / Layer<SemanticVoxel> exactly as the oracle's layers.
    factory="reference": the same driver linked with the reference's OWN semantic_tsdf_integrator_factory.cpp (compiled
    unmodified against the shim headers, `make -C oracle ref`), i.e. the reference's creation code constructs our classes."""
    if factory == "reference":
        if not os.path.exists(REF_FACTORY_DEMO):
            pytest.skip("oracle/_ref/shim_demo_ref_factory not built (needs /root/reference)")
        demo = REF_FACTORY_DEMO
    C, w, h, vs = 21, 320, 240, 0.10
    itype = KSG_INTEGRATOR_FAST if method == "fast" else KSG_INTEGRATOR_MERGED
    cfg = make_config(itype, vs, C, max_points=w * h)
    pal = np.array([[cfg.label_color[l][k] for k in range(4)] for l in range(C)], np.uint8)
    ora = OracleIntegrator(cfg)
    ora.set_color_to_label(pal[:, :3], np.arange(C, dtype=np.uint8))
    fr = []
    for cam, depth, label, T in frames(w, h, C, 2):
        xyz, pix = synth.backproject(depth, cam)
        rgba = pal[label.reshape(-1)[pix]].copy()
        rgba[::53] = (9, 8, 7, 255)                      # unknown colour -> label 0
        fr.append((T, xyz, rgba))
        ora.integrate_points(T, xyz, rgba=rgba)
    fpath, opath = tmp_path / "frames.bin", tmp_path / "out.bin"
    write_frames(fpath, fr, vs, 16, pal, [C - 1])
    env = dict(os.environ, KSG_MAX_POINTS=str(w * h), KSG_MAX_UPDATES=str(8 << 20))
    args = [demo, method, str(fpath), str(opath)] + (["lazy"] if mode == "lazy" else [])
    r = subprocess.run(args, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    got = read_shim_output(opath, 16, C)
    assert_parity(compare_maps(got, ora.export()))


def test_checkpoint_and_resume_through_the_shim_continues_bit_exactly(demo, tmp_path):
    """merged has no cross-frame integrator state, so integrate(0..1) -> saveMap | new process: loadMap -> integrate(2..3) must
    equal integrate(0..3) in every voxel of both host layers."""
    C, w, h, vs = 21, 160, 120, 0.10
    cfg = make_config(KSG_INTEGRATOR_MERGED, vs, C, max_points=w * h)
    pal = np.array([[cfg.label_color[l][k] for k in range(4)] for l in range(C)], np.uint8)
    fr = []
    for cam, depth, label, T in frames(w, h, C, 4):
        xyz, pix = synth.backproject(depth, cam)
        fr.append((T, xyz, pal[label.reshape(-1)[pix]].copy()))
    fpath, ckpt = tmp_path / "frames.bin", tmp_path / "map.ksgm"
    write_frames(fpath, fr, vs, 16, pal, [C - 1])
    write_frames(tmp_path / "first.bin", fr[:2], vs, 16, pal, [C - 1])
    env = dict(os.environ, KSG_MAX_POINTS=str(w * h), KSG_MAX_UPDATES=str(8 << 20))
    runs = {"all": [demo, "merged", str(fpath), str(tmp_path / "all.bin")],
            "first": [demo, "merged", str(tmp_path / "first.bin"), str(tmp_path / "first_out.bin"), "--save", str(ckpt)],
            "resumed": [demo, "merged", str(fpath), str(tmp_path / "resumed.bin"), "--load", str(ckpt), "--skip", "2"]}
    for name in ("all", "first", "resumed"):
        r = subprocess.run(runs[name], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, (name, r.stderr + r.stdout)
    assert open(tmp_path / "resumed.bin", "rb").read() == open(tmp_path / "all.bin", "rb").read()
    assert open(tmp_path / "first_out.bin", "rb").read() != open(tmp_path / "all.bin", "rb").read()


BINDING_CHECK = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "gpu_binding_check")


@pytest.mark.parametrize("method", ["fast", "merged"])
def test_integration_md_binding_against_reference_headers_matches_oracle(tmp_path, method):
    """integration/kimera_semantics/semantic_tsdf_integrator_gpu.h (what a kimera_semantics maintainer adds) compiled against the
    reference's REAL SemanticIntegratorBase / SemanticLabel2Color / SemanticVoxel and driven like the shim demo: both host layers
    must equal the oracle's (merged in the reference's bundle order, which the binding selects)."""
    if not os.path.exists(BINDING_CHECK):
        pytest.skip("oracle/_ref/gpu_binding_check not built (needs /root/reference)")
    C, w, h, vs = 21, 320, 240, 0.10
    itype = KSG_INTEGRATOR_FAST if method == "fast" else KSG_INTEGRATOR_MERGED
    cfg = make_config(itype, vs, C, max_points=w * h)
    pal = np.array([[cfg.label_color[l][k] for k in range(4)] for l in range(C)], np.uint8)
    ora = OracleIntegrator(cfg, canonical_merged=False)
    ora.set_color_to_label(pal[:, :3], np.arange(C, dtype=np.uint8))
    fr = []
    for cam, depth, label, T in frames(w, h, C, 2):
        xyz, pix = synth.backproject(depth, cam)
        rgba = pal[label.reshape(-1)[pix]].copy()
        rgba[::53] = (9, 8, 7, 255)
        fr.append((T, xyz, rgba))
        ora.integrate_points(T, xyz, rgba=rgba)
    fpath, opath = tmp_path / "frames.bin", tmp_path / "out.bin"
    write_frames(fpath, fr, vs, 16, pal, [C - 1])
    r = subprocess.run([BINDING_CHECK, method, str(fpath), str(opath)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout
    assert_parity(compare_maps(read_shim_output(opath, 16, C), ora.export()))


# ---- BASELINE.json full-size configurations -------------------------------------------------------------------------
def test_fullsize_config1_fast_640x480_5cm_c21_ten_frames():
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.05, 21)
    gpu, ora = Integrator(cfg), OracleIntegrator(cfg)
    for i, (cam, depth, label, T) in enumerate(frames(640, 480, 21, 10)):
        sg, so = gpu.integrate_depth(T, depth, label, cam.K), ora.integrate_depth(T, depth, label, cam.K)
        ok, why = stats_equal(sg, so)
        assert ok, f"frame {i}: {why}"
    assert_parity(compare_maps(gpu.export(), ora.export()))


def test_fullsize_config2_merged_640x480_2cm_c21_one_frame_vs_oracle():
    cfg = make_config(KSG_INTEGRATOR_MERGED, 0.02, 21, max_updates=48 << 20, max_blocks=4096)
    gpu, ora = Integrator(cfg), OracleIntegrator(cfg)
    cam, depth, label, T = next(frames(640, 480, 21, 1))
    sg, so = gpu.integrate_depth(T, depth, label, cam.K), ora.integrate_depth(T, depth, label, cam.K)   # ~10 s of CPU
    ok, why = stats_equal(sg, so)
    assert ok, why
    assert sg.voxel_updates > 25_000_000
    assert_parity(compare_maps(gpu.export(), ora.export()))


def _invariants(cfg, e):
    C = cfg.num_labels
    trunc = np.float32(cfg.default_truncation_distance)
    assert np.isfinite(e["tsdf_distance"]).all() and (np.abs(e["tsdf_distance"]) <= trunc).all()
    assert (e["tsdf_weight"] >= 0).all() and (e["tsdf_weight"] <= cfg.max_weight).all()
    assert (e["sem_priors"] <= np.float32(-0.60205999132)).all()      # every log-likelihood increment is <= 0
    assert (e["sem_label"] < C).all()
    lut = np.array([[cfg.label_color[l][k] for k in range(4)] for l in range(256)], np.uint8)
    touched = (e["sem_priors"] < np.float32(-0.60205999132)).any(axis=-1)
    assert np.array_equal(e["sem_rgba"][touched], lut[e["sem_label"][touched]])     # base.cpp:370-380
    assert np.array_equal(e["tsdf_rgba"][touched], e["sem_rgba"][touched])          # kSemantic hand-off base.cpp:177-180
    best = e["sem_priors"].max(axis=-1)
    first = (e["sem_priors"] == best[..., None]).argmax(axis=-1)
    assert np.array_equal(first.astype(np.uint8), e["sem_label"])                   # label = first arg-max of the priors


def test_fullsize_fast_weight_checksum_property_40_frames():
    """Size-independent property: with constant weights and no drop-off every update adds exactly 1.0 to its voxel, so the
    sum of all voxel weights equals the number of voxel updates (a checksum over the whole stream, no oracle needed)."""
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.05, 21, use_const_weight=1, use_weight_dropoff=0)
    gpu = Integrator(cfg)
    total = 0
    for cam, depth, label, T in frames(640, 480, 21, 40):
        total += gpu.integrate_depth(T, depth, label, cam.K).voxel_updates
    e = gpu.export()
    assert e["tsdf_weight"].max() < cfg.max_weight
    assert int(e["tsdf_weight"].astype(np.float64).sum()) == total
    assert (e["tsdf_weight"] == np.round(e["tsdf_weight"])).all()
    _invariants(cfg, e)


def test_fullsize_merged_2cm_invariants_three_frames():
    cfg = make_config(KSG_INTEGRATOR_MERGED, 0.02, 21, max_updates=48 << 20, max_blocks=4096)
    gpu = Integrator(cfg)
    for cam, depth, label, T in frames(640, 480, 21, 3):
        st = gpu.integrate_depth(T, depth, label, cam.K)
        assert st.voxel_updates == st.ray_steps > 25_000_000
    _invariants(cfg, gpu.export())


# BASELINE.json configs[3] geometry on one GPU: 1280x720, 5 cm, 150 classes (one frame of the 8-frame batch)
def test_config3_1280x720_5cm_c150_one_frame():
    for itype in (KSG_INTEGRATOR_FAST,):
        cfg = make_config(itype, 0.05, 150, max_points=1280 * 720, max_blocks=2048)
        gpu, ora = Integrator(cfg), OracleIntegrator(cfg)
        for cam, depth, label, T in frames(1280, 720, 150, 2):
            sg, so = gpu.integrate_depth(T, depth, label, cam.K), ora.integrate_depth(T, depth, label, cam.K)
            ok, why = stats_equal(sg, so)
            assert ok, why
        assert_parity(compare_maps(gpu.export(), ora.export()))


# ---- spatial hash-block sharding of one map: shards run one after the other on ONE device here -------------
@pytest.mark.parametrize("itype,G", [(KSG_INTEGRATOR_FAST, 2), (KSG_INTEGRATOR_FAST, 3), (KSG_INTEGRATOR_MERGED, 2), (KSG_INTEGRATOR_MERGED, 4)])
def test_spatially_sharded_map_equals_unsharded(itype, G):
    """Every shard sees every frame and casts every ray but applies only the tiles it owns; assembling the per-shard exports
    with the ownership masks must give exactly the single-integrator (= oracle) map."""
    from kimera_semantics_b200.capi import merge_shard_exports
    C_, w, h = 21, 320, 240
    cfg = make_config(itype, 0.05, C_, max_points=w * h, max_updates=8 << 20)
    ora = OracleIntegrator(cfg)
    shards = []
    for r in range(G):
        c = make_config(itype, 0.05, C_, max_points=w * h, max_updates=8 << 20, shard_rank=r, shard_count=G)
        shards.append(Integrator(c))
    for cam, depth, label, T in frames(w, h, C_, 3):
        so = ora.integrate_depth(T, depth, label, cam.K)
        tiles = 0
        for s in shards:
            st = s.integrate_depth(T, depth, label, cam.K)
            assert st.voxel_updates == so.voxel_updates and st.blocks_allocated == so.blocks_allocated
            tiles += st.tiles_touched
        assert tiles > 0
    merged = merge_shard_exports([s.export() for s in shards], 16)
    assert_parity(compare_maps(merged, ora.export()))
    # each shard really skipped work: its own export differs from the full map somewhere
    assert any((s.export()["tsdf_weight"] != merged["tsdf_weight"]).any() for s in shards)
    for s in shards:
        s.close()


def test_import_blocks_restores_a_map_and_integration_continues_identically():
    """ksg_import_blocks (SURVEY.md 8f NEXT-3): export -> import into a fresh integrator -> identical export; for `merged` (no
    cross-frame state besides the map) integrating further frames into the restored map equals the uninterrupted run."""
    C_, w, h = 21, 320, 240
    cfg = make_config(KSG_INTEGRATOR_MERGED, 0.05, C_, max_points=w * h, max_updates=8 << 20)
    a, ora = Integrator(cfg), OracleIntegrator(cfg)
    fr = list(frames(w, h, C_, 4))
    for cam, depth, label, T in fr[:2]:
        a.integrate_depth(T, depth, label, cam.K)
        ora.integrate_depth(T, depth, label, cam.K)
    snap = a.export()
    b = Integrator(cfg)
    b.import_blocks(snap)
    assert b.num_blocks() == a.num_blocks()
    rep = compare_maps(b.export(), snap)
    assert rep["same_blocks"] == 1 and rep["tsdf_distance_bit_mismatch"] == 0 and rep["sem_priors_bit_mismatch"] == 0 and rep["label_mismatch"] == 0
    for cam, depth, label, T in fr[2:]:
        sb, so = b.integrate_depth(T, depth, label, cam.K), ora.integrate_depth(T, depth, label, cam.K)
        ok, why = stats_equal(sb, so)
        assert ok, why
    assert_parity(compare_maps(b.export(), ora.export()))
    # partial import: only distances of two blocks, everything else untouched
    two = {k: v[:2].copy() for k, v in snap.items()}
    two["tsdf_distance"][:] = 0.125
    c = Integrator(make_config(KSG_INTEGRATOR_FAST, 0.05, C_, max_points=w * h))
    c.import_blocks(snap)
    c._check(c.lib.ksg_import_blocks(c.handle, 2, two["block_index"].ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_int32)),
                                     two["tsdf_distance"].ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)),
                                     None, None, None, None, None), "partial import")
    e = c.export()
    idx = [int(np.where((e["block_index"] == two["block_index"][i]).all(axis=1))[0][0]) for i in range(2)]
    assert (e["tsdf_distance"][idx] == 0.125).all()
    mask = np.ones(len(e["block_index"]), bool); mask[idx] = False
    assert np.array_equal(e["tsdf_distance"][mask], snap["tsdf_distance"][mask]) and np.array_equal(e["tsdf_weight"], snap["tsdf_weight"])


@pytest.mark.gpu
def test_fast_sets_full_reset_after_10000_frames():
    """ApproxHashSet::resetApproxSet (A.4): the per-frame offset reaches full_reset_threshold = 10 000, both tables are
    wiped and the offset restarts.  10 012 tiny clouds through the C-ABI vs the oracle (the same sequence is checked against
    the reference's own sources in tests/test_oracle_vs_ref_hybrid.py)."""
    from test_oracle_vs_ref_hybrid import tiny_frames
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.10, 21, max_points=64, max_blocks=512)
    gpu, ora = Integrator(cfg), OracleIntegrator(cfg)
    worst = None
    for i, (T, xyz, lab) in enumerate(tiny_frames(10012)):
        sg = gpu.integrate_points(T, xyz, labels=lab)
        so = ora.integrate_points(T, xyz, labels=lab)
        if worst is None and (sg.voxel_updates, sg.rays_cast) != (so.voxel_updates, so.rays_cast):
            worst = (i, sg.as_dict(), so.as_dict())
    assert worst is None, f"first frame whose counters differ: {worst}"
    assert_parity(compare_maps(gpu.export(), ora.export()))
    gpu.close()


def test_depth_entry_with_float64_intrinsics_matches_oracle():
    """fx = 415.69219381653056 (60 degree FOV, 480 lines) is not representable in float; the k64 entry must reproduce the reference's
    float(1.0 / fx_double) exactly (depth_map_to_pointcloud.h:228-230)."""
    cfg = make_config(KSG_INTEGRATOR_FAST, 0.10, 21, max_points=320 * 240)
    gpu, ora = Integrator(cfg), OracleIntegrator(cfg)
    for cam, depth, label, T in frames(320, 240, 21, 2):
        K64 = np.array([415.69219381653056 / 2, 415.69219381653056 / 2 * 1.0001, 159.5 + 0.123456789, 119.5], np.float64)
        sg, so = gpu.integrate_depth_k64(T, depth, label, K64), ora.integrate_depth_k64(T, depth, label, K64)
        ok, why = stats_equal(sg, so)
        assert ok, why
    assert_parity(compare_maps(gpu.export(), ora.export()))
    gpu.close()


def test_shim_depth_frame_entry_matches_oracle(demo, tmp_path):
    C, w, h, vs = 21, 320, 240, 0.10
    cfg = make_config(KSG_INTEGRATOR_FAST, vs, C, max_points=w * h)
    pal = np.array([[cfg.label_color[l][k] for k in range(4)] for l in range(C)], np.uint8)
    ora = OracleIntegrator(cfg)
    K64 = np.array([207.84609690826528, 207.9, 159.5, 119.5], np.float64)
    dpath = tmp_path / "depth.bin"
    with open(dpath, "wb") as f:
        fr = list(frames(w, h, C, 2))
        f.write(np.int32(len(fr)).tobytes() + np.int32(w).tobytes() + np.int32(h).tobytes() + K64.tobytes())
        for cam, depth, label, T in fr:
            f.write(np.ascontiguousarray(T, np.float32).tobytes() + np.ascontiguousarray(depth, np.float32).tobytes() + np.ascontiguousarray(label, np.uint8).tobytes())
            ora.integrate_depth_k64(T, depth, label, K64)
    fpath, opath = tmp_path / "frames.bin", tmp_path / "out.bin"
    write_frames(fpath, [], vs, 16, pal, [C - 1])
    env = dict(os.environ, KSG_MAX_POINTS=str(w * h))
    r = subprocess.run([demo, "fast", str(fpath), str(opath), "--depth", str(dpath)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    assert_parity(compare_maps(read_shim_output(opath, 16, C), ora.export()))


def test_warp_chain_scan_equals_sequential_float_loop():
    """ksg_debug_chain_sum: one warp, lanes = records, composition of two-entry tables over shuffles; must reproduce the sequential
    float32 recurrence of a hot `merged` voxel bit for bit (tests/test_exact_float_chain.py holds the same claim for the CPU model)."""
    from kimera_semantics_b200 import capi
    from test_exact_float_chain import realistic_terms, xc, F, bits
    rng = np.random.default_rng(11)
    for n, s0 in ((1, -0.60205999132), (31, -0.60205999132), (33, -5.0), (4097, -0.60205999132), (92000, -0.60205999132), (92000, -2.5e5)):
        terms = realistic_terms(rng, n)
        assert bits(capi.debug_chain_sum(terms, s0)) == bits(xc.sequential(F(s0), terms)), (n, s0)
    mant = rng.integers(1 << 23, 1 << 24, 6000)
    terms = -np.ldexp(mant.astype(np.float64), rng.integers(-30, 6, 6000) - 23).astype(np.float32)
    terms[rng.random(6000) < 0.05] = 0.0
    for s0 in (-1e-3, -777.25, -3.0e7):
        assert bits(capi.debug_chain_sum(terms, s0)) == bits(xc.sequential(F(s0), terms)), s0
