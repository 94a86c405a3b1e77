"""The generated marching-cubes table (tools/make_mc_table.py -> csrc/ksg_mc_table.h) and the numpy mesher built on it (tests/mesh_ref.py,
the twin of csrc/ksg_mesh.cuh): table invariants, watertightness and orientation on analytic distance fields.  CPU only."""
import os
from collections import Counter

import numpy as np

import mesh_ref as mr
from mesh_ref import mc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_is_the_generators_output():
    assert open(os.path.join(ROOT, "kimera_semantics_b200", "csrc", "ksg_mc_table.h")).read() == mc.header_text()


def test_every_configuration_uses_exactly_its_crossed_edges_and_closes_on_the_faces():
    table = mc.build_table()
    assert table[0] == [] and table[255] == []
    assert max(len(r) for r in table) == 15
    for cfg in range(256):
        row = table[cfg]
        crossed = {e for e, (a, b) in enumerate(mc.EDGES) if ((cfg >> a) & 1) != ((cfg >> b) & 1)}
        assert set(row) == crossed, cfg
        # boundary of the patch (triangle sides used once) = the face segments, which depend on the face's corner signs only:
        # two cubes sharing a face cut it identically -> no cracks
        sides = Counter()
        for t in range(0, len(row), 3):
            tri = row[t:t + 3]
            assert len(set(tri)) == 3, cfg
            for k in range(3):
                sides[frozenset((tri[k], tri[(k + 1) % 3]))] += 1
        boundary = {s for s, n in sides.items() if n == 1}
        assert all(n <= 2 for n in sides.values()), cfg
        want = set()
        for face in mc.FACES:
            for a, b in mc.face_segments(cfg, face):
                want.add(frozenset((a, b)))
        assert boundary == want, cfg


def test_ambiguous_faces_depend_on_the_face_signs_only():
    # the x = 1 face of one cube is the x = 0 face of its neighbour: same corner signs -> same pairs of crossed edges (by position on the face)
    lo, hi = mc.FACES[4], mc.FACES[5]          # (0, 3, 7, 4) and (1, 2, 6, 5): corner k of one coincides with corner k of the other
    pos = lambda face, e: sorted(face.index(c) for c in mc.EDGES[e])
    for cfg_a in range(256):
        signs = [(cfg_a >> c) & 1 for c in hi]
        cfg_b = sum(s << c for s, c in zip(signs, lo))
        seg_a = sorted(sorted([pos(hi, a), pos(hi, b)]) for a, b in mc.face_segments(cfg_a, hi))
        seg_b = sorted(sorted([pos(lo, a), pos(lo, b)]) for a, b in mc.face_segments(cfg_b, lo))
        assert seg_a == seg_b, cfg_a


def _closed_and_outward(mesh, centre_fn, min_area_frac=1.0):
    v = mesh["vertices"].astype(np.float64)
    assert len(v) % 3 == 0 and len(v) > 0
    q = np.round(v / 1e-4).astype(np.int64)                     # weld: the same edge is interpolated from either end by neighbouring cubes
    _, ids = np.unique(q, axis=0, return_inverse=True)
    ids = ids.reshape(-1, 3)
    good = (ids[:, 0] != ids[:, 1]) & (ids[:, 1] != ids[:, 2]) & (ids[:, 0] != ids[:, 2])    # drop triangles collapsed by the weld
    directed = Counter()
    for a, b, c in ids[good].tolist():
        for e in ((a, b), (b, c), (c, a)):
            directed[e] += 1
    for (a, b), n in directed.items():
        assert n == 1 and directed.get((b, a), 0) == 1, "open or inconsistently wound edge"
    tri = v.reshape(-1, 3, 3)
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    out = centre_fn(tri.mean(axis=1))
    area = np.linalg.norm(nrm, axis=1)
    big = area > 1e-3 * area.max()
    outward = (nrm[big] * out[big]).sum(axis=1) > 0
    assert area[big][outward].sum() >= min_area_frac * area[big].sum(), "triangles wound towards the inside"


def test_sphere_mesh_is_closed_wound_outwards_and_on_the_surface():
    r, vs, vps = 0.62, 0.1, 8
    exp = mr.sdf_export(lambda x, y, z: np.sqrt(x * x + y * y + z * z) - r, vs, vps, -1, 1)
    mesh = mr.extract(exp, vs, vps)
    assert mesh["block_first"][-1] == len(mesh["vertices"]) and len(mesh["block_first"]) == 9
    rad = np.linalg.norm(mesh["vertices"].astype(np.float64), axis=1)
    assert np.abs(rad - r).max() < 0.15 * vs                    # linear interpolation of an exact distance field
    _closed_and_outward(mesh, lambda c: c)
    # every vertex lies in an observed voxel: colour and label are that voxel's
    assert np.all(mesh["rgba"][:, 3] == 255) and np.all(mesh["labels"] > 0)


def test_two_touching_blobs_exercise_the_ambiguous_faces():
    # two spheres on a cube diagonal: the cubes between them see alternating corner signs
    vs, vps = 0.1, 8
    c1, c2 = np.array([-0.2, -0.2, -0.2]), np.array([0.2, 0.2, 0.2])
    fn = lambda x, y, z: np.minimum(np.sqrt((x - c1[0]) ** 2 + (y - c1[1]) ** 2 + (z - c1[2]) ** 2),
                                    np.sqrt((x - c2[0]) ** 2 + (y - c2[1]) ** 2 + (z - c2[2]) ** 2)) - 0.33
    exp = mr.sdf_export(fn, vs, vps, -1, 1)
    mesh = mr.extract(exp, vs, vps)

    def outward(c):
        d1, d2 = c - c1, c - c2
        return np.where((np.linalg.norm(d1, axis=1) < np.linalg.norm(d2, axis=1))[:, None], d1, d2)
    # consistent winding is checked exactly (every edge is shared by two triangles running in opposite directions); "outward" per triangle
    # only as an area fraction: in the saddle cubes between the blobs the fans of non-planar loops contain slivers facing sideways
    _closed_and_outward(mesh, outward, 0.98)


def test_unobserved_or_missing_neighbours_produce_no_triangles():
    vs, vps = 0.1, 8
    exp = mr.sdf_export(lambda x, y, z: z - 0.03, vs, vps, -1, 1)            # a horizontal plane through all blocks
    full = mr.extract(exp, vs, vps)
    # z = 0.03 lies between the voxel centres -0.05 and 0.05: cubes of the lower blocks' top layer, which need the upper blocks
    keep = exp["block_index"][:, 2] < 0
    lower = {k: v[keep] for k, v in exp.items()}
    assert len(mr.extract(lower, vs, vps)["vertices"]) == 0
    # cubes along the outer +x / +y faces have no neighbour block: 15 x 15 cubes of the 16 x 16 columns remain
    assert len(full["vertices"]) == 15 * 15 * 2 * 3
    exp["tsdf_weight"][:, :] = 1e-4                                          # weight <= min_weight: unobserved
    assert len(mr.extract(exp, vs, vps)["vertices"]) == 0
