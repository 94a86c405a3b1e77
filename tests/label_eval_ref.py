"""CPU twin (numpy float32, same operation order) of ksg_evaluate_labels / csrc/ksg_eval.cuh: ground-truth label accuracy of an exported
map against an analytic world, after SemanticSimulationWorld::generateSemanticSdfFromWorld
(kimera_semantics/src/simulation/semantic_simulation_world.cpp:35-97).  Test infrastructure."""
import numpy as np

WORLD_DTYPE = np.dtype([("type", np.int32), ("a", np.float32, 3), ("b", np.float32, 3), ("label", np.int32)])
F = np.float32


def synthetic_scene_world() -> np.ndarray:
    """The benchmark scene of kimera_semantics_b200/synth.py (the reference's simulation world of semantic_simulation_eval.cpp:16-34 inside
    a 12 x 12 x 5 m room), object label = the object id the generator uses."""
    o = [(0, (0, 0, 2), (2, 0, 0), 1),              # sphere
         (1, (-2, -4, 2), (0, 1, 0), 2),            # plane y = -4
         (1, (4, 0, 0), (-1, 0, 0), 3),             # plane x = 4
         (2, (-4, 4, 2), (4, 4, 4), 4),             # cube
         (1, (0, 0, 0.03), (0, 0, 1), 5),           # ground
         (1, (-6, 0, 0), (1, 0, 0), 6), (1, (0, 6, 0), (0, -1, 0), 6), (1, (0, 0, 5), (0, 0, -1), 6)]   # visible faces of the room
    w = np.zeros(len(o), WORLD_DTYPE)
    for i, (t, a, b, l) in enumerate(o):
        w[i] = (t, a, b, l)
    return w


def _dist(o, x, y, z):
    a, b = o["a"], o["b"]
    if o["type"] == 0:
        dx, dy, dz = x - a[0], y - a[1], z - a[2]
        return np.sqrt((dx * dx + dy * dy) + dz * dz).astype(F) - b[0]
    if o["type"] == 1:
        return ((b[0] * (x - a[0]) + b[1] * (y - a[1])) + b[2] * (z - a[2])).astype(F)
    p = (x, y, z)
    dv = [np.maximum(np.maximum((a[k] - b[k] / F(2)) - p[k], F(0)), (p[k] - a[k]) - b[k] / F(2)).astype(F) for k in range(3)]
    d = np.sqrt((dv[0] * dv[0] + dv[1] * dv[1]) + dv[2] * dv[2]).astype(F)
    iv = [np.maximum((a[k] - b[k] / F(2)) - p[k], (p[k] - a[k]) - b[k] / F(2)).astype(F) for k in range(3)]
    inside = np.maximum(iv[0], np.maximum(iv[1], iv[2]))
    return np.where(d < F(1e-6), inside, d).astype(F)


def evaluate(exp, world, voxel_size, vps, num_labels, max_dist, band, checker_size=0.0, checker_margin=0.0):
    """(evaluated, correct, observed) over an export dict (block_index [nb,3], tsdf_distance / tsdf_weight / sem_label [nb, vps^3])."""
    V = vps ** 3
    lin = np.arange(V)
    lx, ly, lz = lin % vps, (lin // vps) % vps, lin // (vps * vps)
    bi = exp["block_index"].astype(np.int64)
    gx = (bi[:, 0:1] * vps + lx[None, :]).astype(F)
    gy = (bi[:, 1:2] * vps + ly[None, :]).astype(F)
    gz = (bi[:, 2:3] * vps + lz[None, :]).astype(F)
    vs = F(voxel_size)
    cx, cy, cz = ((gx + F(0.5)) * vs).astype(F), ((gy + F(0.5)) * vs).astype(F), ((gz + F(0.5)) * vs).astype(F)
    obs = exp["tsdf_weight"] > 0
    sel = obs & (np.abs(exp["tsdf_distance"]) <= F(band))
    best = np.full(cx.shape, F(max_dist), F)
    gt = np.zeros(cx.shape, np.int64)
    anyo = np.zeros(cx.shape, bool)
    for o in world:
        d = _dist(o, cx, cy, cz)
        closer = d < best
        best = np.where(closer, d, best)
        gt = np.where(closer, int(o["label"]), gt)
        anyo |= closer
    sel &= anyo
    if checker_size > 0:
        s = F(checker_size)
        cell = np.zeros(cx.shape, np.int64)
        near = np.zeros(cx.shape, bool)
        for c in (cx, cy, cz):
            q = (c / s).astype(F)
            fl = np.floor(q)
            cell += fl.astype(np.int64)
            fr = ((q - fl).astype(F) * s).astype(F)
            near |= (fr < F(checker_margin)) | ((s - fr).astype(F) < F(checker_margin))
        sel &= ~near
        gt = 1 + ((cell + gt) % (num_labels - 1))
    return int(sel.sum()), int((sel & (exp["sem_label"].astype(np.int64) == gt)).sum()), int(obs.sum())
