"""CPU twin (numpy float32, same operation order) of ksg_extract_mesh / csrc/ksg_mesh.cuh: marching cubes over an exported map with the
vertex colour / label taken from the voxel that contains the vertex (SURVEY.md 8f NEXT-4; restates voxblox's MeshIntegrator, which is not
under /root/reference - see the header of csrc/ksg_mesh.cuh).  Test infrastructure."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import make_mc_table as mc  # noqa: E402

F = np.float32
CORNER = np.array(mc.CORNERS, np.int64)          # [8, 3] (x, y, z) offsets
EDGE = np.array(mc.EDGES, np.int64)              # [12, 2]
_T = mc.build_table()
NTRI = np.array([len(r) // 3 for r in _T], np.int64)
TABLE = np.full((256, 15), -1, np.int64)
for _i, _r in enumerate(_T):
    TABLE[_i, :len(_r)] = _r


def extract(exp, voxel_size, vps, min_weight=1e-4):
    """exp: export dict (block_index [nb, 3] in (z, y, x) order, tsdf_distance / tsdf_weight / sem_label [nb, vps^3], tsdf_rgba [nb, vps^3, 4]).
    -> dict(vertices [n, 3] f32, rgba [n, 4] u8, labels [n] u8, block_first [nb + 1] i64), same layout as Integrator.extract_mesh()."""
    vs = F(voxel_size)
    vsi = F(1.0 / float(vs))              # DevCfg.vsi
    eps = F(1e-6)
    mw = F(min_weight)
    bidx = exp["block_index"].astype(np.int64)
    nb = len(bidx)
    where = {tuple(b): i for i, b in enumerate(bidx.tolist())}
    D = exp["tsdf_distance"].reshape(nb, vps, vps, vps)      # [z, y, x]
    W = exp["tsdf_weight"].reshape(nb, vps, vps, vps)
    RGBA = exp["tsdf_rgba"].reshape(nb, vps ** 3, 4)
    LAB = exp["sem_label"].reshape(nb, vps ** 3)
    block_size = F(vps) * vs
    verts, cols, labs, first = [], [], [], [0]
    lz, ly, lx = np.meshgrid(np.arange(vps), np.arange(vps), np.arange(vps), indexing="ij")
    for i in range(nb):
        b = bidx[i]
        d = np.zeros((vps + 1,) * 3, F)
        w = np.zeros((vps + 1,) * 3, F)
        have = np.zeros((vps + 1,) * 3, bool)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    j = where.get((b[0] + dx, b[1] + dy, b[2] + dz))
                    if j is None:
                        continue
                    sz = slice(0, vps) if dz == 0 else slice(vps, vps + 1)
                    sy = slice(0, vps) if dy == 0 else slice(vps, vps + 1)
                    sx = slice(0, vps) if dx == 0 else slice(vps, vps + 1)
                    tz = slice(0, vps) if dz == 0 else slice(0, 1)
                    ty = slice(0, vps) if dy == 0 else slice(0, 1)
                    tx = slice(0, vps) if dx == 0 else slice(0, 1)
                    d[sz, sy, sx] = D[j][tz, ty, tx]
                    w[sz, sy, sx] = W[j][tz, ty, tx]
                    have[sz, sy, sx] = True
        ok = np.ones((vps,) * 3, bool)
        idx = np.zeros((vps,) * 3, np.int64)
        cd = []
        for c in range(8):
            ox, oy, oz = CORNER[c]
            sl = (slice(oz, oz + vps), slice(oy, oy + vps), slice(ox, ox + vps))
            ok &= have[sl] & (w[sl] > mw)
            cd.append(d[sl])
            idx |= (d[sl] < 0).astype(np.int64) << c
        ntri = np.where(ok, NTRI[idx], 0)
        sel = np.nonzero(ntri.reshape(-1) > 0)[0]                 # linear voxel order x + vps * (y + vps * z)
        if len(sel) == 0:
            first.append(first[-1])
            continue
        cfg = idx.reshape(-1)[sel]
        nt = ntri.reshape(-1)[sel]
        sdf = np.stack([c.reshape(-1)[sel] for c in cd], axis=1)     # [m, 8]
        origin = (b.astype(F) * block_size).astype(F)
        l = np.stack([lx.reshape(-1)[sel], ly.reshape(-1)[sel], lz.reshape(-1)[sel]], axis=1)
        base = (origin[None, :] + ((l.astype(F) + F(0.5)) * vs).astype(F)).astype(F)       # [m, 3]
        e = TABLE[cfg]                                               # [m, 15]
        valid = np.arange(15)[None, :] < (3 * nt)[:, None]
        ee = np.where(valid, e, 0)
        a, bb = EDGE[ee, 0], EDGE[ee, 1]                             # [m, 15]
        rows = np.arange(len(sel))[:, None]
        sa, sb = sdf[rows, a], sdf[rows, bb]
        pa = (base[:, None, :] + (CORNER[a].astype(F) * vs).astype(F)).astype(F)
        pb = (base[:, None, :] + (CORNER[bb].astype(F) * vs).astype(F)).astype(F)
        diff = (sa - sb).astype(F)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (sa / diff).astype(F)
            p_int = (pa + (t[:, :, None] * (pb - pa).astype(F)).astype(F)).astype(F)
        p_mid = (F(0.5) * (pa + pb).astype(F)).astype(F)
        p = np.where((np.abs(diff) >= F(1e-6))[:, :, None], p_int, p_mid)
        p = p[valid]                                                 # cube-major, then table order
        # colour / label of the voxel containing the vertex
        g = np.floor((p * vsi).astype(F) + eps).astype(np.int64)
        gb = g // vps
        gl = g & (vps - 1)
        lin = gl[:, 0] + vps * (gl[:, 1] + vps * gl[:, 2])
        col = np.zeros((len(p), 4), np.uint8)
        lab = np.zeros(len(p), np.uint8)
        for k, key in enumerate(map(tuple, gb.tolist())):
            j = where.get(key)
            if j is not None and W[j].reshape(-1)[lin[k]] > mw:
                col[k] = RGBA[j][lin[k]]
                lab[k] = LAB[j][lin[k]]
        verts.append(p.astype(F))
        cols.append(col)
        labs.append(lab)
        first.append(first[-1] + len(p))
    return {"vertices": np.concatenate(verts) if verts else np.zeros((0, 3), F),
            "rgba": np.concatenate(cols) if cols else np.zeros((0, 4), np.uint8),
            "labels": np.concatenate(labs) if labs else np.zeros(0, np.uint8),
            "block_first": np.array(first, np.int64)}


def sdf_export(fn, voxel_size, vps, block_lo, block_hi, weight=1.0):
    """export dict of an analytic signed distance function sampled at the voxel centres of the blocks block_lo <= index < block_hi"""
    vs = F(voxel_size)
    blocks = [(x, y, z) for z in range(block_lo, block_hi) for y in range(block_lo, block_hi) for x in range(block_lo, block_hi)]
    V = vps ** 3
    lin = np.arange(V)
    lx, ly, lz = lin % vps, (lin // vps) % vps, lin // (vps * vps)
    exp = {"block_index": np.array(blocks, np.int32), "tsdf_distance": np.zeros((len(blocks), V), F),
           "tsdf_weight": np.full((len(blocks), V), weight, F), "tsdf_rgba": np.zeros((len(blocks), V, 4), np.uint8),
           "sem_label": np.zeros((len(blocks), V), np.uint8)}
    for i, (bx, by, bz) in enumerate(blocks):
        cx = ((bx * vps + lx).astype(F) + F(0.5)) * vs
        cy = ((by * vps + ly).astype(F) + F(0.5)) * vs
        cz = ((bz * vps + lz).astype(F) + F(0.5)) * vs
        exp["tsdf_distance"][i] = fn(cx, cy, cz).astype(F)
        exp["sem_label"][i] = 1 + (lin % 7)
        exp["tsdf_rgba"][i, :, 0] = lin % 251
        exp["tsdf_rgba"][i, :, 3] = 255
    return exp
