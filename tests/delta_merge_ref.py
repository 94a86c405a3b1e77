"""CPU twin (numpy float32, same operation order) of ksg_merge_blocks_device / csrc/ksg_merge.cuh: the frame-per-GPU batch mode merges
the map of one frame (integrated into an EMPTY map: the "delta") into the base map voxel by voxel.  Test infrastructure."""
from typing import Dict

import numpy as np

P_INIT = np.float32(-0.60205999132)


def _round_half_away(x: np.ndarray) -> np.ndarray:     # roundf for x >= 0
    r = np.floor(x)
    return r + ((x - r) >= np.float32(0.5)).astype(np.float32)


def _blend(c1: np.ndarray, w1: np.ndarray, c2: np.ndarray, w2: np.ndarray) -> np.ndarray:
    """voxblox Color::blendTwoColors on uint8 [..., 4], float32."""
    total = (w1 + w2).astype(np.float32)
    a = (w1 / total).astype(np.float32)[..., None]
    b = (w2 / total).astype(np.float32)[..., None]
    v = _round_half_away((c1.astype(np.float32) * a).astype(np.float32) + (c2.astype(np.float32) * b).astype(np.float32))
    return v.astype(np.uint8)


def empty_map(vps: int, C: int) -> Dict[str, np.ndarray]:
    V = vps ** 3
    return {"block_index": np.zeros((0, 3), np.int32), "tsdf_distance": np.zeros((0, V), np.float32), "tsdf_weight": np.zeros((0, V), np.float32),
            "tsdf_rgba": np.zeros((0, V, 4), np.uint8), "sem_label": np.zeros((0, V), np.uint8), "sem_priors": np.zeros((0, V, C), np.float32),
            "sem_rgba": np.zeros((0, V, 4), np.uint8)}


def merge(base: Dict[str, np.ndarray], delta: Dict[str, np.ndarray], label_rgba: np.ndarray, max_weight: float, color_mode: int) -> Dict[str, np.ndarray]:
    """Returns base with delta merged in; blocks sorted by (z, y, x) as ksg_export_blocks returns them."""
    V = delta["tsdf_distance"].shape[1]
    C = delta["sem_priors"].shape[2]
    idx = {tuple(b): i for i, b in enumerate(base["block_index"].tolist())}
    new = [tuple(b) for b in delta["block_index"].tolist() if tuple(b) not in idx]
    nb0, nn = len(idx), len(new)
    out = {k: v.copy() for k, v in base.items()}
    if nn:
        out["block_index"] = np.concatenate([out["block_index"], np.array(new, np.int32).reshape(-1, 3)])
        out["tsdf_distance"] = np.concatenate([out["tsdf_distance"], np.zeros((nn, V), np.float32)])
        out["tsdf_weight"] = np.concatenate([out["tsdf_weight"], np.zeros((nn, V), np.float32)])
        out["tsdf_rgba"] = np.concatenate([out["tsdf_rgba"], np.zeros((nn, V, 4), np.uint8)])
        out["sem_label"] = np.concatenate([out["sem_label"], np.zeros((nn, V), np.uint8)])
        out["sem_priors"] = np.concatenate([out["sem_priors"], np.full((nn, V, C), P_INIT, np.float32)])
        grey = np.tile(np.array([127, 127, 127, 255], np.uint8), (nn, V, 1))
        out["sem_rgba"] = np.concatenate([out["sem_rgba"], grey])
        for k, b in enumerate(new):
            idx[b] = nb0 + k
    rows = np.array([idx[tuple(b)] for b in delta["block_index"].tolist()], np.int64)
    wa = delta["tsdf_weight"]
    sem = (delta["sem_priors"].view(np.uint32) != P_INIT.view(np.uint32)).any(axis=-1)
    touched = (wa > 0) | sem
    wb = out["tsdf_weight"][rows]
    cw = (wa + wb).astype(np.float32)
    tsdf = (wa > 0) & (cw > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        nd = (((delta["tsdf_distance"] * wa).astype(np.float32) + (out["tsdf_distance"][rows] * wb).astype(np.float32)).astype(np.float32) / cw).astype(np.float32)
        blended = _blend(delta["tsdf_rgba"], wa, out["tsdf_rgba"][rows], wb)
    d_new = np.where(tsdf, nd, out["tsdf_distance"][rows])
    w_new = np.where(tsdf, np.minimum(cw, np.float32(max_weight)), wb)
    rgba = np.where(tsdf[..., None], blended, out["tsdf_rgba"][rows])
    p_new = (out["sem_priors"][rows] + (delta["sem_priors"] - P_INIT).astype(np.float32)).astype(np.float32)
    lab = p_new.argmax(axis=-1).astype(np.uint8)            # first maximum wins
    sc = label_rgba[lab]
    if color_mode == 1:
        rgba = sc
    elif color_mode == 2:
        raise NotImplementedError("kSemanticProbability uses expf: not part of the bit-exact merge check")
    out["tsdf_distance"][rows] = np.where(touched, d_new, out["tsdf_distance"][rows])
    out["tsdf_weight"][rows] = np.where(touched, w_new, out["tsdf_weight"][rows])
    out["tsdf_rgba"][rows] = np.where(touched[..., None], rgba, out["tsdf_rgba"][rows])
    out["sem_priors"][rows] = np.where(touched[..., None], p_new, out["sem_priors"][rows])
    out["sem_label"][rows] = np.where(touched, lab, out["sem_label"][rows])
    out["sem_rgba"][rows] = np.where(touched[..., None], sc, out["sem_rgba"][rows])
    order = np.lexsort((out["block_index"][:, 0], out["block_index"][:, 1], out["block_index"][:, 2]))
    return {k: v[order] for k, v in out.items()}
