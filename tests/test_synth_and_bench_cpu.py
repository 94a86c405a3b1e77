"""Synthetic generator determinism, the product/oracle import boundary, and the N > 1 plumbing of bench.py on gloo."""
import ast
import os
import subprocess
import sys

import numpy as np
import pytest

from kimera_semantics_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frames_are_deterministic_and_cover_the_label_set():
    cam = synth.make_camera(160, 120)
    d1, l1, T1 = synth.frame(cam, 3, 21)
    d2, l2, T2 = synth.frame(cam, 3, 21)
    assert np.array_equal(d1, d2) and np.array_equal(l1, l2) and np.array_equal(T1, T2)
    assert d1.dtype == np.float32 and l1.dtype == np.uint8 and T1.shape == (7,)
    assert abs(np.linalg.norm(T1[:4]) - 1) < 1e-6
    assert l1.max() < 21 and len(np.unique(l1)) == 21
    assert d1.min() > 0.3 and d1.max() > 5.0          # near sphere and walls beyond max_ray_length (clearing rays)
    d3, _, _ = synth.frame(cam, 4, 21)
    assert not np.array_equal(d1, d3)


def test_backprojection_matches_float32_reference_formula():
    cam = synth.make_camera(64, 48)
    depth, _, _ = synth.frame(cam, 0, 5, invalid_fraction=0.2)
    xyz, pix = synth.backproject(depth, cam)
    assert len(pix) == np.isfinite(depth).sum() and len(pix) < depth.size
    K = cam.K
    u, v = (pix % 64).astype(np.float32), (pix // 64).astype(np.float32)
    d = depth.reshape(-1)[pix]
    x = ((u - K[2]) * d) * np.float32(1.0 / np.float64(K[0]))
    assert np.array_equal(xyz[:, 0], x) and np.array_equal(xyz[:, 2], d)


def test_product_package_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/ (the checker)."""
    pkg = os.path.join(ROOT, "kimera_semantics_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            path = os.path.join(dirpath, fn)
            if fn.endswith(".py"):
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        names = [node.module or ""]
                    assert not any(n.split(".")[0] == "oracle" for n in names), path
            elif fn.endswith((".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(path, errors="ignore").read()
                assert "ks_oracle" not in text and "oracle/" not in text and "kso_" not in text, path


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import bench
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
# every rank integrates ITS OWN sequence (the way bench.py shards: one stream + map per rank, no data-path collective)
cam, frames = bench.gen_frames("fast10", 2, rank)
from oracle.oracle_py import OracleIntegrator   # CPU stand-in for the device in this no-GPU test of the plumbing
integ = OracleIntegrator(bench.make_cfg("fast10"))
upd = sum(integ.integrate_depth(T, d, l, cam.K).voxel_updates for d, l, T in frames)
ms = 10.0 * (rank + 1)
t = torch.tensor([ms, float(upd)], dtype=torch.float64)
tmax, tsum = t.clone(), t.clone()
dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
if rank == 0:
    print("RESULT", float(tmax[0]), int(tsum[1]), upd, flush=True)
digest = float(frames[0][0].sum())
allv = [None] * world
dist.all_gather_object(allv, digest)
if rank == 0:
    print("DISTINCT", len(set(allv)), flush=True)
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharding_and_max_over_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0].split()
    assert float(line[1]) == 20.0                      # max over ranks of the per-rank time
    assert int(line[2]) > int(line[3]) > 0             # whole-job updates = sum over ranks
    assert "DISTINCT 2" in outs[0][0]                  # the ranks really integrate different streams


def test_reference_impl_line_shape():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "fast10", "--steps", "3",
                          "--warmup", "3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0   # 5 labels: only the port can run it


def test_reference_impl_times_the_reference_sources_when_built():
    """At the reference's 21 labels both CPU arms are calibrated and the faster (arm, threads) pair is the reported one."""
    from oracle import ref_py
    if not ref_py.available(fast_build=True):
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "fast5", "--steps", "3",
                          "--warmup", "3"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    cal = line["cpu_baseline"]["calibration_fps"]
    assert any(k.startswith("reference@") for k in cal) and any(k.startswith("port@") for k in cal)
    best = max(cal, key=cal.get)
    assert best == f'{line["cpu_baseline"]["kind"]}@{line["cpu_baseline"]["cores"]}'
    assert line["value"] > 0 and line["mvoxel_updates_per_s"] > 0
