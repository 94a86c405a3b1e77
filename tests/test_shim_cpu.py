"""The C++ host shim (kimera_semantics_b200/cpp): builds without CUDA / Eigen / glog, mirrors the reference's factory
error convention (abort with a message) and never integrates on the CPU."""
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "kimera_semantics_b200", "cpp")


@pytest.fixture(scope="module")
def demo():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "kimera_semantics_b200", "csrc"), "libksg.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", CPP], stdout=subprocess.DEVNULL)
    return os.path.join(CPP, "shim_demo")


def write_frames(path, frames_xyz_rgba, voxel_size, vps, palette, dynamic):
    import numpy as np
    with open(path, "wb") as f:
        f.write(np.int32(len(frames_xyz_rgba)).tobytes())
        f.write(np.float32(voxel_size).tobytes())
        f.write(np.int32(vps).tobytes())
        f.write(np.int32(len(palette)).tobytes())
        for l, c in enumerate(palette):
            f.write(bytes([int(c[0]), int(c[1]), int(c[2]), int(c[3]), l]))
        f.write(np.int32(len(dynamic)).tobytes())
        f.write(bytes(dynamic))
        for T, xyz, rgba in frames_xyz_rgba:
            f.write(np.int32(len(xyz)).tobytes())
            f.write(np.ascontiguousarray(T, np.float32).tobytes())
            f.write(np.ascontiguousarray(xyz, np.float32).tobytes())
            f.write(np.ascontiguousarray(rgba, np.uint8).tobytes())


def test_unknown_integrator_type_aborts_like_the_reference(demo, tmp_path):
    fr = tmp_path / "f.bin"
    write_frames(fr, [], 0.1, 16, [(255, 255, 255, 255)], [])
    r = subprocess.run([demo, "bogus", str(fr), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode != 0
    assert "Unknown TSDF integrator type: bogus" in r.stderr       # LOG(FATAL) factory.cpp:61


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_shim_has_no_cpu_fallback(demo, tmp_path):
    fr = tmp_path / "f.bin"
    write_frames(fr, [], 0.1, 16, [(255, 255, 255, 255)], [])
    r = subprocess.run([demo, "fast", str(fr), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode != 0
    assert "ksg_create failed" in r.stderr and "no CPU fallback" in r.stderr


def test_shim_library_has_no_cuda_or_oracle_dependency():
    out = subprocess.run(["ldd", os.path.join(CPP, "libkimera_semantics_gpu.so")], capture_output=True, text=True).stdout
    assert "libksg.so" in out and "oracle" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", os.path.join(CPP, "libkimera_semantics_gpu.so")], capture_output=True, text=True).stdout
    for name in ("SemanticTsdfIntegratorFactory6create", "FastSemanticTsdfIntegrator19integratePointCloud",
                 "MergedSemanticTsdfIntegrator19integratePointCloud", "SemanticLabel2Color25getSemanticLabelFromColor"):
        assert name in syms, name
