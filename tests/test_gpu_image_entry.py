"""ksg_integrate_image: the depth / semantic encodings of the reference's front end (kimera_semantics_ros/include/kimera_semantics_ros/
depth_map_to_pointcloud.h:183-193,213-266): uint16 millimetre depth (DepthTraits<uint16_t>) and an RGB semantic image whose colours name
the labels.  The oracle gets the cloud the reference's PointCloudFromDepth::convert<uint16_t> would hand to integratePointCloud (numpy
float32 in the reference's operation order) through its points entry; the CUDA path gets the raw images."""
import numpy as np
import pytest

from kimera_semantics_b200 import synth
from kimera_semantics_b200.capi import Integrator, KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED, KSG_COLOR_MODE_COLOR
from oracle.oracle_py import OracleIntegrator
from parity_utils import assert_parity, compare_maps, make_config

pytestmark = pytest.mark.gpu


def reference_cloud_u16(depth_mm: np.ndarray, K64):
    """PointCloudFromDepth::convert<uint16_t> (depth_map_to_pointcloud.h:222-266) + the finite filter of the cloud conversion."""
    h, w = depth_mm.shape
    fx, fy, cx, cy = [float(x) for x in K64]
    center_x, center_y = np.float32(cx), np.float32(cy)
    unit = float(np.float32(0.001))                       # double unit_scaling = DepthTraits<uint16_t>::toMeters(1) = 1 * 0.001f
    constant_x, constant_y = np.float32(unit / fx), np.float32(unit / fy)
    v, u = np.mgrid[0:h, 0:w]
    d = depth_mm.astype(np.float32)
    x = ((u.astype(np.float32) - center_x) * d).astype(np.float32) * constant_x
    y = ((v.astype(np.float32) - center_y) * d).astype(np.float32) * constant_y
    z = d * np.float32(0.001)
    ok = depth_mm.reshape(-1) != 0
    xyz = np.stack([x.reshape(-1), y.reshape(-1), z.reshape(-1)], axis=1).astype(np.float32)
    return np.ascontiguousarray(xyz[ok]), np.nonzero(ok)[0]


@pytest.mark.parametrize("itype,color_mode", [(KSG_INTEGRATOR_FAST, 1), (KSG_INTEGRATOR_FAST, KSG_COLOR_MODE_COLOR), (KSG_INTEGRATOR_MERGED, 1)])
def test_uint16_depth_and_rgb_semantic_image_entry_matches_the_reference_front_end(itype, color_mode):
    W, H, C = 320, 240, 21
    cam = synth.make_camera(W, H)
    cfg = make_config(itype, 0.05, C, max_points=W * H, max_updates=16 << 20, color_mode=color_mode)
    pal = np.array([[cfg.label_color[l][k] for k in range(4)] for l in range(256)], np.uint8)
    table_rgb, table_lab = pal[:C, :3].copy(), np.arange(C, dtype=np.uint8)
    gpu, ora = Integrator(cfg), OracleIntegrator(cfg)
    gpu.set_color_to_label(table_rgb, table_lab)
    ora.set_color_to_label(table_rgb, table_lab)
    K64 = np.array([float(cam.K[0]), float(cam.K[1]), float(cam.K[2]), float(cam.K[3])], np.float64)
    for f in range(3):
        depth, label, T = synth.frame(cam, f, C)
        mm = np.clip(np.rint(depth.astype(np.float64) * 1000.0), 0, 65535).astype(np.uint16)
        mm.reshape(-1)[(f * 7)::53] = 0                       # invalid measurements
        rgb = np.ascontiguousarray(pal[label][:, :, :3])
        rgb.reshape(-1, 3)[(f * 5)::97] = (1, 2, 3)           # colours that are not in the table -> label 0 (color.cpp:75-80)
        sg = gpu.integrate_image(T, mm, rgb, K64)
        xyz, pix = reference_cloud_u16(mm, K64)
        rgba = np.concatenate([rgb.reshape(-1, 3)[pix], np.full((len(pix), 1), 255, np.uint8)], axis=1)
        so = ora.integrate_points(T, xyz, rgba=np.ascontiguousarray(rgba))
        assert sg.points_in == so.points_in and sg.voxel_updates == so.voxel_updates and sg.rays_cast == so.rays_cast, (sg.as_dict(), so.as_dict())
    assert_parity(compare_maps(gpu.export(), ora.export()))
    gpu.close(); ora.close()
