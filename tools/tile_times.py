"""Development aid: per-tile (records, cycles) of the tile-apply kernel for the last frame of a short run."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import WORKLOADS, make_cfg, gen_frames
from kimera_semantics_b200.capi import Integrator, _ptr
wl = sys.argv[1]; n = int(sys.argv[2])
itype, w, h, vs, Cc, _, _ = WORKLOADS[wl]
cam, frames = gen_frames(wl, n)
integ = Integrator(make_cfg(wl))
integ.lib.ksg_debug_tile_times(integ.handle, 1, 0, None)
for i in range(n):
    st = integ.integrate_depth(frames[i][2], frames[i][0], frames[i][1], cam.K)
nt = integ.lib.ksg_debug_tile_times(integ.handle, 1, 0, None)
buf = np.zeros((nt, 2), np.int64)
integ.lib.ksg_debug_tile_times(integ.handle, 1, nt, _ptr(buf, C.c_int64))
order = np.argsort(-buf[:, 1])
print("tiles", nt, "records", buf[:, 0].sum(), "sum cycles", buf[:, 1].sum(), "max cycles", buf[:, 1].max())
for k in order[:12]:
    print(f"  tile {k:5d}: records {buf[k,0]:8d} cycles {buf[k,1]:10d}  cycles/record {buf[k,1]/max(1,buf[k,0]):8.1f}")
small = buf[buf[:, 0] < 2000]
print("small tiles: n", len(small), "mean cycles", small[:, 1].mean() if len(small) else 0, "mean cycles/record", (small[:, 1] / np.maximum(1, small[:, 0])).mean() if len(small) else 0)
