"""Development aid for ncu: integrate a few frames of a workload through the device entry point."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import WORKLOADS, make_cfg, gen_frames
from kimera_semantics_b200.capi import Integrator
wl = sys.argv[1]; n = int(sys.argv[2])
itype, w, h, vs, C, _, _ = WORKLOADS[wl]
cam, frames = gen_frames(wl, n)
d_depth = [torch.from_numpy(f[0]).cuda() for f in frames]
d_label = [torch.from_numpy(f[1]).cuda() for f in frames]
integ = Integrator(make_cfg(wl))
stream = torch.cuda.current_stream().cuda_stream
for i in range(n):
    st = integ.integrate_depth_device(frames[i][2], d_depth[i].data_ptr(), d_label[i].data_ptr(), w, h, cam.K, stream, want_stats=True)
    print(i, st.as_dict(), flush=True)
