"""Development aid: where does a frame integrated into an EMPTY map (the delta map of the frame-per-GPU mode) spend its time?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import WORKLOADS, make_cfg, gen_frames
from kimera_semantics_b200.capi import Integrator
wl = sys.argv[1] if len(sys.argv) > 1 else "fast5"
itype, w, h, vs, C, _, _ = WORKLOADS[wl]
cam, frames = gen_frames(wl, 12)
dd = [torch.from_numpy(f[0]).cuda() for f in frames]
dl = [torch.from_numpy(f[1]).cuda() for f in frames]
integ = Integrator(make_cfg(wl))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); s = ts.cuda_stream
def ev(): return torch.cuda.Event(enable_timing=True)
for mode in ("reset every frame", "continuing map"):
    integ.reset()
    for i in range(12):
        t0 = time.perf_counter()
        if mode.startswith("reset"):
            integ.reset()
        t1 = time.perf_counter()
        a, b = ev(), ev()
        a.record(ts)
        integ.integrate_depth_device(frames[i][2], dd[i].data_ptr(), dl[i].data_ptr(), w, h, cam.K, s)
        b.record(ts)
        t2 = time.perf_counter()
        nb = integ.device_map_view()[0]
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"{mode:18s} frame {i:2d}: reset {1e3*(t1-t0):7.3f} ms host | integrate {a.elapsed_time(b):7.3f} ms device, {1e3*(t2-t1):7.3f} ms host | view {1e3*(t3-t2):7.3f} ms host | blocks {nb}", flush=True)
integ.reset()
integ.set_profiling(True)
for i in range(6):
    integ.reset(); integ.set_profiling(True)
    integ.integrate_depth_device(frames[i][2], dd[i].data_ptr(), dl[i].data_ptr(), w, h, cam.K, s)
    integ.sync()
    print("profile (into empty map):", {k: round(v, 4) if isinstance(v, float) else v for k, v in integ.get_profile().items() if not isinstance(v, (list, dict))}, flush=True)
print("timeline:", integ.fast_timeline() if hasattr(integ, "fast_timeline") else None)
