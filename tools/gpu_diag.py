"""Development aid: run a few configurations on the GPU next to the oracle and print full reports
(does not stop at the first difference)."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from kimera_semantics_b200 import synth
from kimera_semantics_b200.capi import Integrator, KSG_INTEGRATOR_FAST, KSG_INTEGRATOR_MERGED, load_library
from oracle.oracle_py import OracleIntegrator
from parity_utils import compare_maps, make_config, frames

print(load_library().ksg_build_info().decode())
CASES = [
    ("fast  320x240 10cm C5 tma", KSG_INTEGRATOR_FAST, 320, 240, 0.10, 5, 3, {}),
    ("fast  320x240 10cm C5 coop", KSG_INTEGRATOR_FAST, 320, 240, 0.10, 5, 3, {"apply_mode": 1}),
    ("merged 320x240 10cm C5 tma", KSG_INTEGRATOR_MERGED, 320, 240, 0.10, 5, 2, {"max_updates": 8 << 20}),
    ("merged 320x240 10cm C5 coop", KSG_INTEGRATOR_MERGED, 320, 240, 0.10, 5, 2, {"max_updates": 8 << 20, "apply_mode": 1}),
    ("fast  640x480 5cm C21", KSG_INTEGRATOR_FAST, 640, 480, 0.05, 21, 4, {}),
    ("merged 640x480 5cm C21", KSG_INTEGRATOR_MERGED, 640, 480, 0.05, 21, 2, {"max_updates": 16 << 20}),
]
for name, itype, w, h, vs, C, nf, kw in CASES:
    try:
        cfg = make_config(itype, vs, C, max_points=w * h, **kw)
        gpu, ora = Integrator(cfg), OracleIntegrator(cfg)
        for i, (cam, depth, label, T) in enumerate(frames(w, h, C, nf)):
            t0 = time.time(); sg = gpu.integrate_depth(T, depth, label, cam.K); tg = time.time() - t0
            t0 = time.time(); so = ora.integrate_depth(T, depth, label, cam.K); to = time.time() - t0
            rep = compare_maps(gpu.export(), ora.export())
            print(f"[{name}] frame {i}: gpu {tg*1e3:.2f} ms oracle {to*1e3:.1f} ms\n   gpu    {sg.as_dict()}\n   oracle {so.as_dict()}\n   {rep}", flush=True)
        gpu.close()
    except Exception:
        traceback.print_exc()
