"""Helper run in its OWN process by tests/test_gpu_bundle_order.py: `merged` with ksg_config.merged_bundle_order =
KSG_BUNDLE_ORDER_LIBSTDCXX through the C-ABI, compared field by field with the digests produced by the reference's own sources
(tests/golden/ref_hybrid_golden.json).  Prints one JSON object {case: {field: equal?}}.  A separate process keeps a CUDA fault in
this young mode from poisoning the context of the rest of the GPU suite."""
import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
_spec = importlib.util.spec_from_file_location("make_ref_golden", os.path.join(HERE, "golden", "make_ref_golden.py"))
mrg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mrg)


def main(make_integrator=None):
    from kimera_semantics_b200.capi import KSG_BUNDLE_ORDER_LIBSTDCXX, KSG_INTEGRATOR_MERGED
    if make_integrator is None:
        from kimera_semantics_b200.capi import Integrator as make_integrator
    golden = json.load(open(os.path.join(HERE, "golden", "ref_hybrid_golden.json")))
    report = {}
    for name in sorted(mrg.CASES):
        if mrg.CASES[name][0] != KSG_INTEGRATOR_MERGED:
            continue

        def make(cfg):
            cfg.merged_bundle_order = KSG_BUNDLE_ORDER_LIBSTDCXX
            return make_integrator(cfg)
        got = mrg.digest(mrg.run_case(name, make))
        report[name] = {k: got[k] == golden[name][k] for k in mrg.KEYS if not (k == "tsdf_rgba" and "probability" in name)}
    print("REPORT " + json.dumps(report), flush=True)
    return report


if __name__ == "__main__":
    main()
