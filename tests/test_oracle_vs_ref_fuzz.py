"""Randomised differential test: the oracle against the reference's own sources (oracle/_ref) on random clouds, poses and
configuration switches - far/near/behind-the-camera points, |z| ~ 0 (zero weight), unknown colours, freespace clouds, every
Config flag.  Bit-exact in every exported field; `merged` in the oracle's faithful (libstdc++ bundle order) mode.
The same generator (tests/fuzz_cases.py) drives the CUDA path in tests/test_gpu_fuzz.py."""
import pytest

from oracle import ref_py
from oracle.oracle_py import OracleIntegrator
from parity_utils import compare_maps
import fuzz_cases

pytestmark = pytest.mark.skipif(not ref_py.available(), reason="oracle/_ref/libks_ref_hybrid.so not built (needs /root/reference)")


@pytest.mark.parametrize("seed", range(24))
def test_random_case_oracle_equals_reference_sources(seed):
    cfg, frames = fuzz_cases.make_case(seed)
    ora = OracleIntegrator(cfg, canonical_merged=False)
    ora.set_color_to_label(*fuzz_cases.color_table(cfg))
    ref = ref_py.RefHybridIntegrator(cfg)
    with fuzz_cases.quiet_stderr():
        for T, pts, rgba, freespace in frames:
            ora.integrate_points(T, pts, rgba=rgba, freespace=freespace)
            ref.integrate_points(T, pts, rgba=rgba, freespace=freespace)
    rep = compare_maps(ref.export(), ora.export())
    assert rep["same_blocks"] == 1.0, rep
    assert not {k: v for k, v in rep.items() if k.endswith("mismatch") and v}, rep
