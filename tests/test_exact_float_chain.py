"""tools/exact_float_chain.py: the scan formulation of a long same-sign float32 addition chain must be bit-identical to the
sequential loop (the per-voxel, per-class log-probability recurrence of `merged`'s hot voxels)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import exact_float_chain as xc  # noqa: E402

F = np.float32
LM, LN = F(np.log(F(0.9))), F(np.log(F(1.0) - F(0.9)))


def bits(x):
    return int(np.array(x, np.float32).view(np.uint32))


def realistic_terms(rng, n):
    """L * freq entries: small non-negative counts times log(p) / log(1-p), summed in float32, plus zeros (label 0 column)."""
    counts = rng.integers(0, 4, (n, 3))
    t = (F(counts[:, 0]) * LM + F(counts[:, 1]) * LN).astype(np.float32) + (F(counts[:, 2]) * LN).astype(np.float32)
    t[rng.random(n) < 0.1] = 0.0
    return t.astype(np.float32)


@pytest.mark.parametrize("seed", range(8))
def test_scan_equals_sequential_on_realistic_chains(seed):
    rng = np.random.default_rng(seed)
    terms = realistic_terms(rng, 20000)
    s0 = F(-0.60205999132)
    assert bits(xc.scan_sum(s0, terms)) == bits(xc.sequential(s0, terms))


@pytest.mark.parametrize("seed", range(6))
def test_scan_equals_sequential_on_adversarial_chains(seed):
    """Many exact ties (terms whose low bits are exactly half a grid step), terms of wildly different magnitude, terms larger than
    the running value, subnormals, zeros, and chains that cross many binades."""
    rng = np.random.default_rng(100 + seed)
    n = 6000
    mant = rng.integers(1 << 23, 1 << 24, n)
    exps = rng.integers(-30, 6, n)
    terms = -np.ldexp(mant.astype(np.float64), exps - 23).astype(np.float32)
    tie = rng.random(n) < 0.3                          # force ...1000 patterns in the low bits -> ties on coarser grids
    low = rng.integers(1, 12, n)
    tm = (mant >> low << low) | (1 << (low - 1))
    terms[tie] = -np.ldexp(tm[tie].astype(np.float64), exps[tie] - 23).astype(np.float32)
    terms[rng.random(n) < 0.05] = 0.0
    terms[rng.random(n) < 0.01] = -F(1e-42)            # subnormal terms
    for s0 in (F(-1e-3), F(-0.60205999132), F(-777.25), F(-3.0e7)):
        assert bits(xc.scan_sum(s0, terms)) == bits(xc.sequential(s0, terms)), float(s0)


def test_block_size_does_not_matter_and_long_run_of_one_term():
    terms = np.full(50000, LN, np.float32)             # 50 000 identical updates: the hot free-space voxel
    want = xc.sequential(F(-0.60205999132), terms)
    for block in (1, 7, 32, 1024):
        assert bits(xc.scan_sum(F(-0.60205999132), terms, block=block)) == bits(want)
    assert float(want) < -1.0e5


@pytest.mark.parametrize("seed", range(4))
def test_two_level_chunked_version_is_exact_and_rarely_falls_back(seed):
    rng = np.random.default_rng(50 + seed)
    terms = realistic_terms(rng, 92000)                 # one frame of the hottest `merged2` voxel
    for s0 in (F(-0.60205999132), F(-2.5e5)):           # first frame / steady state
        got, fallbacks = xc.chunked_sum(s0, terms, chunk=1024)
        assert bits(got) == bits(xc.sequential(s0, terms))
        assert fallbacks <= 20                          # of 90 chunks: only the ones in which |s| doubles


def test_host_device_primitives_of_the_future_kernel_are_exact():
    """kimera_semantics_b200/csrc/ksg_chain.cuh (__host__ __device__ integer code: record tables, composition, binade crossing) compiled
    for the CPU and checked against the sequential float loop on 92 000-record realistic chains and tie-heavy adversarial ones."""
    import subprocess
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kimera_semantics_b200", "csrc")
    subprocess.check_call(["make", "-C", csrc, "-s", "test/chain_host_test"])
    out = subprocess.run([os.path.join(csrc, "test", "chain_host_test")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "chain ok" in out.stdout, out.stdout[-2000:]
