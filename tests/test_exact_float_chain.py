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


def test_index_level_model_of_the_hot_voxel_prepass():
    """Line-by-line Python model of csrc/ksg_hot.cuh (chunking with a ragged last chunk, float64 binade guesses, lane-wise folding of
    32-record sub-chunks in the skewed shared-memory layout, warp scan, table application with the plain-loop fallback) on a
    synthetic hot voxel: the finished row must equal the sequential sums for every class."""
    C, CH, STRIDE = 21, 1024, 1024 + 32 + 1
    rng = np.random.default_rng(9)
    n_bundles, n_rec = 12000, 4096 + 3 * 1024 + 517                     # ragged tail
    counts = rng.integers(0, 3, (n_bundles, C)).astype(np.float32)
    L = np.full((C, C), LN, np.float32); np.fill_diagonal(L, LM); L[:, 0] = 0.0
    tmp = np.zeros((n_bundles, C), np.float32)
    for j in range(C):                                                  # j ascending, one multiply and one add per term
        tmp = (tmp + (L[:, j][None, :] * counts[:, j][:, None]).astype(np.float32)).astype(np.float32)
    orders = np.sort(rng.choice(n_bundles, n_rec, replace=False))       # the voxel's records, in bundle order
    n_chunks = (n_rec + CH - 1) // CH
    for regime, prior0 in (("first frames", (-rng.uniform(0.6, 3.0e3, C)).astype(np.float32)),
                           ("steady state", (-rng.uniform(2.0e6, 4.0e6, C)).astype(np.float32))):
        fallbacks = _run_hot_model(C, CH, STRIDE, tmp, orders, prior0, n_chunks)
        if regime == "steady state":
            assert fallbacks <= C                   # a class crosses a binade at most once in these 8 chunks


def _run_hot_model(C, CH, STRIDE, tmp, orders, prior0, n_chunks):
    # k_hot_chunk_sums + k_hot_guess
    sums = np.zeros((n_chunks, C))
    for w in range(n_chunks):
        sums[w] = tmp[orders[w * CH:(w + 1) * CH]].astype(np.float64).sum(axis=0)
    guess = np.zeros((n_chunks, C), np.int64)
    for c in range(C):
        run = float(prior0[c])
        for w in range(n_chunks):
            guess[w, c] = xc._decompose(F(run))[1]
            run += sums[w, c]
    # k_hot_chunk_tables
    tables = {}
    for w in range(n_chunks):
        rows = orders[w * CH:(w + 1) * CH]
        s_cols = np.zeros(C * STRIDE, np.float32)
        for r in range(CH):
            v = tmp[rows[r]] if r < len(rows) else np.zeros(C, np.float32)
            for lane in range(C):
                s_cols[lane * STRIDE + r + (r >> 5)] = v[lane]
        for c in range(C):
            lane_tables = []
            for lane in range(32):
                col = c * STRIDE + lane * 33
                t = xc.record_table(s_cols[col], int(guess[w, c]))
                for k in range(1, 32):
                    t = xc.compose(t, xc.record_table(s_cols[col + k], int(guess[w, c])))
                lane_tables.append(t)
            off = 1
            while off < 32:                                             # Hillis-Steele inclusive scan
                lane_tables = [xc.compose(lane_tables[l - off], lane_tables[l]) if l >= off else lane_tables[l] for l in range(32)]
                off <<= 1
            tables[(w, c)] = lane_tables[31]
    # k_hot_apply
    fallbacks = 0
    for c in range(C):
        p = prior0[c]
        for w in range(n_chunks):
            m, g = xc._decompose(p)
            inc = tables[(w, c)][m & 1][0]
            if p < 0 and m >= 0x800000 and g == guess[w, c] and m + inc < (1 << 24):
                p = F(-np.ldexp(float(m + inc), g))
            else:
                for o in orders[w * CH:(w + 1) * CH]:
                    p = F(p + tmp[o, c])
                fallbacks += 1
        assert bits(p) == bits(xc.sequential(prior0[c], tmp[orders, c])), c
    return fallbacks
