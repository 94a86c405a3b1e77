"""Exact, parallelisable evaluation of a long chain of float32 additions of same-sign terms,
        s <- fl(s + a_k),   k = 0 .. n-1,   s < 0,  a_k <= 0        (round to nearest even, no FMA)
which is what one voxel's log-probability of one class goes through when many bundles hit it (`priors += L * freq`,
base.cpp:306-307): in the 2 cm `merged` workload the voxels next to the camera receive ~92 000 such updates per frame, strictly
in bundle order, and that single dependent chain is the critical path of the tile kernel.

Observation.  While |s| stays inside one binade [2^e, 2^(e+1)), s is a multiple of u = 2^(e-23) and
        fl(s + a) = s + round_to_nearest_even_on_grid_u(a)      -- the magnitudes add as INTEGERS  M <- M + q_k,  M = |s| / u,
with q_k = floor(x_k) + (frac(x_k) > 1/2), x_k = |a_k| / u, except for exact ties frac(x_k) = 1/2, where the result is forced
to an even M (the only place the running value matters, and only through its parity).  Each record is therefore a function
parity -> (increment, new parity) with two table entries, function composition is associative, and the whole chain becomes a
parallel prefix scan over 2-entry tables plus a search for the first prefix that leaves the binade (|M| >= 2^24), where one
ordinary float addition is done and the grid coarsens.  Same-sign terms make M monotone, so there are at most ~150 binade
crossings in a chain of any length.  The result is bit-identical to the sequential loop; tests/test_exact_float_chain.py checks
it on random and adversarial chains.  (Device plan: lanes = records, one warp scan per 32 records and class group, long chains
split over several warps whose segment functions are composed - DESIGN.md, next round.)
"""
import numpy as np

F = np.float32


def sequential(s0, terms):
    s = F(s0)
    for a in terms:
        s = F(s + F(a))
    return s


def _decompose(x):
    """float32 -> (integer mantissa m, exponent e) with |x| = m * 2^e, m < 2^24 (m = 0 for zero)."""
    bits = int(np.array(abs(float(x)), dtype=np.float32).view(np.uint32))
    exp, frac = bits >> 23, bits & 0x7FFFFF
    if exp == 0:
        return frac, -149                      # subnormal
    return frac | 0x800000, exp - 150


def record_table(a, grid_exp):
    """The two-entry function of one record on the grid u = 2^grid_exp: for parity p of the running magnitude M returns
    (increment, new parity)."""
    m, e = _decompose(a)
    shift = grid_exp - e                        # |a| / u = m * 2^(e - grid_exp)
    if m == 0:
        return ((0, 0), (0, 1))
    if shift <= 0:                              # |a| is a multiple of u: exact
        q = m << (-shift)
        return ((q, q & 1), (q, (q & 1) ^ 1))
    n, rem, half = m >> shift, m & ((1 << shift) - 1), 1 << (shift - 1)
    if rem != half:                             # ordinary rounding, independent of the running value
        q = n + (1 if rem > half else 0)
        return ((q, q & 1), (q, (q & 1) ^ 1))
    # exact tie: M + n + 1/2 rounds to the even neighbour
    out = []
    for p in (0, 1):
        base_parity = p ^ (n & 1)
        q = n + (1 if base_parity else 0)
        out.append((q, 0))
    return tuple(out)


def compose(f, g):
    """(g after f) as a two-entry table."""
    out = []
    for p in (0, 1):
        inc1, p1 = f[p]
        inc2, p2 = g[p1]
        out.append((inc1 + inc2, p2))
    return tuple(out)


def scan_sum(s0, terms, block=32):
    """Same value as sequential(s0, terms), computed block-wise with prefix scans (a block = what one warp would take)."""
    s = F(s0)
    assert s < 0 and np.isfinite(s)
    terms = [F(a) for a in terms]
    i, n = 0, len(terms)
    while i < n:
        m, e = _decompose(s)
        if m < 0x800000:                        # subnormal running value: not worth a fast path
            s = F(s + terms[i]); i += 1
            continue
        grid_exp = e                            # u = 2^e, M = m in [2^23, 2^24)
        chunk = terms[i:i + block]
        tables = [record_table(a, grid_exp) for a in chunk]
        # inclusive scan of the record functions (sequential here; a Hillis-Steele / shuffle scan on the device)
        prefix, acc = [], None
        for t in tables:
            acc = t if acc is None else compose(acc, t)
            prefix.append(acc)
        p0 = m & 1
        # first record whose prefix leaves the binade
        leave = next((k for k, f in enumerate(prefix) if m + f[p0][0] >= (1 << 24)), None)
        if leave is None:
            m_new = m + prefix[-1][p0][0]
            s = F(-np.ldexp(float(m_new), grid_exp))
            i += len(chunk)
            continue
        if leave > 0:
            m_new = m + prefix[leave - 1][p0][0]
            s = F(-np.ldexp(float(m_new), grid_exp))
        s = F(s + chunk[leave])                 # the crossing step itself: one ordinary addition on the coarser grid
        i += leave + 1
    return s


def chunk_table(terms, grid_exp):
    """Composite two-entry function of a whole chunk on one grid (what one warp produces for its share of a long chain)."""
    acc = None
    for a in terms:
        t = record_table(F(a), grid_exp)
        acc = t if acc is None else compose(acc, t)
    return acc


def chunked_sum(s0, terms, chunk=1024):
    """Two-level version for chains that are split over many warps (the hot voxels): pass A is embarrassingly parallel over chunks,
    pass B is a short sequential walk over one table per chunk.
      A: every chunk guesses the binade it will start in from a float64 prefix sum of the raw terms (cheap, order-insensitive)
         and folds its records into ONE two-entry table on that grid;
      B: with the exact running value, a chunk's table is applied if the guess was right and the chunk stays inside the binade
         (the overwhelmingly common case: same-sign terms cross a binade only ~once per doubling of |s|); otherwise that one
         chunk is re-evaluated exactly with scan_sum()."""
    s = F(s0)
    terms = np.asarray(terms, dtype=np.float32)
    n = len(terms)
    starts = list(range(0, n, chunk))
    approx = float(s) + np.concatenate([[0.0], np.cumsum(terms.astype(np.float64))])
    guesses, tables = [], []
    for c0 in starts:                                                     # pass A (parallel on the device)
        g = _decompose(F(approx[c0]))[1] if approx[c0] != 0 else -149
        guesses.append(g)
        tables.append(chunk_table(terms[c0:c0 + chunk], g))
    fallbacks = 0
    for k, c0 in enumerate(starts):                                       # pass B
        m, e = _decompose(s)
        inc, _ = tables[k][m & 1]
        if m >= 0x800000 and e == guesses[k] and m + inc < (1 << 24):
            s = F(-np.ldexp(float(m + inc), e))
        else:
            s = scan_sum(s, terms[c0:c0 + chunk])
            fallbacks += 1
    return s, fallbacks
