"""The shortcut in tsdf_batch (csrc/ksg_kernels.cuh): when every weight of a batch is >= 0, the voxel's weight is already >= 1e-6 and
(w + tree_sum(u)) * 1.001 < max_weight, the recurrence  w <- (fl(w + u) < 1e-6) ? w : min(max_weight, fl(w + u))  equals the bare float
chain w <- fl(w + u).  Checked here in numpy float32 with the device's butterfly summation order, on random and adversarial batches
(sums just below / above the bound, tiny weights, zeros).  CPU only."""
import numpy as np

F = np.float32
EPS = F(1e-6)


def butterfly_sum(u):
    v = u.astype(F).copy()
    o = 16
    while o > 0:
        v = (v + v[np.arange(32) ^ o]).astype(F)      # usum += __shfl_xor_sync(usum, o): every lane ends with the same value
        o >>= 1
    assert np.all(v == v[0])
    return v[0]


def general(w, u, max_w):
    before = np.zeros(32, F)
    wc = F(w)
    for j in range(32):
        before[j] = wc
        nw = F(wc + u[j])
        if not (nw < EPS):
            wc = min(F(max_w), nw)
    return before, wc


def bare(w, u):
    before = np.zeros(32, F)
    wc = F(w)
    for j in range(32):
        before[j] = wc
        wc = F(wc + u[j])
    return before, wc


def plain_condition(w, u, max_w):
    return bool(np.all(u >= 0)) and bool(F(w) >= EPS) and bool(F(F(F(w) + butterfly_sum(u)) * F(1.001)) < F(max_w))


def test_bare_chain_equals_the_clamped_recurrence_whenever_the_shortcut_is_taken():
    rng = np.random.default_rng(7)
    taken = refused = 0
    for case in range(4000):
        max_w = F(rng.choice([10000.0, 100.0, 1.0, 3.5]))
        kind = case % 5
        if kind == 0:
            u = rng.random(32).astype(F) * F(rng.choice([1e-3, 1.0, 50.0]))
            w = F(rng.random() * float(max_w))
        elif kind == 1:      # sum lands within a few ulp..1 % of the bound
            u = rng.random(32).astype(F)
            w = F(0.5)
            target = float(max_w) / 1.001 * (1.0 + rng.uniform(-0.01, 0.01))
            u = (u * F((target - 0.5) / float(u.sum(dtype=np.float64)))).astype(F)
        elif kind == 2:      # tiny weights around 1e-6
            u = (rng.random(32) * 3e-6).astype(F)
            w = F(rng.uniform(0, 3e-6))
        elif kind == 3:      # zeros and one large
            u = np.zeros(32, F)
            u[rng.integers(32)] = F(rng.uniform(0, 2 * float(max_w)))
            w = F(rng.uniform(1e-6, float(max_w)))
        else:                # already saturated or negative entries: shortcut must be refused or harmless
            u = (rng.standard_normal(32) * 0.3).astype(F)
            w = F(max_w)
        if plain_condition(w, u, max_w):
            taken += 1
            b0, w0 = general(w, u, max_w)
            b1, w1 = bare(w, u)
            assert w0.tobytes() == F(w1).tobytes() and b0.tobytes() == b1.tobytes(), (case, w, max_w)
        else:
            refused += 1
    assert taken > 800 and refused > 800
