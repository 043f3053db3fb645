"""CPU: the ALGORITHM of the on-device categorical draws (k_greedy_sample with draws, csrc/wb_kernels.cu) restated in NumPy, against the
sequential arithmetic of std::discrete_distribution (libstdc++: probabilities normalised by their sum, running sums, last one forced to 1,
lower_bound) that the reference's whisper_sample_token_topk uses (src/whisper.cpp:6545-6618).

The kernel gives every one of its 1024 threads a contiguous slice of the vocabulary, adds the slice sums with a tree, chains the normalised
slices with a Hillis-Steele scan in double and counts the running sums below the uniform.  Only the association of the double additions
differs from the sequential loop, so a draw can change only when the uniform falls within ~1e-13 of a boundary of the cumulative
distribution: on thousands of draws from flat and peaked distributions the ids must be identical, and the largest difference between
the two cumulative distributions is reported."""
import numpy as np


def canonical_uniforms(seed, n):
    """std::generate_canonical<double, 53>(std::mt19937(seed)): two 32-bit outputs per number, (a + b * 2^32) / 2^64"""
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)                     # init_genrand(seed), the seeding of std::mt19937
    raw = bg.random_raw(2 * n).astype(np.float64)
    u = (raw[0::2] + raw[1::2] * 4294967296.0) / 18446744073709551616.0
    return np.minimum(u, np.nextafter(1.0, 0.0))


def sequential(probs, u):
    p = probs.astype(np.float64)
    total = 0.0
    for v in p:                                  # std::accumulate(..., 0.0)
        total += v
    cp = np.empty_like(p)
    run = 0.0
    for i, v in enumerate(p / total):            # std::partial_sum of the normalised values
        run += v
        cp[i] = run
    cp[-1] = 1.0
    return np.searchsorted(cp, u, side="left"), cp     # lower_bound


def kernel_restated(probs, u, T=1024):
    V = len(probs)
    S = (V + T - 1) // T
    p = np.zeros(S * T, np.float64); p[:V] = probs.astype(np.float64)
    sl = p.reshape(T, S)
    ls = np.zeros(T)
    for j in range(S):                           # every thread adds its slice in index order
        ls += sl[:, j]
    red = ls.copy()
    o = T // 2
    while o > 0:                                 # tree reduction in shared memory
        red[:o] += red[o:2 * o]
        o //= 2
    total = red[0]
    ln = np.zeros(T)
    for j in range(S):
        ln += sl[:, j] / total
    sc = ln.copy()
    o = 1
    while o < T:                                 # Hillis-Steele inclusive scan
        add = np.concatenate([np.zeros(o), sc[:-o]])
        sc = sc + add
        o *= 2
    base = sc - ln
    run = base.copy()
    cp = np.empty((T, S))
    for j in range(S):
        run = run + sl[:, j] / total
        cp[:, j] = run
    cp = cp.reshape(-1)[:V]
    ids = np.array([int(np.count_nonzero(cp[:V - 1] < x)) for x in u])       # the last running sum is forced to 1: never below u
    return np.minimum(ids, V - 1), cp


def test_parallel_formulation_draws_the_same_ids_as_libstdcxx_sequential_sums():
    rng = np.random.default_rng(17)
    V = 51866
    n_draws = 0; worst = 0.0
    for trial in range(40):
        logits = (rng.standard_normal(V) * (3.0 if trial % 2 else 9.0)).astype(np.float32)
        if trial % 3 == 0:
            logits[rng.integers(0, V, 20000)] = -np.inf                         # suppressed tokens: probability 0 (flat stretches of the distribution)
        m = np.max(logits[np.isfinite(logits)])
        lse = np.float32(np.log(np.sum(np.exp((logits[np.isfinite(logits)] - m).astype(np.float32)), dtype=np.float32)) + m)
        probs = np.where(np.isfinite(logits), np.exp((logits - lse).astype(np.float32)), np.float32(0.0)).astype(np.float32)
        u = canonical_uniforms(trial, 64)
        a, cpa = sequential(probs, u)
        b, cpb = kernel_restated(probs, u)
        worst = max(worst, float(np.max(np.abs(cpa[:-1] - cpb[:-1]))))
        assert np.array_equal(a, b), (trial, np.nonzero(a != b)[0][:5])
        n_draws += len(u)
    print("%d draws identical; largest difference between the two cumulative distributions %.2e" % (n_draws, worst))
    assert worst < 1e-12
