"""The decomposition behind the level-by-level early drop (csrc/k_early2.cu), checked on the CPU.

project_(early_out = true) (include/impl/scene.hpp:411-510) walks a subset in order; at the first *reaching*
element whose 1-based position is >= tests[i] = uint32(0.05f * (i + 1) * n) it extrapolates the final count
and drops the hypothesis when the bound is below the acceptance threshold; each element serves at most one
checkpoint (:492-506).  tm_query_run(early_out = 2) evaluates that per checkpoint range:
range L = [tm_early_level_begin(n, L), tm_early_level_begin(n, L + 1)).  The claim the kernels rest on:

    if every range 1..18 is non-empty and holds a reaching element, checkpoint L fires at the first reaching
    element of range L, with corrs = (inliers before the range) + (that element's inlier bit);

otherwise the hypothesis is walked sequentially.  This file states the sequential rule and the range rule
side by side in plain Python and compares them on random reach / inlier patterns; the range bounds come from
the C-ABI export the device code shares (no GPU needed)."""
import math

import numpy as np

F = np.float32


def _tests(n):
    return [int(F(F(F(0.05) * F(i + 1)) * F(n))) for i in range(18)]


def _upper(tried, nsub, corrs):
    """scene.hpp:493-500 with the casts as gcc/x86-64 compiles them (see oracle.hpp early_drop_upper)."""
    N, x, n = -2.0 - tried, -2.0 - nsub, -1.0 - corrs
    v = (x * n + math.sqrt((x * n * (N - x) * (N - n)) / (N - 1.0))) / N
    a = int(v) & 0xFFFFFFFF
    return int(-1.0 - a) & 0xFFFFFFFF


def _sequential(reach, inl, bound):
    n, t = len(reach), _tests(len(reach))
    nt, corrs = 0, 0
    for p in range(n):
        corrs += int(inl[p])
        if reach[p] and nt < 18 and p + 1 >= t[nt]:  # one checkpoint per element: nt advances once per p
            if F(_upper(p + 1, n, corrs)) < bound:
                return corrs, True, p + 1
            nt += 1
    return corrs, False, n


def _by_ranges(reach, inl, bound, begin):
    n, corrs = len(reach), 0
    for L in range(19):
        b0, b1 = begin(n, L), begin(n, L + 1)
        if L >= 1:
            first = next((p for p in range(b0, b1) if reach[p]), None)
            if first is None:
                return None  # irregular: walked sequentially by the product
            c = corrs + int(inl[first])
            if F(_upper(first + 1, n, c)) < bound:
                return c, True, first + 1
        corrs += int(np.sum(inl[b0:b1]))
    return corrs, False, n


def test_level_bounds_follow_the_reference_thresholds(built):
    from triplet_match_b200 import capi
    for n in (0, 1, 2, 5, 19, 20, 21, 39, 40, 41, 100, 1000, 56789, 186213, (1 << 24) + 1, 2 ** 31 - 1):
        b = [capi.early_level_begin(n, L) for L in range(-1, 21)]
        assert b[0] == 0 and b[1] == 0 and b[20] == n and b[21] == n
        assert all(x <= y for x, y in zip(b, b[1:]))
        t = _tests(n)
        for L in range(1, 19):
            assert b[L + 1] == min(max(t[L - 1] - 1, 0), n)
    # known answers (part of the C-ABI contract)
    assert [capi.early_level_begin(56789, L) for L in (1, 2, 10, 18, 19)] == [2838, 5677, 28393, 51109, 56789]
    assert [capi.early_level_begin(30, L) for L in range(20)] == [0, 0, 2, 3, 5, 6, 8, 9, 11, 12, 14, 15, 17, 18, 20,
                                                                  21, 23, 24, 26, 30]


def test_range_rule_equals_the_sequential_walk(built):
    from triplet_match_b200 import capi
    cache = {}

    def begin(n, L):
        if (n, L) not in cache:
            cache[(n, L)] = capi.early_level_begin(n, L)
        return cache[(n, L)]

    rng = np.random.default_rng(0)
    regular = irregular = dropped = 0
    for it in range(6000):
        n = int(rng.integers(0, 40)) if it % 3 == 0 else int(rng.integers(20, 600))
        pr, pi = rng.random() ** 0.5, rng.random()
        reach = rng.random(n) < pr
        inl = reach & (rng.random(n) < pi)
        bound = F(rng.random() * n * 1.2)
        seq, rng_rule = _sequential(reach, inl, bound), _by_ranges(reach, inl, bound, begin)
        if rng_rule is None:
            irregular += 1
            continue
        regular += 1
        dropped += int(seq[1])
        assert seq == rng_rule, (n, seq, rng_rule)
    assert regular > 1500 and irregular > 500 and 100 < dropped < regular - 100
