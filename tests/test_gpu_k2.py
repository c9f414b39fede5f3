"""The bivariate pipeline (sort-free adaptive grid, ennemi_b200/csrc/eb2_ksg2.cu) on a real B200, through the C ABI.

Bar as everywhere (BASELINE.json north_star): eps and both neighbour counts BIT-EXACT against the reference's
arithmetic, the estimate within 1e-10.  The pipeline is held (1) to the fixtures captured from the unmodified reference
(forced onto small inputs it would normally leave to the general path), (2) to the library's own brute-force kernels
and the CPU oracle on inputs chosen to stress its grid (ties, duplicates, outliers, degenerate ranges, heavy tails,
thin bands, huge offsets), and (3) to the properties its exact fixed-point reduction promises: bitwise identical
values whatever the shard boundaries, the launch order or the batch a pair is estimated in."""
import numpy as np
import pytest

from ennemi_b200 import _native as nat

pytestmark = pytest.mark.gpu

TOL = 1e-10
BRUTE = nat.FLAG_NO_PRUNE | nat.FLAG_BRUTE_COUNT


def close(a, b):
    return a == b or (np.isnan(a) and np.isnan(b)) or abs(a - b) <= TOL


def same_as_brute(x, y, k):
    co = nat.pack_coords([x, y])
    v, d = nat.ksg_mi(co, k, details=True)
    vb, b = nat.ksg_mi(co, k, flags=BRUTE, details=True)
    for key in ("eps", "nx", "ny"):
        assert np.array_equal(d[key], b[key]), (key, int(np.sum(d[key] != b[key])))
    assert close(v, vb), (v, vb)
    return v, d


def test_reference_fixtures_through_the_pipeline(golden_estimators, monkeypatch):
    """Every bivariate fixture of the unmodified reference (n = 40 .. 2,000 rows), forced through the pipeline."""
    monkeypatch.setenv("EB2_K2_MIN", "2")
    seen = 0
    for name, c in golden_estimators.items():
        if not name.startswith("ksg"):
            continue
        k = int(c["k"])
        v, d = nat.ksg_mi(nat.pack_coords([c["x"], c["y"]]), k, details=True)
        assert nat.last_pipeline() == (1 if k + 1 <= 8 else 0), (name, k, len(c["x"]))      # (k > 7: the heap variant of the general path)
        for key in ("eps", "nx", "ny"):
            assert np.array_equal(d[key], c[key]), (name, key)
        assert close(v, float(c["value"])), name
        seen += 1
    assert seen >= 3


def _cases():
    rng = np.random.default_rng(7)
    t = rng.standard_t(2, size=(60_000, 2))
    yield "student-t2", t[:, 0], t[:, 1], 3
    u = rng.uniform(size=(40_000, 2)); u[:, 1] = u[:, 0] + 1e-3 * u[:, 1]
    yield "thin-band", u[:, 0], u[:, 1], 3
    g = np.round(rng.normal(size=(30_000, 2)), 1)
    yield "ties", g[:, 0], g[:, 1], 3
    g0 = np.round(rng.normal(size=(30_000, 2)), 0)
    yield "bucket-overflow", g0[:, 0] + 1e-9 * rng.normal(size=30_000), g0[:, 1], 3
    g2 = np.round(rng.normal(size=(30_000, 2)), 3)
    yield "some-ties", g2[:, 0], g2[:, 1], 5
    c = rng.normal(size=(40_000, 2)); c[::7] = c[0]
    yield "duplicates", c[:, 0], c[:, 1], 3
    yield "sorted-x", np.sort(rng.normal(size=50_000)), rng.normal(size=50_000), 3
    yield "constant-y", rng.normal(size=20_000), np.full(20_000, 2.5), 2
    o = rng.normal(size=(60_000, 2)); o[5] = (40.0, -35.0); o[6] = (0.0, 30.0)
    yield "outliers", o[:, 0], o[:, 1], 3
    bm = np.concatenate([rng.normal(-5, 0.1, size=(30_000, 2)), rng.normal(5, 0.1, size=(30_000, 2)), [[0.0, 0.0]]])
    yield "bimodal+lonely", bm[:, 0], bm[:, 1], 3
    big = rng.normal(size=(50_000, 2)) * 1e15 + 1e15
    yield "huge-offset", big[:, 0], big[:, 1], 3
    tiny = rng.normal(size=(30_000, 2)) * 1e-300
    yield "denormal-scale", tiny[:, 0], tiny[:, 1], 3
    e = rng.exponential(size=(50_000, 2))
    yield "exponential", e[:, 0], e[:, 1], 7
    s = rng.normal(size=(2_500, 2))
    yield "barely-above-threshold", s[:, 0], s[:, 1], 1


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_grid_stress_cases_bit_exact(case):
    name, x, y, k = case
    same_as_brute(np.ascontiguousarray(x), np.ascontiguousarray(y), k)


@pytest.mark.parametrize("n,k", [(2_048, 3), (2_049, 1), (4_095, 3), (33_333, 7), (100_000, 3), (262_144, 2)])
def test_against_the_oracle(n, k):
    import oracle
    rng = np.random.default_rng(n)
    d = rng.multivariate_normal([0, 0], [[1, 0.9], [0.9, 1]], size=n)
    x, y = np.ascontiguousarray(d[:, 0]), np.ascontiguousarray(d[:, 1])
    want = oracle.ksg_mi(x, y, k, backend="scipy")
    v, got = nat.ksg_mi(nat.pack_coords([x, y]), k, details=True)
    for key in ("eps", "nx", "ny"):
        assert np.array_equal(got[key], want[key]), key
    assert close(v, want["value"])


def test_value_is_independent_of_sharding_and_order():
    """The digamma sum travels as exact integer limbs: partial blocks of ANY shard boundaries add up to the bits of the
    unsharded estimate, run after run (the reduction does not depend on which GPU, block or warp reduced what)."""
    rng = np.random.default_rng(11)
    n = 400_000
    d = rng.multivariate_normal([0, 0], [[1, 0.6], [0.6, 1]], size=n)
    co = nat.pack_coords([d[:, 0], d[:, 1]])
    whole = nat.ksg_mi(co, 3)
    assert nat.ksg_mi(co, 3) == whole
    for cuts in ([0, n], [0, 1, n], [0, 123_457, n], [0, 50_000, 100_000, 150_000, 200_000, 250_000, 300_000, 350_000, n],
                 [0, n // 3, n // 3, n]):
        parts = [nat.ksg_mi_rows(co.ctypes.data, n, 3, lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:])]
        total = np.sum(parts, axis=0)
        assert total[nat.P_ROWS] == n
        assert nat.ksg_mi_finish(total, n, 3) == whole, cuts


def test_single_process_sharding_over_the_visible_gpus():
    """eb2_sharded_ksg_mi: same bits for every number of GPUs (here: as many as the box shows, at least one)."""
    rng = np.random.default_rng(2)
    n = 300_000
    d = rng.multivariate_normal([0, 0], [[1, 0.3], [0.3, 1]], size=n)
    co = nat.pack_coords([d[:, 0], d[:, 1]])
    whole = nat.ksg_mi(co, 3)
    for g in range(1, min(nat.device_count(), 8) + 1):
        assert nat.sharded_ksg_mi(co, 3, g) == whole, g
    with pytest.raises(ValueError):
        nat.sharded_ksg_mi(co, 3, nat.device_count() + 1)


def test_all_pairs_call_matches_single_estimates():
    """eb2_ksg_mi_pairs (every pair of pairwise_mi in one call, batched per stage) against one estimate per pair, and a
    few pairs against the oracle; a NaN column fails only the pairs that use it."""
    import oracle
    rng = np.random.default_rng(5)
    n, nvar = 30_000, 6
    data = rng.normal(size=(n, nvar)) @ rng.normal(size=(nvar, nvar))
    keys = list(range(9_000, 9_000 + nvar))
    for j, key in enumerate(keys):
        nat.cache_put(key, np.ascontiguousarray(data[:, j]))
    try:
        cols = [nat.ColDesc(key, 0, 1, 0.0, 0.0, 0, 0, 1) for key in keys]          # std = 0: values pass through
        pairs = np.array([(i, j) for i in range(nvar) for j in range(i + 1, nvar)], dtype=np.int32)
        values, status = nat.ksg_mi_pairs(cols, pairs, n, 3)
        assert not status.any()
        for t, (i, j) in enumerate(pairs):
            single = nat.ksg_mi(nat.pack_coords([data[:, i], data[:, j]]), 3)
            assert values[t] == single, (i, j)               # same bits: the sum is order-free
        for t in (0, 7, len(pairs) - 1):
            i, j = pairs[t]
            assert close(values[t], oracle.ksg_mi(data[:, i], data[:, j], 3, backend="scipy")["value"])
        bad = data[:, 2].copy(); bad[17] = np.nan
        nat.cache_put(keys[2], bad)
        values, status = nat.ksg_mi_pairs(cols, pairs, n, 3)
        for t, (i, j) in enumerate(pairs):
            assert (status[t] != 0) == (2 in (i, j)), (i, j, status[t])
            if status[t]:
                assert (int(status[t]) & 0xFF) == nat.ERR_NONFINITE and np.isnan(values[t])
    finally:
        for key in keys:
            nat.cache_drop(key)


def test_all_pairs_call_outside_the_pipeline():
    """Sizes / k the pipeline does not take: the same entry point estimates pair by pair on the general path."""
    rng = np.random.default_rng(9)
    for n, k in ((1_500, 3), (6_000, 9)):
        data = rng.normal(size=(n, 4))
        data[:, 1] += 0.7 * data[:, 0]
        keys = list(range(9_500, 9_504))
        for j, key in enumerate(keys):
            nat.cache_put(key, np.ascontiguousarray(data[:, j]))
        try:
            cols = [nat.ColDesc(key, 0, 1, 0.0, 0.0, 0, 0, 1) for key in keys]
            pairs = np.array([(0, 1), (0, 3), (2, 3)], dtype=np.int32)
            values, status = nat.ksg_mi_pairs(cols, pairs, n, k)
            assert not status.any()
            for t, (i, j) in enumerate(pairs):
                assert close(values[t], nat.ksg_mi(nat.pack_coords([data[:, i], data[:, j]]), k, flags=BRUTE))
        finally:
            for key in keys:
                nat.cache_drop(key)


def test_general_path_takes_over_when_a_bucket_overflows(monkeypatch):
    """Heavily tied data (one value holds most of a column) does not fit the grid's buckets: the call falls back to the
    general path and the answer is still the reference's."""
    import oracle
    rng = np.random.default_rng(3)
    n = 40_000
    x = np.where(rng.uniform(size=n) < 0.7, 1.25, rng.normal(size=n))
    y = rng.normal(size=n)
    want = oracle.ksg_mi(x, y, 3, backend="scipy")
    v, got = nat.ksg_mi(nat.pack_coords([x, y]), 3, details=True)
    assert nat.last_pipeline() == 0, "expected the general path"
    for key in ("eps", "nx", "ny"):
        assert np.array_equal(got[key], want[key]), key
    assert close(v, want["value"])


def test_full_deferral_list_loses_nothing(monkeypatch):
    """More deferred queries than the list holds (forced here by shrinking it): the queries that do not fit are finished
    by the search kernel itself, the ones that fit by the leftover kernel, none twice and none never - run after run
    (the reservations race; the count of the list must not be taken back)."""
    monkeypatch.setenv("EB2_K2_LEFTCAP", "40")
    rng = np.random.default_rng(21)
    x = rng.standard_cauchy(6_000); y = x + rng.standard_cauchy(6_000)
    co = nat.pack_coords([x, y])
    vb, b = nat.ksg_mi(co, 3, flags=BRUTE, details=True)
    for _ in range(25):
        v, d = nat.ksg_mi(co, 3, details=True)
        assert nat.last_pipeline() == 1
        for key in ("eps", "nx", "ny"):
            assert np.array_equal(d[key], b[key]), key
        assert close(v, vb)
    t = rng.standard_t(2, size=(60_000, 2))
    same_as_brute(np.ascontiguousarray(t[:, 0]), np.ascontiguousarray(t[:, 1]), 3)
