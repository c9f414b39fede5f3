"""The public API end to end on the GPU against the reference's outputs (``tests/golden/api.npz``)
— tolerance 1e-10 absolute on every MI / entropy value (north_star) — plus the reference's
bitwise-determinism contract (``tests/unit/test_driver.py:921-923, 989-990, 1555-1557``)."""
import warnings

import numpy as np
import pytest

import ennemi_b200 as eb

pytestmark = pytest.mark.gpu
TOL = 1e-10


def near(a, b, tol=TOL):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    with np.errstate(invalid="ignore"):
        return a.shape == b.shape and bool(np.all(same | (np.abs(a - b) <= tol)))


def test_api_cases_match_reference_outputs(golden_api):
    g = golden_api
    x3, y, cond, mask, lags = (g["inputs"][k] for k in ("x3", "y", "cond", "mask", "lags"))
    yd, xd = g["mi_discrete_y"]["yd"], g["mi_discrete_x"]["xd"]
    ynan, xnan = g["mi_dropnan"]["ynan"], g["mi_dropnan"]["xnan"]
    data5 = np.column_stack((x3, y, cond[:, 0]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert near(eb.estimate_mi(y, x3, lags), g["mi_lags"]["out"])
        assert near(eb.estimate_mi(y, x3, lags, k=5, preprocess=False), g["mi_lags_k5_nopre"]["out"])
        assert near(eb.estimate_mi(y, x3[:, :2], lags, cond=cond), g["mi_cond"]["out"])
        assert near(eb.estimate_mi(y, x3[:, :2], lags, cond=cond, cond_lag=1), g["mi_cond_lag1"]["out"])
        assert near(eb.estimate_mi(y, x3[:, :2], lags, cond=cond, cond_lag=g["mi_cond_lag2d"]["cond_lag"]), g["mi_cond_lag2d"]["out"])
        assert near(eb.estimate_mi(y, x3, lags, mask=mask), g["mi_mask"]["out"])
        assert near(eb.estimate_mi(y, x3[:, 0], [0, 2], mask=mask, cond=cond[:, 0]), g["mi_mask_cond"]["out"])
        assert near(eb.estimate_corr(y, x3, lags), g["corr_lags"]["out"])
        assert near(eb.estimate_mi(ynan, xnan, [0, 1], drop_nan=True), g["mi_dropnan"]["out"])
        assert near(eb.estimate_mi(yd, x3, [0, 2], discrete_y=True), g["mi_discrete_y"]["out"])
        assert near(eb.estimate_mi(y, xd, [0, 2], discrete_x=True, k=4), g["mi_discrete_x"]["out"])
        assert near(eb.estimate_mi(yd, x3[:, :2], [0, 1], discrete_y=True, cond=cond), g["mi_discrete_y_cond"]["out"])
        assert near(eb.estimate_mi(y, xd, [0, 1], discrete_x=True, cond=cond[:, 0]), g["mi_discrete_x_cond"]["out"])
        assert near(eb.estimate_mi(yd, xd, [0, 1], discrete_x=True, discrete_y=True), g["mi_discrete_both"]["out"])
        assert near(eb.pairwise_mi(x3), g["pairwise"]["out"])
        assert near(eb.pairwise_mi(data5, k=4), g["pairwise5"]["out"])
        assert near(eb.pairwise_corr(data5), g["pairwise5_corr"]["out"])
        assert near(eb.pairwise_mi(np.column_stack((x3, y)), cond=cond, mask=mask), g["pairwise_cond_mask"]["out"])
        assert near(eb.pairwise_mi(np.column_stack((x3[:, 0], xd, yd, y)), discrete=[False, True, True, False]), g["pairwise_discrete"]["out"])
        assert near(eb.pairwise_mi(np.column_stack((xnan, ynan)), drop_nan=True), g["pairwise_dropnan"]["out"])
        assert near(eb.estimate_entropy(x3), g["ent_cols"]["out"])
        assert near(eb.estimate_entropy(x3, multidim=True, k=5), g["ent_multidim"]["out"])
        assert near(eb.estimate_entropy(x3[:, :2], cond=cond), g["ent_cond"]["out"])
        assert near(eb.estimate_entropy(x3[:, :2], cond=cond[:, 0], multidim=True, mask=mask), g["ent_cond_multidim_mask"]["out"])
        assert near(eb.estimate_entropy(ynan, drop_nan=True), g["ent_1d_dropnan"]["out"])


def test_docs_known_answers_on_gpu(golden_api):
    g = golden_api
    rng = np.random.default_rng(1234)
    data = rng.multivariate_normal([0, 0], [[1, 0.8], [0.8, 1]], size=800)
    z = rng.normal(0, 1, size=800)
    out = eb.estimate_corr(data[:, 1], np.column_stack((data[:, 0], z)))
    assert near(out, g["doc_tutorial_165"]["out"]) and near(out, g["doc_tutorial_165"]["printed"], 5e-9)
    rng = np.random.default_rng(1234)
    x = rng.gamma(1.0, 1.0, size=400); y = np.zeros(400); y[1:] = x[0:-1]; y += rng.normal(0, 0.01, size=400)
    assert near(eb.estimate_corr(y, x, lag=[1, 0, -1]), g["doc_tutorial_208"]["printed"], 5e-9)
    rng = np.random.default_rng(1234)
    data = rng.multivariate_normal([0, 0], [[1, 0.8], [0.8, 1]], size=800)
    assert near(eb.estimate_mi(np.exp(data[:, 1]), np.exp(5 * data[:, 0])), g["doc_issues_68"]["printed"], 5e-9)
    rng = np.random.default_rng(1234)
    data = rng.multivariate_normal([0.5, 0.5], [[1, 0.8], [0.8, 1]], size=800)
    assert eb.estimate_mi(np.maximum(0, data[:, 1]), np.maximum(0, data[:, 0]), preprocess=False)[0, 0] == -np.inf
    rng = np.random.default_rng(1234)
    data = rng.multivariate_normal([0, 0], [[1, 0.8], [0.8, 1]], size=800)
    x = np.concatenate((data[:, 0], data[:, 0] + rng.normal(0, 0.01, size=800), data[:, 0] + rng.normal(0, 0.01, size=800)))
    y = np.concatenate((data[:, 1], data[:, 1] + rng.normal(0, 0.01, size=800), data[:, 1] + rng.normal(0, 0.01, size=800)))
    assert near(eb.estimate_mi(y, x), g["doc_issues_168"]["printed"], 5e-9)


def test_bitwise_determinism_contract():
    rng = np.random.default_rng(2)
    x = rng.normal(size=(3_000, 4)); y = x[:, 0] + rng.normal(size=3_000)
    a = eb.estimate_mi(y, x, lag=[0, 1, 2])
    b = eb.estimate_mi(y, x, lag=[0, 1, 2], max_threads=1)
    assert np.array_equal(a, b)                                             # scheduling never changes bits
    assert np.array_equal(eb.estimate_corr(y, x), eb.estimate_mi(y, x, normalize=True))
    pw = eb.pairwise_mi(x)
    assert np.array_equal(pw, pw.T, equal_nan=True)
    d = rng.integers(0, 4, 3_000)
    m1 = eb.estimate_mi(d, y, discrete_y=True)
    m2 = eb.estimate_mi(y, d, discrete_x=True)
    assert m1[0, 0] == m2[0, 0]                                             # x<->y symmetry with a discrete variable


def test_analytic_values():
    """The reference's statistical tests in miniature: Gaussian MI, CMI chain, uniform entropy."""
    rng = np.random.default_rng(0)
    for rho in (0.0, 0.5, 0.9):
        d = rng.multivariate_normal([0, 0], [[1, rho], [rho, 1]], size=20_000)
        mi = eb.estimate_mi(d[:, 1], d[:, 0])[0, 0]
        assert abs(mi - (-0.5 * np.log(1 - rho ** 2))) < 0.03
    u = rng.uniform(0, 2, size=20_000)
    assert abs(float(eb.estimate_entropy(u)) - np.log(2)) < 0.02
    z = rng.normal(size=20_000); x = z + rng.normal(size=20_000) * 0.5; y = z + rng.normal(size=20_000) * 0.5
    assert abs(eb.estimate_mi(y, x, cond=z)[0, 0]) < 0.02                   # conditionally independent
    assert eb.estimate_mi(y, x)[0, 0] > 0.3


def test_device_column_path_on_gpu(golden_api, monkeypatch):
    """Cached device columns + prep_kernel: (a) bit-identical to the host-prepared block through the
    same kernels, (b) the reference's API outputs within 1e-10, (c) NaN / inf reporting."""
    from ennemi_b200 import api, _align, _native as nat, _columns
    rng = np.random.default_rng(3)
    n = 50_000
    x = rng.normal(2.0, 3.0, size=n); y = 0.4 * x + rng.normal(size=n); z = rng.normal(size=(n, 2)) + x[:, None] * 0.2
    xs, ys, zs = _align.rescaled(x, y, z, False, False)
    assert api.DEVICE_COLUMNS_MIN_ROWS <= n
    assert eb.estimate_mi(y, x)[0, 0] == nat.ksg_mi(nat.pack_coords([xs, ys]), 3)            # same bits in, same bits out
    assert eb.estimate_mi(y, x, cond=z)[0, 0] == nat.cmi(nat.pack_coords([xs, ys, zs]), 3)
    lagged = eb.estimate_mi(y, x, lag=[0, 3, -2], cond=z, cond_lag=[[0, 1], [1, 0], [2, 2]])
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)                             # general host path
    assert np.array_equal(lagged, eb.estimate_mi(y, x, lag=[0, 3, -2], cond=z, cond_lag=[[0, 1], [1, 0], [2, 2]]))
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 0)
    g = golden_api
    x3, yy, cond, lags = (g["inputs"][k] for k in ("x3", "y", "cond", "lags"))
    assert near(eb.estimate_mi(yy, x3, lags), g["mi_lags"]["out"])
    assert near(eb.estimate_mi(yy, x3[:, :2], lags, cond=cond, cond_lag=g["mi_cond_lag2d"]["cond_lag"]), g["mi_cond_lag2d"]["out"])
    assert near(eb.pairwise_mi(np.column_stack((x3, yy, cond[:, 0])), k=4), g["pairwise5"]["out"])
    with pytest.raises(ValueError, match="input contains NaNs"):
        eb.estimate_mi(np.where(np.arange(600) == 5, np.nan, yy), x3)
    with pytest.raises(ValueError, match="data must be finite"):
        eb.estimate_mi(yy, np.where(np.arange(600) == 7, np.inf, x3[:, 0]), preprocess=False)


def test_device_statistics_have_numpy_bits():
    """eb2_cache_stats reproduces ndarray.mean() / ndarray.std() bit for bit (NumPy's pairwise summation order)
    for window sizes around every branch of the recursion, offsets and strides."""
    from ennemi_b200 import _native as nat
    rng = np.random.default_rng(5)
    col = rng.normal(3.0, 2.5, size=1_200_000)
    nat.cache_put(777001, col)
    try:
        for n in (1, 5, 7, 8, 9, 127, 128, 129, 136, 255, 256, 257, 1000, 4097, 50_000, 99_991, 262_144, 1_000_000, 1_200_000):
            for off in (0, 3):
                if off + n > col.size:
                    continue
                view = col[off: off + n]
                assert nat.cache_stats(777001, off, n) == (float(view.mean()), float(view.std())), (n, off)
        view = col[5: 5 + 3 * 100_000: 3]
        assert nat.cache_stats(777001, 5, 100_000, stride=3) == (float(view.mean()), float(view.std()))
    finally:
        nat.cache_drop(777001)


def test_one_call_statistics_on_gpu(monkeypatch):
    """One-task calls on large windows: upload, NumPy-exact statistics, rescaling and the estimate in one
    library call (EB2_FLAG_DEVICE_STATS).  Same bits as the host-prepared path; a constant window falls back
    to the path that warns like the reference; page-locked input arrays behave like any other array."""
    import torch
    from ennemi_b200 import api, _align, _native as nat
    rng = np.random.default_rng(11)
    n = 120_000
    x = rng.normal(-1.0, 0.3, size=n); y = np.exp(x) + rng.normal(size=n) * 0.1
    z = rng.normal(size=(n, 2)) + x[:, None]
    xs, ys, zs = _align.rescaled(x, y, z, False, False)
    want = nat.ksg_mi(nat.pack_coords([xs, ys]), 3)
    assert eb.estimate_mi(y, x)[0, 0] == want
    assert nat.last_timing()["launches"] > 8                 # statistics kernels ran inside the estimate call
    assert eb.estimate_mi(y, x, cond=z)[0, 0] == nat.cmi(nat.pack_coords([xs, ys, zs]), 3)
    assert eb.estimate_mi(y[3:], x[:-3])[0, 0] == eb.estimate_mi(y, x, lag=3)[0, 0]
    pin = torch.empty((2, n), dtype=torch.float64, pin_memory=True).numpy()
    pin[0] = x; pin[1] = y
    assert eb.estimate_mi(pin[1], pin[0])[0, 0] == want
    const = np.full(n, 0.25)
    with pytest.warns(UserWarning, match="takes only a single value"):
        got = eb.estimate_mi(const, x)[0, 0]
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    with pytest.warns(UserWarning, match="takes only a single value"):
        ref = eb.estimate_mi(const, x)[0, 0]
    assert got == ref or (np.isnan(got) and np.isnan(ref))
    with pytest.raises(ValueError, match="input contains NaNs"):
        monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 0)
        eb.estimate_mi(np.where(np.arange(n) == 5, np.nan, y), x)


def test_block_upload_on_gpu(monkeypatch):
    """eb2_cache_put_block: a row-major (n, ncols) block uploaded with one copy and split into columns on the
    device holds exactly the caller's values (checked through the NumPy-exact statistics and through estimates
    on the cached columns), for dense and strided blocks and sizes off the 32 x 32 tile grid; pairwise_mi /
    estimate_mi through the block path return the same bits as through per-column uploads."""
    from ennemi_b200 import api, _columns, _native as nat
    rng = np.random.default_rng(21)
    for n, ncols, ld in ((70_001, 37, 37), (50_000, 5, 9), (33, 2, 2), (131_072, 64, 64)):
        wide = rng.normal(1.0, 2.0, size=(n, ld))
        block = wide[:, :ncols]
        keys = list(range(880_000, 880_000 + ncols))
        nat.cache_put_block(keys, block)
        try:
            means, stds = nat.cache_stats_many(keys, [0] * ncols, n)
            for j in range(ncols):
                col = np.ascontiguousarray(block[:, j])
                assert (means[j], stds[j]) == (col.mean(), col.std()), (n, ncols, j)
                assert nat.cache_stats(keys[j], 0, n) == (float(col.mean()), float(col.std()))
            off = 7
            m2, s2 = nat.cache_stats_many(keys[:2], [off, 0], n - off)
            assert m2[0] == block[off:, 0].mean() and s2[1] == block[:n - off, 1].std()
            if n > 1000:
                descs = [nat.ColDesc(keys[0], 0, 1, 0.0, 0.0, 0, 0, 1), nat.ColDesc(keys[ncols - 1], 0, 1, 0.0, 0.0, 0, 0, 1)]
                want = nat.ksg_mi(nat.pack_coords([block[:, 0], block[:, ncols - 1]]), 3)
                assert nat.ksg_mi_cols(descs, n, 3, flags=nat.FLAG_SINGLE_USE) == want
        finally:
            for key in keys:
                nat.cache_drop(key)
    data = rng.normal(size=(60_000, 7)) @ rng.normal(size=(7, 7))
    cond = rng.normal(size=(60_000, 2)) + data[:, :2]
    assert api.DEVICE_COLUMNS_MIN_ROWS <= 60_000
    got = (eb.pairwise_mi(data), eb.estimate_mi(data[:, 0], data[:, 1:], lag=[0, 5]), eb.pairwise_mi(data[:, :4], cond=cond),
           eb.pairwise_mi(data, preprocess=False))
    monkeypatch.setattr(_columns.ColumnStore, "MIN_BLOCK_DENSITY", 2.0)              # column-by-column uploads
    per_col = (eb.pairwise_mi(data), eb.estimate_mi(data[:, 0], data[:, 1:], lag=[0, 5]), eb.pairwise_mi(data[:, :4], cond=cond),
               eb.pairwise_mi(data, preprocess=False))
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)                      # general host path
    host = eb.pairwise_mi(data)
    for a, b in zip(got, per_col):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(got[0], host, equal_nan=True)


def test_graph_replay_matches_eager(monkeypatch):
    """A repeated resident estimate (device pointer input) is captured into a CUDA graph on its third call and
    replayed afterwards (EB2_GRAPH=0 keeps every call eager).  The partial block (sums, zero counters, rows, pairs evaluated) must be identical
    to the eager call's, the replay must read whatever the caller's buffer holds now, a row shard is a signature of
    its own, and non-finite data are still reported."""
    import torch
    from ennemi_b200 import _native as nat
    rng = np.random.default_rng(31)
    n = 400_000
    d = rng.multivariate_normal([0, 0], [[1, 0.5], [0.5, 1]], size=n)
    dev = torch.from_numpy(nat.pack_coords([d[:, 0], d[:, 1]])).cuda()
    ptr = int(dev.data_ptr())

    def call(lo=0, hi=n):
        part = nat.ksg_mi_rows(ptr, n, 3, lo, hi, flags=nat.FLAG_DEVICE_INPUT)
        part[nat.P_PAIRS] = 0.0          # the work counter depends on the (unordered) slot layout of the run, the result does not
        return part

    monkeypatch.setenv("EB2_GRAPH", "0")
    eager = call()
    monkeypatch.setenv("EB2_GRAPH", "1")
    for _ in range(6):
        assert np.array_equal(call(), eager)
    assert nat.last_timing()["knn_ms"] > 0
    d2 = rng.multivariate_normal([0, 0], [[1, -0.3], [-0.3, 1]], size=n)
    dev.copy_(torch.from_numpy(nat.pack_coords([d2[:, 0], d2[:, 1]])))
    torch.cuda.synchronize()
    replayed = call()
    monkeypatch.setenv("EB2_GRAPH", "0")
    assert np.array_equal(replayed, call()) and not np.array_equal(replayed, eager)
    half = call(0, n // 2)
    monkeypatch.delenv("EB2_GRAPH")                       # the default: graphs on
    for _ in range(5):
        assert np.array_equal(call(0, n // 2), half)
    assert np.array_equal(call(), replayed)
    dev[0, 5] = float("nan")
    torch.cuda.synchronize()
    with pytest.raises(ValueError, match="data must be finite"):
        call()


def test_pairwise_full_size_vs_reference_calls():
    """BASELINE.json configs[3] at full size through the public API: pairwise_mi on (100000, 64) — the block upload,
    device statistics, prepared-variable cache and batched pair path the bench's pairwise number comes from — against
    the reference's own SciPy calls (oracle "scipy" backend) on the columns prepared as the reference prepares them
    (ennemi/_driver.py:703-707, 871-902).  18 sampled pairs within 1e-10; eps / n_x / n_y of two pairs bit for bit."""
    import oracle
    from ennemi_b200 import _align, _native as nat
    rng = np.random.default_rng(0)
    mix = np.eye(64) + 0.15 * rng.normal(size=(64, 64))
    data = rng.normal(size=(100_000, 64)) @ mix
    got = eb.pairwise_mi(data, k=3)
    assert got.shape == (64, 64) and np.all(np.isnan(np.diag(got))) and np.array_equal(got, got.T, equal_nan=True)
    pairs = [(0, 1), (0, 63), (62, 63), (31, 32)] + [tuple(sorted(rng.choice(64, 2, replace=False))) for _ in range(14)]
    for t, (i, j) in enumerate(pairs):
        xs, ys, _ = _align.rescaled(data[:, i].copy(), data[:, j].copy(), None, False, False)
        want = oracle.ksg_mi(xs, ys, 3, backend="scipy")
        assert abs(got[i, j] - want["value"]) <= TOL, (i, j, got[i, j], want["value"])
        if t < 2:
            value, parts = nat.ksg_mi(nat.pack_coords([xs, ys]), 3, details=True)
            assert np.array_equal(parts["eps"], want["eps"])
            assert np.array_equal(parts["nx"], want["nx"]) and np.array_equal(parts["ny"], want["ny"])
            assert abs(value - got[i, j]) <= 1e-13


def test_cmi_lag_sweep_full_size_vs_reference_calls():
    """BASELINE.json configs[2] at full size through the public API: estimate_mi(y, x, lag=range(50), cond=z) at
    N = 200,000 with a 3-D condition (ennemi/_driver.py:477-483); three sampled lags against the reference's SciPy
    calls on host-prepared windows (oracle.conditional_mi), eps and all three counts of one lag bit for bit."""
    import oracle
    from ennemi_b200 import _align, _native as nat
    rng = np.random.default_rng(0)
    n = 200_000
    z = rng.normal(size=(n, 3)); x = rng.normal(size=n) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=n)
    lags = np.arange(50)
    got = eb.estimate_mi(y, x, lag=lags, k=3, cond=z)
    assert got.shape == (50, 1)
    for t, lag in enumerate((0, 17, 49)):
        task = _align.MiTask(x, y, lag, 49, 0, 3, None, z, np.zeros(3, dtype=int), False, False, True, False)
        xs, ys, zs = _align.prepare(task)
        want = oracle.conditional_mi(xs, ys, zs, 3, backend="scipy")
        assert abs(got[lag, 0] - want["value"]) <= TOL, (lag, got[lag, 0], want["value"])
        if t == 1:
            value, parts = nat.cmi(nat.pack_coords([xs, ys, zs]), 3, details=True)
            for key in ("eps", "nxz", "nyz", "nz"):
                assert np.array_equal(parts[key], want[key]), key
            assert abs(value - got[lag, 0]) <= 1e-13


def test_conditional_entropy_on_device_columns_gpu(monkeypatch):
    """estimate_entropy(x, cond=...) on device-resident columns (eb2_entropy_cols): the same bits as the
    host route through eb2_entropy, and the reference's chain rule on SciPy's k-th distances within 1e-10."""
    import oracle
    from ennemi_b200 import api
    rng = np.random.default_rng(12)
    n = 30_000
    c = rng.normal(size=(n, 2))
    x = np.column_stack((c[:, 0] + rng.normal(size=n), rng.normal(size=n), c[:, 1] - 0.5 * rng.normal(size=n)))
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    host = eb.estimate_entropy(x, cond=c)
    host_md = eb.estimate_entropy(x, cond=c, multidim=True, k=5)
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 20_000)
    dev = eb.estimate_entropy(x, cond=c)
    assert np.array_equal(dev, host)
    assert np.array_equal(eb.estimate_entropy(x, cond=c, multidim=True, k=5), host_md)
    assert np.array_equal(eb.estimate_entropy(x[:, 1], cond=c[:, 0]), eb.estimate_entropy(x[:, 1], cond=c[:, :1]))
    h_c = oracle.knn_entropy(c, 3, backend="scipy")["value"]
    for j in range(3):
        want = oracle.knn_entropy(np.column_stack((x[:, j], c)), 3, backend="scipy")["value"] - h_c
        assert abs(dev[j] - want) <= 1e-10, j
    # a multi-column variable without a condition: block upload + eb2_entropy_cols, the same bits as the host route
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    host_block = eb.estimate_entropy(x, multidim=True, k=4)
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 20_000)
    assert eb.estimate_entropy(x, multidim=True, k=4) == host_block
    assert abs(float(host_block) - oracle.knn_entropy(x, 4, backend="scipy")["value"]) <= 1e-10
    # separate variables: one block upload, one estimate per cached column (1-D: the two-pointer k-NN kernel)
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    host_cols = eb.estimate_entropy(x, k=2)
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 20_000)
    assert np.array_equal(eb.estimate_entropy(x, k=2), host_cols)
    assert abs(host_cols[1] - oracle.knn_entropy(x[:, 1], 2, backend="scipy")["value"]) <= 1e-10
    bad = x.copy(); bad[17, 2] = np.inf
    with pytest.raises(ValueError, match="data must be finite"):
        eb.estimate_entropy(bad, cond=c)
    with pytest.raises(ValueError, match="data must be finite"):
        eb.estimate_entropy(bad, multidim=True)
    with pytest.raises(ValueError, match="input contains NaNs"):
        eb.estimate_entropy(np.where(np.isinf(bad), np.nan, bad), multidim=True)
    with pytest.raises(ValueError, match="input contains NaNs"):
        eb.estimate_entropy(np.where(np.isinf(bad), np.nan, bad), cond=c)
