"""Host logic (lag alignment, masks, NaN dropping, rescaling, dispatch, pandas wrapping, scheduling)
against the outputs of the reference's public API stored in ``tests/golden/api.npz``.

The estimator seam is routed to the CPU oracle (fixture ``oracle_backend``), whose values are
bit-identical to the reference's, so every comparison here is exact equality."""
import threading
import warnings

import numpy as np
import pytest

import ennemi_b200 as eb


def eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_estimate_mi_cases(oracle_backend, golden_api):
    g = golden_api
    x3, y, cond, mask, lags = (g["inputs"][k] for k in ("x3", "y", "cond", "mask", "lags"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert eq(eb.estimate_mi(y, x3, lags), g["mi_lags"]["out"])
        assert eq(eb.estimate_mi(y, x3, lags, k=5, preprocess=False), g["mi_lags_k5_nopre"]["out"])
        assert eq(eb.estimate_mi(y, x3[:, :2], lags, cond=cond), g["mi_cond"]["out"])
        assert eq(eb.estimate_mi(y, x3[:, :2], lags, cond=cond, cond_lag=1), g["mi_cond_lag1"]["out"])
        assert eq(eb.estimate_mi(y, x3[:, :2], lags, cond=cond, cond_lag=g["mi_cond_lag2d"]["cond_lag"]),
                  g["mi_cond_lag2d"]["out"])
        assert eq(eb.estimate_mi(y, x3, lags, mask=mask), g["mi_mask"]["out"])
        assert eq(eb.estimate_mi(y, x3[:, 0], [0, 2], mask=mask, cond=cond[:, 0]), g["mi_mask_cond"]["out"])
        assert eq(eb.estimate_corr(y, x3, lags), g["corr_lags"]["out"])
        assert eq(eb.estimate_mi(g["mi_dropnan"]["ynan"], g["mi_dropnan"]["xnan"], [0, 1], drop_nan=True),
                  g["mi_dropnan"]["out"])
        yd, xd = g["mi_discrete_y"]["yd"], g["mi_discrete_x"]["xd"]
        assert eq(eb.estimate_mi(yd, x3, [0, 2], discrete_y=True), g["mi_discrete_y"]["out"])
        assert eq(eb.estimate_mi(y, xd, [0, 2], discrete_x=True, k=4), g["mi_discrete_x"]["out"])
        assert eq(eb.estimate_mi(yd, x3[:, :2], [0, 1], discrete_y=True, cond=cond), g["mi_discrete_y_cond"]["out"])
        assert eq(eb.estimate_mi(y, xd, [0, 1], discrete_x=True, cond=cond[:, 0]), g["mi_discrete_x_cond"]["out"])
        assert eq(eb.estimate_mi(yd, xd, [0, 1], discrete_x=True, discrete_y=True), g["mi_discrete_both"]["out"])
        assert eq(eb.estimate_mi(yd, xd, 0, discrete_x=True, discrete_y=True, cond=(cond[:, 0] > 0).astype(int)),
                  g["mi_discrete_both_cond"]["out"])


def test_pairwise_and_entropy_cases(oracle_backend, golden_api):
    g = golden_api
    x3, y, cond, mask = (g["inputs"][k] for k in ("x3", "y", "cond", "mask"))
    yd, xd = g["mi_discrete_y"]["yd"], g["mi_discrete_x"]["xd"]
    ynan, xnan = g["mi_dropnan"]["ynan"], g["mi_dropnan"]["xnan"]
    data5 = np.column_stack((x3, y, cond[:, 0]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert eq(eb.pairwise_mi(x3), g["pairwise"]["out"])
        assert eq(eb.pairwise_mi(data5, k=4), g["pairwise5"]["out"])
        assert eq(eb.pairwise_corr(data5), g["pairwise5_corr"]["out"])
        assert eq(eb.pairwise_mi(np.column_stack((x3, y)), cond=cond, mask=mask), g["pairwise_cond_mask"]["out"])
        assert eq(eb.pairwise_mi(np.column_stack((x3[:, 0], xd, yd, y)), discrete=[False, True, True, False]),
                  g["pairwise_discrete"]["out"])
        assert eq(eb.pairwise_mi(np.column_stack((xnan, ynan)), drop_nan=True), g["pairwise_dropnan"]["out"])
        assert eq(eb.estimate_entropy(x3), g["ent_cols"]["out"])
        assert eq(eb.estimate_entropy(x3, multidim=True, k=5), g["ent_multidim"]["out"])
        assert eq(eb.estimate_entropy(x3[:, :2], cond=cond), g["ent_cond"]["out"])
        assert eq(eb.estimate_entropy(x3[:, :2], cond=cond[:, 0], multidim=True, mask=mask), g["ent_cond_multidim_mask"]["out"])
        assert eq(eb.estimate_entropy(ynan, drop_nan=True), g["ent_1d_dropnan"]["out"])
        assert eq(eb.estimate_entropy(np.column_stack((xd, yd)), discrete=True), g["ent_discrete"]["out"])
    assert eq(eb.normalize_mi(g["normalize"]["mi"]), g["normalize"]["out"])


def test_docs_known_answers(oracle_backend, golden_api):
    """The 8-digit outputs printed in the reference's docs (tutorial.md:165, :208-210;
    potential-issues.md:68, :117, :168), regenerated from their snippets."""
    g = golden_api
    rng = np.random.default_rng(1234)
    data = rng.multivariate_normal([0, 0], [[1, 0.8], [0.8, 1]], size=800)
    z = rng.normal(0, 1, size=800)
    out = eb.estimate_corr(data[:, 1], np.column_stack((data[:, 0], z)))
    assert eq(out, g["doc_tutorial_165"]["out"]) and np.allclose(out, g["doc_tutorial_165"]["printed"], atol=5e-9)
    rng = np.random.default_rng(1234)
    x = rng.gamma(1.0, 1.0, size=400); y = np.zeros(400); y[1:] = x[0:-1]; y += rng.normal(0, 0.01, size=400)
    out = eb.estimate_corr(y, x, lag=[1, 0, -1])
    assert eq(out, g["doc_tutorial_208"]["out"]) and np.allclose(out, g["doc_tutorial_208"]["printed"], atol=5e-9)
    rng = np.random.default_rng(1234)
    data = rng.multivariate_normal([0.5, 0.5], [[1, 0.8], [0.8, 1]], size=800)
    out = eb.estimate_mi(np.maximum(0, data[:, 1]), np.maximum(0, data[:, 0]), preprocess=False)
    assert out.shape == (1, 1) and out[0, 0] == -np.inf


def test_error_messages_and_types(oracle_backend):
    x = np.arange(50.0)
    y = x ** 2
    cases = [
        (lambda: eb.estimate_mi(y[:40], x), ValueError, "x and y must have same length"),
        (lambda: eb.estimate_mi(y, np.zeros((50, 2, 2))), ValueError, "x must be one- or two-dimensional"),
        (lambda: eb.estimate_mi(np.zeros((50, 2)), x), ValueError, "y must be one-dimensional"),
        (lambda: eb.estimate_mi(y, x, cond=np.zeros(30)), ValueError, "x and cond must have same length"),
        (lambda: eb.estimate_mi(y, x, k=0), ValueError, "k must be greater than zero"),
        (lambda: eb.estimate_mi(y, x, k=3.0), TypeError, "k must be int"),
        (lambda: eb.estimate_mi(y, x, k=50), ValueError, r"k must be smaller than number of observations \(after lag and mask\)"),
        (lambda: eb.estimate_mi(y, x, lag=50), ValueError, "lag is too large, no observations left"),
        (lambda: eb.estimate_mi(y, x, mask=np.ones(20, dtype=bool)), ValueError, "mask length does not match input length"),
        (lambda: eb.estimate_mi(y, x, mask=np.ones(50)), TypeError, "mask must contain only booleans"),
        (lambda: eb.estimate_mi(y, x, mask=np.ones((50, 2), dtype=bool)), ValueError, "mask must be one-dimensional"),
        (lambda: eb.estimate_mi(y, np.where(x == 3, np.nan, x)), ValueError, "input contains NaNs"),
        (lambda: eb.estimate_entropy(np.zeros((5, 2, 2))), ValueError, "x must be one- or two-dimensional"),
        (lambda: eb.estimate_entropy(x, cond=np.zeros((50, 2, 2))), ValueError, "cond must be one- or two-dimensional"),
        (lambda: eb.estimate_entropy(x, k=50), ValueError, "k must be smaller"),
        (lambda: eb.pairwise_mi(np.column_stack((x, y)), discrete=[True, False], cond=x), ValueError, "Conditioning is not supported"),
        (lambda: eb.estimate_mi(y, x, lag=1.5), TypeError, None),
    ]
    for fn, exc, msg in cases:
        with pytest.raises(exc, match=msg):
            fn()
    with pytest.raises(ValueError, match="data must be finite"):
        eb.estimate_mi(y, np.where(x == 3, np.inf, x), preprocess=False)


def test_shapes_pandas_and_warnings(oracle_backend):
    import pandas as pd
    rng = np.random.default_rng(3)
    df = pd.DataFrame({"a": rng.normal(size=200), "b": rng.normal(size=200), "c": rng.normal(size=200)})
    out = eb.estimate_mi(df["a"], df[["b", "c"]], lag=[0, 1])
    assert isinstance(out, pd.DataFrame) and list(out.columns) == ["b", "c"] and list(out.index) == [0, 1]
    out = eb.estimate_mi(df["a"], df["b"])
    assert isinstance(out, pd.DataFrame) and list(out.columns) == ["b"]
    pw = eb.pairwise_corr(df)
    assert isinstance(pw, pd.DataFrame) and list(pw.index) == ["a", "b", "c"] and np.isnan(pw.loc["a", "a"])
    assert pw.loc["a", "b"] == pw.loc["b", "a"]
    ent = eb.estimate_entropy(df)
    assert isinstance(ent, pd.DataFrame) and ent.shape == (1, 3)
    assert np.asarray(eb.estimate_entropy(df, multidim=True)).shape == ()
    assert isinstance(eb.normalize_mi(pw), pd.DataFrame)
    assert eb.pairwise_mi(np.arange(10.0)).shape == (1, 1)
    assert eb.estimate_mi(df["a"].values, df["b"].values).shape == (1, 1)
    with pytest.warns(UserWarning, match="normalize=True while at least one variable is discrete"):
        eb.estimate_mi(rng.integers(0, 3, 200), df["b"].values, discrete_y=True, normalize=True)
    with pytest.warns(UserWarning, match="takes only a single value"):
        eb.estimate_mi(df["a"].values, np.zeros(200) + 1e-30 * np.arange(200))
    with pytest.warns(UserWarning, match="relatively many unique values"):
        eb.estimate_mi(np.arange(200), df["b"].values, discrete_y=True)


def test_callbacks_threads_and_determinism(oracle_backend, monkeypatch):
    from ennemi_b200 import _schedule, _devices
    rng = np.random.default_rng(5)
    x = rng.normal(size=(300, 3)); y = x[:, 0] + rng.normal(size=300)
    seen, threads = [], set()
    lock = threading.Lock()

    def cb(var, lag):
        with lock:
            seen.append((var, int(lag))); threads.add(threading.current_thread().name)

    monkeypatch.setattr(_devices, "visible", lambda: [0, 1])       # pretend two GPUs: 4 worker threads
    monkeypatch.setattr(_schedule, "INLINE_BUDGET_S", 0.0)
    a = eb.estimate_mi(y, x, lag=[0, 1, -1], callback=cb)
    assert sorted(seen) == sorted((v, l) for l in (0, 1, -1) for v in range(3))
    assert any(t.startswith("ennemi-b200-work") for t in threads)
    seen.clear()
    b = eb.estimate_mi(y, x, lag=[0, 1, -1], max_threads=1, callback=cb)
    assert len(seen) == 9 and eq(a, b)                              # same bits whatever the scheduling
    pairs = []
    c = eb.pairwise_mi(x, callback=lambda i, j: pairs.append((i, j)))
    assert sorted(pairs) == [(0, 1), (0, 2), (1, 2)] and eq(c, c.T)

    def boom(t):
        raise ValueError("from a worker")
    with pytest.raises(ValueError, match="from a worker"):
        _schedule.run_tasks(boom, [1, 2, 3, 4], None, 1.0, lambda i: None)


def test_seam_signatures_match_reference(oracle_backend):
    """``ennemi/_driver.py:18-21`` imports these names; positional (x, y[, cond], k)."""
    from ennemi_b200 import _estimators as e
    rng = np.random.default_rng(9)
    x = rng.normal(size=400); y = x + rng.normal(size=400); z = rng.normal(size=(400, 2)); d = rng.integers(0, 3, 400)
    import oracle
    assert e._estimate_single_mi(x, y, 3) == oracle.ksg_mi(x, y, 3, backend="scipy")["value"]
    assert e._estimate_conditional_mi(x, y, z, 3) == oracle.conditional_mi(x, y, z, 3, backend="scipy")["value"]
    assert e._estimate_conditional_mi(x, y, z[:, 0], k=2) == oracle.conditional_mi(x, y, z[:, 0], 2, backend="scipy")["value"]
    assert e._estimate_semidiscrete_mi(x, d, 3) == oracle.semidiscrete_mi(x, d, 3, backend="scipy")["value"]
    assert e._estimate_conditional_semidiscrete_mi(x, d, z, 3) == oracle.conditional_semidiscrete_mi(x, d, z, 3, backend="scipy")["value"]
    assert e._estimate_single_entropy(z, 3) == oracle.knn_entropy(z, 3, backend="scipy")["value"]
    assert e._estimate_single_entropy(x[::2], 3) == oracle.knn_entropy(x[::2], 3, backend="scipy")["value"]   # strided view
    assert np.isinf(e._psi(np.array([1, 0])))


def test_device_column_path_gives_identical_results(oracle_backend, golden_api, monkeypatch):
    """Lag sweeps / pairwise on cached device columns (rescaling done by the device's prep kernel,
    emulated here with the same IEEE operations) must reproduce the reference bit for bit, including
    cond_lag windows, the noise draw order and constant-data handling."""
    from ennemi_b200 import api
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 0)
    g = golden_api
    x3, y, cond, lags = (g["inputs"][k] for k in ("x3", "y", "cond", "lags"))
    ynan, xnan = g["mi_dropnan"]["ynan"], g["mi_dropnan"]["xnan"]
    data5 = np.column_stack((x3, y, cond[:, 0]))
    puts = []
    real_put = oracle_backend.cache_put
    monkeypatch.setattr(eb.api._columns._native, "cache_put", lambda key, col, dev=0: (puts.append(key), real_put(key, col, dev))[1])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert eq(eb.estimate_mi(y, x3, lags), g["mi_lags"]["out"])
        assert len(puts) >= 3 and oracle_backend.block_puts == 1         # the fast path really ran (x3 as one block)
        assert eq(eb.estimate_mi(y, x3, lags, k=5, preprocess=False), g["mi_lags_k5_nopre"]["out"])
        assert eq(eb.estimate_mi(y, x3[:, :2], lags, cond=cond), g["mi_cond"]["out"])
        assert eq(eb.estimate_mi(y, x3[:, :2], lags, cond=cond, cond_lag=1), g["mi_cond_lag1"]["out"])
        assert eq(eb.estimate_mi(y, x3[:, :2], lags, cond=cond, cond_lag=g["mi_cond_lag2d"]["cond_lag"]), g["mi_cond_lag2d"]["out"])
        assert eq(eb.estimate_corr(y, x3, lags), g["corr_lags"]["out"])
        assert eq(eb.pairwise_mi(x3), g["pairwise"]["out"])
        assert eq(eb.pairwise_mi(data5, k=4), g["pairwise5"]["out"])
        assert eq(eb.pairwise_corr(data5), g["pairwise5_corr"]["out"])
        # NaNs present: drop_nan is not a no-op -> general path; without drop_nan -> the reference's error
        assert eq(eb.estimate_mi(ynan, xnan, [0, 1], drop_nan=True), g["mi_dropnan"]["out"])
    with pytest.raises(ValueError, match="input contains NaNs"):
        eb.estimate_mi(ynan, x3)
    with pytest.raises(ValueError, match="data must be finite"):
        eb.estimate_mi(y, np.where(np.arange(600) == 7, np.inf, x3[:, 0]), preprocess=False)
    with pytest.raises(TypeError):
        eb.estimate_mi(y, x3[:, 0], lag=1.5)
    with pytest.raises(ValueError, match="k must be smaller"):
        eb.estimate_mi(y[:5], x3[:5, 0], k=5)
    with pytest.warns(UserWarning, match="takes only a single value"):
        out = eb.estimate_mi(y, np.column_stack((np.full(600, 2.0), x3[:, 0])), preprocess=True)
    # a constant x consumes no noise draw: the second variable must still match the plain call bit for bit
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)      # the general host path must agree bit for bit
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert eq(out, eb.estimate_mi(y, np.column_stack((np.full(600, 2.0), x3[:, 0])), preprocess=True))
    noise_keys = {entry[0] for entry in eb.api._columns.NoiseBank._keys.values()}
    assert all(k[1] in noise_keys for k in oracle_backend.cache)          # only noise vectors outlive a call


def test_window_stats_bits():
    from ennemi_b200._columns import window_stats
    rng = np.random.default_rng(1)
    for n in (5, 8, 127, 128, 129, 1000, 4097, 100_003):
        a = rng.normal(3.0, 2.5, size=(n, 3))
        for view in (np.ascontiguousarray(a[:, 0]), a[:, 1], a[::2, 2], a[3:, 0]):
            if len(view):
                m, s = window_stats(view)
                assert m == view.mean() and s == view.std()


def test_class_ids_match_np_unique():
    """The counting path for integer / boolean labels numbers classes exactly as np.unique does."""
    from ennemi_b200._estimators import _classes
    rng = np.random.default_rng(0)
    for y in (rng.integers(0, 16, 5000), rng.integers(-5, 3, 1000), rng.integers(0, 2, 100).astype(bool), np.array([7, 7, 7]),
              rng.integers(0, 200, 5000).astype(np.uint8), np.array(["a", "b", "a"]), rng.normal(size=50),
              np.array([0, 2 ** 40, 5], dtype=np.int64)):
        labels, inv, cnt = np.unique(y, return_inverse=True, return_counts=True)
        c, n, sz = _classes(y)
        assert n == len(labels) and np.array_equal(c, np.ravel(inv)) and np.array_equal(sz, cnt)


def test_one_call_statistics_path(oracle_backend, golden_api, monkeypatch):
    """A one-task call leaves the window statistics to the library call (FLAG_DEVICE_STATS, descriptors
    with mean = NaN); a constant window comes back as ConstantWindow and is repeated on the path that
    knows std, which warns exactly like the reference."""
    from ennemi_b200 import api, _columns, _native
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 0)
    monkeypatch.setattr(_columns, "DEVICE_STATS_MIN_ROWS", 0)
    g = golden_api
    x3, y, cond = (g["inputs"][k] for k in ("x3", "y", "cond"))
    seen = []
    real = oracle_backend.ksg_mi_cols
    monkeypatch.setattr(_columns._native, "ksg_mi_cols",
                        lambda descs, n, k, dev=0, flags=0: (seen.append((flags, [d.mean for d in descs])), real(descs, n, k, dev, flags))[1])
    one = eb.estimate_mi(y, x3[:, 0], lag=1)
    assert seen and seen[-1][0] & _native.FLAG_DEVICE_STATS and all(m != m for m in seen[-1][1])
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    assert eq(one, eb.estimate_mi(y, x3[:, 0], lag=1))
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 0)
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    host = eb.estimate_mi(y, x3[:, 1], cond=cond)
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 0)
    assert eq(eb.estimate_mi(y, x3[:, 1], cond=cond), host)
    # constant x: first attempt raises ConstantWindow inside, the repeat warns and agrees with the host path
    const = np.full(600, 2.0)
    n_calls = len(seen)
    with pytest.warns(UserWarning, match="takes only a single value"):
        dev_out = eb.estimate_mi(const, x3[:, 0])
    assert len(seen) == n_calls + 2 and not (seen[-1][0] & _native.FLAG_DEVICE_STATS)
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    with pytest.warns(UserWarning, match="takes only a single value"):
        assert eq(dev_out, eb.estimate_mi(const, x3[:, 0]))


def test_block_upload_path(oracle_backend, monkeypatch):
    """Row-major (n, nvar) arrays are registered as one block per call: one upload per device instead of one
    strided gather per column, whole-column statistics in one call, descriptors memoised per (variable, role) —
    and results identical to the general host path, warnings included."""
    from ennemi_b200 import api, _columns, _native
    rng = np.random.default_rng(5)
    data = rng.normal(size=(400, 6)) @ rng.normal(size=(6, 6))
    cond = rng.normal(size=(400, 2)) + data[:, :2]
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    host_pw = eb.pairwise_mi(data)
    host_pw_c = eb.pairwise_mi(data, cond=cond)
    host_mi = eb.estimate_mi(data[:, 0], data[:, 1:], lag=[0, 2, -1])
    host_np = eb.pairwise_mi(data, preprocess=False)
    wide = np.zeros((400, 40)); wide[:, :6] = data
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 0)
    for stats_rows in (0, 10 ** 9):                 # whole-column statistics on the "device" / lazily on the host
        monkeypatch.setattr(_columns, "DEVICE_STATS_MIN_ROWS", stats_rows)
        single_puts = []
        real_put = oracle_backend.cache_put
        monkeypatch.setattr(_columns._native, "cache_put",
                            lambda key, col, dev=0: (single_puts.append(key), real_put(key, col, dev))[1])
        oracle_backend.block_puts = 0
        assert eq(eb.pairwise_mi(data), host_pw)
        assert oracle_backend.block_puts == 1 and len(single_puts) <= 2          # only noise vectors go one by one
        assert eq(eb.pairwise_mi(data, cond=cond), host_pw_c)
        assert eq(eb.estimate_mi(data[:, 0], data[:, 1:], lag=[0, 2, -1]), host_mi)      # a strided block (ld = 6, 5 columns)
        assert eq(eb.pairwise_mi(data, preprocess=False), host_np)
        assert eq(eb.pairwise_mi(np.asfortranarray(data)), host_pw)                     # contiguous columns: no block
        before = oracle_backend.block_puts
        assert eq(eb.pairwise_mi(wide[:, :6]), host_pw)                                  # sparse view: column by column
        assert oracle_backend.block_puts == before
    # layouts
    assert _native.block_layout(data) == 6 and _native.block_layout(data[:, 1:]) == 6
    assert _native.block_layout(data[:, ::2]) is None and _native.block_layout(np.asfortranarray(data)) is None
    assert _native.block_layout(data.astype(np.float32)) is None and _native.block_layout(data[:, :1]) is None
    # a constant column warns for every task that touches it, on both paths, and is never memoised
    const = data.copy(); const[:, 2] = 1.5
    with warnings.catch_warnings(record=True) as w_dev:
        warnings.simplefilter("always")
        out_dev = eb.pairwise_mi(const)
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    with warnings.catch_warnings(record=True) as w_host:
        warnings.simplefilter("always")
        out_host = eb.pairwise_mi(const)
    assert eq(out_dev, out_host) and len(w_dev) == len(w_host) == 5


def test_noise_stream_skip_keeps_the_draw_sequence():
    """Skipping a draw (the caller already holds its values: memoised descriptors) must leave the stream exactly
    where drawing it would have — the reference's noise depends only on the sequence of draw shapes
    (``ennemi/_driver.py:874-899``)."""
    from ennemi_b200 import _align
    ref = np.random.default_rng(_align.NOISE_SEED)
    a, b, c = (ref.normal(0.0, _align.NOISE_SCALE, sh) for sh in ((97,), (97,), (97, 3)))
    s1 = _align._NoiseStream()
    s1.skip((97,))
    assert np.array_equal(s1.normal((97,)), b) and np.array_equal(s1.normal((97, 3)), c)
    s2 = _align._NoiseStream()
    assert np.array_equal(s2.normal((97,)), a)
    s2.skip((97,))
    assert np.array_equal(s2.normal((97, 3)), c)
    _align._NoiseStream._cache.clear()                  # nothing memoised: the generator has to be replayed
    s3 = _align._NoiseStream()
    s3.skip((97,)); s3.skip((97,))
    assert np.array_equal(s3.normal((97, 3)), c)


def test_conditional_entropy_on_device_columns(oracle_backend, golden_api, monkeypatch):
    """H(X | C) on device-resident columns: X and C uploaded once each whatever the number of
    variables, the columns of C named in both terms; same values as the host route (which stacks and copies C per
    term) and as the fixtures of the unmodified reference; the reference's errors in the reference's words."""
    from ennemi_b200 import api
    g = golden_api
    rng = np.random.default_rng(4)
    n = 400
    c = rng.normal(size=(n, 2))
    x = np.column_stack((c[:, 0] + rng.normal(size=n), rng.normal(size=n), c[:, 1] - rng.normal(size=n)))
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    host = eb.estimate_entropy(x, cond=c)
    host_md = eb.estimate_entropy(x, cond=c, multidim=True)
    host_1d = eb.estimate_entropy(x[:, 0], cond=c[:, 0], k=5)
    puts0 = getattr(oracle_backend, "block_puts", 0)
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 0)
    assert eq(eb.estimate_entropy(x, cond=c), host)
    assert getattr(oracle_backend, "block_puts", 0) == puts0 + 2           # one upload of X, one of C, for three variables
    assert eq(eb.estimate_entropy(x, cond=c, multidim=True), host_md)
    assert eq(eb.estimate_entropy(x[:, 0], cond=c[:, 0], k=5), host_1d)
    x3, cond = g["inputs"]["x3"], g["inputs"]["cond"]
    assert eq(eb.estimate_entropy(x3[:, :2], cond=cond), g["ent_cond"]["out"])          # fixture of the unmodified reference
    # a multi-column variable without a condition: the block goes up as it is, no host transposition
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    host_block = eb.estimate_entropy(x, multidim=True, k=4)
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 0)
    puts1 = oracle_backend.block_puts
    assert eq(eb.estimate_entropy(x, multidim=True, k=4), host_block) and oracle_backend.block_puts == puts1 + 1
    assert eq(eb.estimate_entropy(g["inputs"]["x3"], multidim=True, k=5), g["ent_multidim"]["out"])
    # separate variables (columns of a 2-D x): one block upload, one estimate per cached column
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 10 ** 9)
    host_cols = eb.estimate_entropy(x, k=2)
    monkeypatch.setattr(api, "DEVICE_COLUMNS_MIN_ROWS", 0)
    puts2 = oracle_backend.block_puts
    assert eq(eb.estimate_entropy(x, k=2), host_cols) and oracle_backend.block_puts == puts2 + 1
    assert eq(eb.estimate_entropy(g["inputs"]["x3"]), g["ent_cols"]["out"])
    assert not oracle_backend.cache                                        # the store dropped its columns
    xn = x.copy(); xn[3, 1] = np.nan
    with pytest.raises(ValueError, match="input contains NaNs"):
        eb.estimate_entropy(xn)
    with pytest.raises(ValueError, match="input contains NaNs"):
        eb.estimate_entropy(xn, multidim=True)
    with pytest.raises(ValueError, match="input contains NaNs"):
        eb.estimate_entropy(xn, cond=c)
    with pytest.raises(ValueError, match="k must be smaller"):
        eb.estimate_entropy(x[:3], cond=c[:3], k=3)
    # rows dropped per variable, a mask, discrete data: the host route (identical to the reference's)
    assert eq(eb.estimate_entropy(xn, cond=c, drop_nan=True)[[0, 2]], host[[0, 2]])
