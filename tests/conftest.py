import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def has_gpu() -> bool:
    try:
        from ennemi_b200 import _native
        return _native.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: nothing to do here.  Without `-m`,
    # GPU tests are skipped when no device is present so that a plain `pytest tests` works anywhere.
    if config.getoption("-m"):
        return
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class OracleBackend:
    """Stands in for ``ennemi_b200._native`` in CPU tests of the HOST logic: same function
    signatures, answers computed by the CPU oracle.  Never used by the product."""

    PATCHED = ("ksg_mi", "cmi", "ross_mi", "ross_cmi", "entropy", "psi",
               "cache_put", "cache_drop", "ksg_mi_cols", "cmi_cols", "mi_cols_batch", "cache_stats",
               "cache_put_block", "cache_stats_many", "ksg_mi_pairs", "entropy_cols")

    def __init__(self, backend="scipy"):
        import oracle
        self.o = oracle
        self.backend = backend
        self.cache = {}

    @staticmethod
    def _rows(coords):
        return [np.ascontiguousarray(r) for r in coords]

    def ksg_mi(self, coords, k, dev=0, flags=0, details=False):
        r = self.o.ksg_mi(coords[0], coords[1], k, backend=self.backend)
        return (r["value"], r) if details else r["value"]

    def cmi(self, coords, k, dev=0, flags=0, details=False):
        r = self.o.conditional_mi(coords[0], coords[1], coords[2:].T, k, backend=self.backend)
        return (r["value"], r) if details else r["value"]

    def ross_mi(self, coords, cls, ncls, k, dev=0, flags=0, details=False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = self.o.semidiscrete_mi(coords[0], cls, k, backend=self.backend)
        return (r["value"], r) if details else r["value"]

    def ross_cmi(self, coords, cls, ncls, k, dev=0, flags=0, details=False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = self.o.conditional_semidiscrete_mi(coords[0], cls, coords[1:].T, k, backend=self.backend)
        return (r["value"], r) if details else r["value"]

    def entropy(self, coords, k, dev=0, flags=0, details=False):
        r = self.o.knn_entropy(coords.T, k, backend=self.backend)
        return (r["value"], r) if details else r["value"]

    def psi(self, counts, dev=0):
        return np.asarray(self.o.psi(np.asarray(counts)), dtype=np.float64)

    # ---- device column cache, emulated with numpy (same three IEEE operations as prep_kernel)
    def cache_put(self, key, column, dev=0):
        self.cache[(dev & 0xFF, key)] = np.array(column, dtype=np.float64).ravel()

    def cache_put_block(self, keys, block, dev=0):
        from ennemi_b200 import _native
        assert _native.block_layout(block) is not None and len(keys) == block.shape[1]
        self.block_puts = getattr(self, "block_puts", 0) + 1
        for j, key in enumerate(keys):
            self.cache[(dev & 0xFF, key)] = np.array(block[:, j], dtype=np.float64)

    def cache_stats_many(self, keys, offs, n, stride=1, dev=0):
        pairs = [self.cache_stats(k, o, n, stride, dev) for k, o in zip(keys, offs)]
        return np.array([p[0] for p in pairs]), np.array([p[1] for p in pairs])

    def cache_drop(self, key, dev=0):
        self.cache.pop((dev & 0xFF, key), None)

    def cache_stats(self, key, off, n, stride=1, dev=0):
        v = self.cache[(dev & 0xFF, key)][off: off + (n - 1) * stride + 1: stride]
        return float(v.mean()), float(v.std())

    def _gather(self, descs, n, dev, flags=0):
        from ennemi_b200 import _native
        rows = []
        for d in descs:
            col = self.cache[(dev & 0xFF, d.key)]
            v = col[d.off: d.off + (n - 1) * d.stride + 1: d.stride].copy()
            if np.isnan(v).any():
                raise _native.NonFiniteInput("data must be finite, check for nan or inf values", True)
            mean, std = d.mean, d.std
            if (flags & _native.FLAG_DEVICE_STATS) and std != 0.0 and mean != mean:
                mean, std = v.mean(), v.std()                 # EB2_FLAG_DEVICE_STATS: statistics inside the call
                if abs(std) < 1e-20:
                    raise _native.ConstantWindow()
            if std != 0.0:
                v = (v - mean) / std
                if d.nkey:
                    noise = self.cache[(dev & 0xFF, d.nkey)]
                    v = v + noise[d.noff: d.noff + (n - 1) * d.nstride + 1: d.nstride]
            if not np.isfinite(v).all():
                raise _native.NonFiniteInput("data must be finite, check for nan or inf values", False)
            rows.append(v)
        return rows

    def ksg_mi_cols(self, descs, n, k, dev=0, flags=0):
        x, y = self._gather(descs, n, dev, flags)
        return self.o.ksg_mi(x, y, k, backend=self.backend)["value"]

    def mi_cols_batch(self, tasks, n, k, dev=0, flags=0):
        from ennemi_b200 import _native
        values, status = np.full(len(tasks), np.nan), np.zeros(len(tasks), dtype=np.int32)
        for t, descs in enumerate(tasks):
            try:
                values[t] = self.ksg_mi_cols(descs, n, k, dev) if len(descs) == 2 else self.cmi_cols(descs, n, k, dev)
            except _native.NonFiniteInput as e:
                status[t] = _native.ERR_NONFINITE | ((1 if e.nan else 2) << 8)
        return values, status

    def ksg_mi_pairs(self, cols, pairs, n, k, dev=0, flags=0):
        return self.mi_cols_batch([[cols[i], cols[j]] for i, j in np.asarray(pairs).reshape(-1, 2)], n, k, dev, flags)

    def entropy_cols(self, descs, n, k, dev=0, flags=0):
        rows = self._gather(descs, n, dev, flags)
        return self.o.knn_entropy(np.column_stack(rows), k, backend=self.backend)["value"]

    def cmi_cols(self, descs, n, k, dev=0, flags=0):
        rows = self._gather(descs, n, dev, flags)
        return self.o.conditional_mi(rows[0], rows[1], np.column_stack(rows[2:]), k, backend=self.backend)["value"]


@pytest.fixture
def oracle_backend(monkeypatch):
    """Routes the estimator seam to the CPU oracle so that host logic can be tested without a GPU."""
    from ennemi_b200 import _native, _devices
    fake = OracleBackend()
    for name in OracleBackend.PATCHED:
        monkeypatch.setattr(_native, name, getattr(fake, name))
    monkeypatch.setattr(_native, "device_count", lambda: 1)
    monkeypatch.setattr(_devices, "visible", lambda: [0])
    from ennemi_b200 import _columns
    monkeypatch.setattr(_columns.NoiseBank, "_on_device", {})      # a fresh fake device holds no noise vectors yet
    return fake


@pytest.fixture(scope="session")
def golden_estimators():
    g = np.load(os.path.join(ROOT, "tests", "golden", "estimators.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    return {n: {k.split("/")[1]: g[k] for k in g.files if k.startswith(n + "/")} for n in names}


@pytest.fixture(scope="session")
def golden_api():
    g = np.load(os.path.join(ROOT, "tests", "golden", "api.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    return {n: {k.split("/")[1]: g[k] for k in g.files if k.startswith(n + "/")} for n in names}
