"""Generates ``tests/golden/*.npz`` by running the UNMODIFIED reference from ``/root/reference``.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Two fixture files are written:

``estimators.npz``  per-estimator cases: the inputs, the value returned by the reference's own
    private estimator, and the intermediate arrays (k-th-neighbour distances and neighbour
    counts) captured from the reference run by wrapping the ``cKDTree`` name inside
    ``ennemi._entropy_estimators`` with a recording subclass — i.e. these are the arrays the
    reference itself computed, not a re-derivation.
``api.npz``  public-API cases (lags, cond, cond_lag, mask, discrete, preprocess, normalize,
    drop_nan, pairwise, entropy) with inputs and the reference's outputs, plus the
    known-answer outputs printed in the reference's docs (tutorial.md:165, :208-210,
    potential-issues.md:68, :117, :168).
"""
import os
import sys
import warnings

import numpy as np

REF = "/root/reference"
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import ennemi  # noqa: E402
from ennemi import _entropy_estimators as ee  # noqa: E402
from scipy.spatial import cKDTree as _RealTree  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

_log = []


class RecordingTree(_RealTree):
    """cKDTree that remembers what ``query`` / ``query_ball_point`` returned."""

    def query(self, *a, **kw):
        res = super().query(*a, **kw)
        _log.append(("query", self.n, np.asarray(a[0]).shape, res[0].ravel().copy()))
        return res

    def query_ball_point(self, *a, **kw):
        res = super().query_ball_point(*a, **kw)
        _log.append(("count", self.n, np.asarray(a[0]).shape, np.asarray(res, dtype=np.int64).copy()))
        return res


def run_recorded(fn, *args):
    _log.clear()
    ee.cKDTree = RecordingTree
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            value = fn(*args)
    finally:
        ee.cKDTree = _RealTree
    return float(value), list(_log)


def scatter_by_class(y, pieces):
    """Per-class recorded arrays -> one array in original row order (classes in np.unique order)."""
    y = np.asarray(y)
    out = np.empty(len(y), dtype=pieces[0].dtype)
    for lab, piece in zip(np.unique(y), pieces):
        out[y == lab] = piece
    return out


def estimator_cases():
    g = {}

    def put(name, **arrs):
        for key, val in arrs.items():
            g[f"{name}/{key}"] = np.asarray(val)

    # ---- KSG (a1) -------------------------------------------------------------------
    ksg_specs = {
        "ksg_gauss_k3": (800, 3, "gauss"), "ksg_gauss_k1": (500, 1, "gauss"),
        "ksg_unif_k7": (600, 7, "unif"), "ksg_small_k_max": (40, 39, "gauss"),
        "ksg_dups": (400, 3, "dups"), "ksg_gauss_k20": (700, 20, "gauss"),
        "ksg_big": (5000, 3, "gauss"),
    }
    for i, (name, (n, k, kind)) in enumerate(ksg_specs.items()):
        rng = np.random.default_rng(100 + i)
        if kind == "gauss":
            d = rng.multivariate_normal([0, 0], [[1, 0.6], [0.6, 1]], size=n)
            x, y = d[:, 0].copy(), d[:, 1].copy()
        elif kind == "unif":
            x = rng.uniform(-3, 5, n); y = x ** 2 + rng.uniform(0, 1, n)
        else:  # many exact duplicates -> eps = 0 -> -inf
            x = rng.integers(0, 6, n).astype(float); y = rng.integers(0, 5, n).astype(float)
        val, log = run_recorded(ee._estimate_single_mi, x, y, k)
        put(name, x=x, y=y, k=k, value=val, eps=log[0][3], nx=log[1][3], ny=log[2][3])

    # ---- Frenzel-Pompe (a2) -----------------------------------------------------------
    for i, (name, (n, k, c)) in enumerate({"cmi_c1_k3": (600, 3, 1), "cmi_c3_k2": (500, 2, 3),
                                           "cmi_c2_k5": (900, 5, 2), "cmi_c7_k3": (400, 3, 7)}.items()):
        rng = np.random.default_rng(200 + i)
        z = rng.normal(size=(n, c))
        x = rng.normal(size=n) + z[:, 0]
        y = 0.5 * x + z[:, -1] + rng.normal(size=n)
        zz = z[:, 0] if c == 1 else z
        val, log = run_recorded(ee._estimate_conditional_mi, x, y, zz, k)
        put(name, x=x, y=y, z=z, k=k, value=val, eps=log[0][3], nxz=log[1][3], nyz=log[2][3], nz=log[3][3])
    # duplicates in every space: psi(0) in all three terms -> inf + inf - inf = nan
    rng = np.random.default_rng(250)
    x = rng.integers(0, 3, 300).astype(float); y = rng.integers(0, 3, 300).astype(float)
    z = rng.integers(0, 2, (300, 1)).astype(float)
    val, log = run_recorded(ee._estimate_conditional_mi, x, y, z, 3)
    put("cmi_dups", x=x, y=y, z=z, k=3, value=val, eps=log[0][3], nxz=log[1][3], nyz=log[2][3], nz=log[3][3])

    # ---- Ross (a3) --------------------------------------------------------------------
    for i, (name, (n, k, ncls)) in enumerate({"ross_4cls_k3": (900, 3, 4), "ross_16cls_k5": (1600, 5, 16),
                                              "ross_tiny_class": (300, 4, 3)}.items()):
        rng = np.random.default_rng(300 + i)
        y = rng.integers(0, ncls, n)
        if name == "ross_tiny_class":
            y[:] = rng.integers(0, 2, n); y[:3] = 2          # class 2 has 3 <= k members -> eps = inf
        x = rng.normal(size=n) + 0.5 * y
        val, log = run_recorded(ee._estimate_semidiscrete_mi, x, y, k)
        eps = scatter_by_class(y, [e[3] for e in log[0::2]])
        nfull = scatter_by_class(y, [e[3] for e in log[1::2]])
        put(name, x=x, y=y, k=k, value=val, eps=eps, n_full=nfull)
    rng = np.random.default_rng(350)
    ys = rng.choice(np.array(["red", "green", "blue"]), 500)
    x = rng.normal(size=500) + (ys == "red") * 1.5
    val, log = run_recorded(ee._estimate_semidiscrete_mi, x, ys, 3)
    put("ross_strings", x=x, y=ys, k=3, value=val, eps=scatter_by_class(ys, [e[3] for e in log[0::2]]),
        n_full=scatter_by_class(ys, [e[3] for e in log[1::2]]))

    # ---- conditional Ross (a4) ----------------------------------------------------------
    for i, (name, (n, k, ncls, c)) in enumerate({"cross_3cls_c2": (700, 3, 3, 2), "cross_5cls_c1": (800, 2, 5, 1)}.items()):
        rng = np.random.default_rng(400 + i)
        y = rng.integers(0, ncls, n)
        z = rng.normal(size=(n, c))
        x = rng.normal(size=n) + 0.7 * y + z[:, 0]
        val, log = run_recorded(ee._estimate_conditional_semidiscrete_mi, x, y, z, k)
        put(name, x=x, y=y, z=z, k=k, value=val,
            eps=scatter_by_class(y, [e[3] for e in log[0::4]]), nxz=scatter_by_class(y, [e[3] for e in log[1::4]]),
            nyz=scatter_by_class(y, [e[3] for e in log[2::4]]), nz=scatter_by_class(y, [e[3] for e in log[3::4]]))

    # ---- k-NN entropy (a5) --------------------------------------------------------------
    cov4 = np.array([[1.0, 0.5, 0.6, -0.2], [0.5, 1.0, 0.7, -0.5], [0.6, 0.7, 2.0, -0.1], [-0.2, -0.5, -0.1, 0.5]])
    for i, (name, (n, k, m)) in enumerate({"ent_1d_k3": (500, 3, 1), "ent_4d_k5": (600, 5, 4), "ent_2d_k1": (300, 1, 2),
                                           "ent_3d_k30": (400, 30, 3)}.items()):
        rng = np.random.default_rng(500 + i)
        if m == 4:
            x = rng.multivariate_normal([0, 0, 0, 0], cov4, size=n)
        elif m == 1:
            x = rng.normal(1.0, 2.0, size=n)
        else:
            x = rng.normal(size=(n, m))
        val, log = run_recorded(ee._estimate_single_entropy, x, k)
        put(name, x=x, k=k, value=val, dist=log[0][3])
    x = np.round(np.random.default_rng(550).normal(size=200), 1)          # duplicates -> log(0) -> -inf
    with np.errstate(divide="ignore"):
        val, log = run_recorded(ee._estimate_single_entropy, x, 3)
    put("ent_dups", x=x, k=3, value=val, dist=log[0][3])

    # ---- digamma (a6) -------------------------------------------------------------------
    n = np.concatenate((np.arange(1, 200), [1000, 12345, 10 ** 6, 2 ** 31, 10 ** 12]))
    put("psi", n=n, value=ee._psi(n), zero=ee._psi(np.array([3, 0, 5])))
    return g


def api_cases():
    g = {}

    def put(name, **arrs):
        for key, val in arrs.items():
            g[f"{name}/{key}"] = np.asarray(val)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # docs known answers -----------------------------------------------------------
        rng = np.random.default_rng(1234)
        data = rng.multivariate_normal([0, 0], [[1, 0.8], [0.8, 1]], size=800)
        z = rng.normal(0, 1, size=800)
        put("doc_tutorial_165", out=ennemi.estimate_corr(data[:, 1], np.column_stack((data[:, 0], z))),
            printed=np.array([[0.79978795, -0.02110195]]))
        rng = np.random.default_rng(1234)
        x = rng.gamma(1.0, 1.0, size=400); y = np.zeros(400); y[1:] = x[0:-1]; y += rng.normal(0, 0.01, size=400)
        put("doc_tutorial_208", out=ennemi.estimate_corr(y, x, lag=[1, 0, -1]),
            printed=np.array([[0.99975754], [-0.04579946], [-0.00918085]]))
        rng = np.random.default_rng(1234)
        data = rng.multivariate_normal([0, 0], [[1, 0.8], [0.8, 1]], size=800)
        put("doc_issues_68", out=ennemi.estimate_mi(np.exp(data[:, 1]), np.exp(5 * data[:, 0])),
            printed=np.array([[0.18815223]]))
        rng = np.random.default_rng(1234)
        data = rng.multivariate_normal([0.5, 0.5], [[1, 0.8], [0.8, 1]], size=800)
        put("doc_issues_117", out=ennemi.estimate_mi(np.maximum(0, data[:, 1]), np.maximum(0, data[:, 0]), preprocess=False),
            printed=np.array([[-np.inf]]))
        rng = np.random.default_rng(1234)
        data = rng.multivariate_normal([0, 0], [[1, 0.8], [0.8, 1]], size=800)
        x = np.concatenate((data[:, 0], data[:, 0] + rng.normal(0, 0.01, size=800), data[:, 0] + rng.normal(0, 0.01, size=800)))
        y = np.concatenate((data[:, 1], data[:, 1] + rng.normal(0, 0.01, size=800), data[:, 1] + rng.normal(0, 0.01, size=800)))
        put("doc_issues_168", out=ennemi.estimate_mi(y, x), printed=np.array([[1.02554819]]))

        # seeded API cases (inputs stored) -----------------------------------------------
        rng = np.random.default_rng(7)
        n = 600
        x3 = rng.normal(size=(n, 3))
        y = 0.7 * np.roll(x3[:, 0], 2) + 0.3 * x3[:, 1] + rng.normal(size=n) * 0.5
        cond = np.column_stack((np.roll(x3[:, 2], 1) + rng.normal(size=n) * 0.3, rng.normal(size=n)))
        mask = rng.random(n) > 0.2
        lags = np.array([0, 1, 2, -1, 3])
        put("inputs", x3=x3, y=y, cond=cond, mask=mask, lags=lags)
        put("mi_lags", out=ennemi.estimate_mi(y, x3, lags))
        put("mi_lags_k5_nopre", out=ennemi.estimate_mi(y, x3, lags, k=5, preprocess=False))
        put("mi_cond", out=ennemi.estimate_mi(y, x3[:, :2], lags, cond=cond))
        put("mi_cond_lag1", out=ennemi.estimate_mi(y, x3[:, :2], lags, cond=cond, cond_lag=1))
        cl2 = np.array([[0, 1], [1, 1], [2, 0], [-1, 0], [1, 3]])
        put("mi_cond_lag2d", cond_lag=cl2, out=ennemi.estimate_mi(y, x3[:, :2], lags, cond=cond, cond_lag=cl2))
        put("mi_mask", out=ennemi.estimate_mi(y, x3, lags, mask=mask))
        put("mi_mask_cond", out=ennemi.estimate_mi(y, x3[:, 0], [0, 2], mask=mask, cond=cond[:, 0]))
        put("corr_lags", out=ennemi.estimate_corr(y, x3, lags))
        ynan = y.copy(); ynan[rng.integers(0, n, 30)] = np.nan
        xnan = x3.copy(); xnan[rng.integers(0, n, 25), 1] = np.nan
        put("mi_dropnan", ynan=ynan, xnan=xnan, out=ennemi.estimate_mi(ynan, xnan, [0, 1], drop_nan=True))
        yd = (y > 0).astype(int) + (y > 1).astype(int)
        put("mi_discrete_y", yd=yd, out=ennemi.estimate_mi(yd, x3, [0, 2], discrete_y=True))
        xd = np.digitize(x3[:, 0], [-1, 0, 1])
        put("mi_discrete_x", xd=xd, out=ennemi.estimate_mi(y, xd, [0, 2], discrete_x=True, k=4))
        put("mi_discrete_y_cond", out=ennemi.estimate_mi(yd, x3[:, :2], [0, 1], discrete_y=True, cond=cond))
        put("mi_discrete_x_cond", out=ennemi.estimate_mi(y, xd, [0, 1], discrete_x=True, cond=cond[:, 0]))
        put("mi_discrete_both", out=ennemi.estimate_mi(yd, xd, [0, 1], discrete_x=True, discrete_y=True))
        put("mi_discrete_both_cond", out=ennemi.estimate_mi(yd, xd, 0, discrete_x=True, discrete_y=True,
                                                            cond=(cond[:, 0] > 0).astype(int)))
        put("pairwise", out=ennemi.pairwise_mi(x3))
        data5 = np.column_stack((x3, y, cond[:, 0]))
        put("pairwise5", out=ennemi.pairwise_mi(data5, k=4))
        put("pairwise5_corr", out=ennemi.pairwise_corr(data5))
        put("pairwise_cond_mask", out=ennemi.pairwise_mi(np.column_stack((x3, y)), cond=cond, mask=mask))
        put("pairwise_discrete", out=ennemi.pairwise_mi(np.column_stack((x3[:, 0], xd, yd, y)), discrete=[False, True, True, False]))
        put("pairwise_dropnan", out=ennemi.pairwise_mi(np.column_stack((xnan, ynan)), drop_nan=True))
        put("ent_cols", out=ennemi.estimate_entropy(x3))
        put("ent_multidim", out=ennemi.estimate_entropy(x3, multidim=True, k=5))
        put("ent_cond", out=ennemi.estimate_entropy(x3[:, :2], cond=cond))
        put("ent_cond_multidim_mask", out=ennemi.estimate_entropy(x3[:, :2], cond=cond[:, 0], multidim=True, mask=mask))
        put("ent_1d_dropnan", out=ennemi.estimate_entropy(ynan, drop_nan=True))
        put("ent_discrete", out=ennemi.estimate_entropy(np.column_stack((xd, yd)), discrete=True))
        put("normalize", mi=np.array([-0.1, 0.0, 0.2, 1.5, 7.0]), out=ennemi.normalize_mi(np.array([-0.1, 0.0, 0.2, 1.5, 7.0])))
    return g


def main():
    os.makedirs(OUT, exist_ok=True)
    est = estimator_cases()
    api = api_cases()
    np.savez_compressed(os.path.join(OUT, "estimators.npz"), **est)
    np.savez_compressed(os.path.join(OUT, "api.npz"), **api)
    for f in ("estimators.npz", "api.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
    for key in sorted(api):
        if key.endswith("/printed"):
            name = key.split("/")[0]
            print(name, np.ravel(api[f"{name}/out"]), "printed", np.ravel(api[key]))


if __name__ == "__main__":
    main()
