"""The three-level grid (k-NN entropy and Frenzel-Pompe CMI in three and more dimensions, ennemi_b200/csrc/eb2_ksg2.cu)
on a real B200, through the C ABI.  Same bar as everywhere: k-th distances and every count BIT-EXACT, estimates within
1e-10 - against the fixtures captured from the unmodified reference (forced onto small inputs), against the library's
brute-force kernels on stress inputs, and against the oracle's SciPy calls."""
import numpy as np
import pytest

from ennemi_b200 import _native as nat

pytestmark = pytest.mark.gpu

TOL = 1e-10
BRUTE = nat.FLAG_NO_PRUNE | nat.FLAG_BRUTE_COUNT


def close(a, b):
    return a == b or (np.isnan(a) and np.isnan(b)) or abs(a - b) <= TOL


@pytest.fixture(autouse=True)
def small_inputs_take_the_grid(monkeypatch):
    monkeypatch.setenv("EB2_G3_MIN", "2")       # (default: 3-D and 4-D entropy from 200,000 rows on)
    monkeypatch.setenv("EB2_G3_CMI", "1")       # (the Frenzel-Pompe variant is opt-in: slower than the general path)


def test_reference_fixtures_through_the_grid(golden_estimators):
    seen = 0
    for name, c in golden_estimators.items():
        if not name.startswith(("ent", "cmi")):
            continue
        k = int(c["k"])
        if name.startswith("ent") and np.asarray(c["x"]).ndim == 2 and 3 <= c["x"].shape[1] <= 8 and k <= 7:
            v, d = nat.entropy(nat.pack_coords([c["x"]]), k, details=True)
            assert nat.last_pipeline() == 2, name
            assert np.array_equal(d["dist"], c["dist"]), name
        elif name.startswith("cmi") and 2 <= np.column_stack((c["z"],)).shape[1] <= 6 and k <= 7:
            v, d = nat.cmi(nat.pack_coords([c["x"], c["y"], c["z"]]), k, details=True)
            assert nat.last_pipeline() == 2, name
            for key in ("eps", "nxz", "nyz", "nz"):
                assert np.array_equal(d[key], c[key]), (name, key)
        else:
            continue
        assert close(v, float(c["value"])), name
        seen += 1
    assert seen >= 2


@pytest.mark.parametrize("dim,k", [(3, 1), (3, 3), (4, 5), (5, 3), (7, 6), (8, 2)])
def test_entropy_against_brute_force(dim, k):
    rng = np.random.default_rng(dim * 10 + k)
    x = rng.normal(size=(25_000, dim)) @ rng.normal(size=(dim, dim))
    co = nat.pack_coords([x])
    v, d = nat.entropy(co, k, details=True)
    assert nat.last_pipeline() == 2
    vb, b = nat.entropy(co, k, flags=BRUTE, details=True)
    assert np.array_equal(d["dist"], b["dist"])
    assert close(v, vb)


@pytest.mark.parametrize("c,k", [(2, 3), (3, 1), (3, 3), (4, 5), (6, 2)])
def test_cmi_against_brute_force(c, k):
    rng = np.random.default_rng(c * 10 + k)
    n = 25_000
    z = rng.normal(size=(n, c)); x = rng.normal(size=n) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=n)
    co = nat.pack_coords([x, y, z])
    v, d = nat.cmi(co, k, details=True)
    assert nat.last_pipeline() == 2
    vb, b = nat.cmi(co, k, flags=BRUTE, details=True)
    for key in ("eps", "nxz", "nyz", "nz"):
        assert np.array_equal(d[key], b[key]), key
    assert close(v, vb)


def _stress():
    rng = np.random.default_rng(17)
    yield "student-t2", rng.standard_t(2, size=(30_000, 4))
    yield "ties", np.round(rng.normal(size=(20_000, 3)), 1)
    yield "duplicates", np.repeat(rng.normal(size=(5_000, 4)), 4, axis=0)
    o = rng.normal(size=(30_000, 3)); o[3] = (50.0, -40.0, 30.0); o[4] = (0.0, 0.0, 60.0)
    yield "outliers", o
    yield "thin-sheet", np.column_stack((rng.uniform(size=20_000), rng.uniform(size=20_000), 1e-6 * rng.normal(size=20_000)))
    yield "constant-coordinate", np.column_stack((rng.normal(size=(15_000, 2)), np.full(15_000, 3.0), rng.normal(size=15_000)))
    yield "huge-offset", rng.normal(size=(20_000, 3)) * 1e12 + 1e15
    # coordinate 0 (the bucket coordinate) tied: five values fit the buckets, a constant column overflows them and the
    # call is repeated on the general path
    yield "tied-bucket-coordinate", np.column_stack((rng.integers(0, 5, 30_000).astype(float), rng.normal(size=(30_000, 3))))
    yield "constant-bucket-coordinate", np.column_stack((np.full(20_000, -1.5), rng.normal(size=(20_000, 2))))


@pytest.mark.parametrize("case", list(_stress()), ids=lambda c: c[0])
def test_stress_distributions(case):
    name, x = case
    co = nat.pack_coords([x])
    v, d = nat.entropy(co, 3, details=True)
    vb, b = nat.entropy(co, 3, flags=BRUTE, details=True)
    assert np.array_equal(d["dist"], b["dist"]), name
    assert close(v, vb) or (np.isinf(v) and v == vb)
    # the same rows as a condition (x, y derived from them): all three counts
    n = len(x)
    rng = np.random.default_rng(1)
    xx = x[:, 0] + rng.normal(size=n); yy = x[:, 1] - xx + rng.normal(size=n)
    cz = nat.pack_coords([xx, yy, x])
    v, d = nat.cmi(cz, 3, details=True)
    vb, b = nat.cmi(cz, 3, flags=BRUTE, details=True)
    for key in ("eps", "nxz", "nyz", "nz"):
        assert np.array_equal(d[key], b[key]), (name, key)
    assert close(v, vb)


def test_against_the_reference_calls():
    import oracle
    rng = np.random.default_rng(23)
    n = 60_000
    z = rng.normal(size=(n, 3)); x = rng.normal(size=n) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=n)
    want = oracle.conditional_mi(x, y, z, 3, backend="scipy")
    v, got = nat.cmi(nat.pack_coords([x, y, z]), 3, details=True)
    assert nat.last_pipeline() == 2
    for key in ("eps", "nxz", "nyz", "nz"):
        assert np.array_equal(got[key], want[key]), key
    assert close(v, want["value"])
    d4 = rng.normal(size=(n, 4)) @ rng.normal(size=(4, 4))
    want = oracle.knn_entropy(d4, 5, backend="scipy")
    v, got = nat.entropy(nat.pack_coords([d4]), 5, details=True)
    assert np.array_equal(got["dist"], want["dist"]) and close(v, want["value"])


def test_default_dispatch_and_full_size_entropy(monkeypatch):
    """Without the test knobs: 3-D and 4-D entropy take the grid from 200,000 rows on, smaller inputs, wider spaces and
    Frenzel-Pompe stay on the general path; BASELINE.json configs[4] (4-D, N = 500,000, k = 5) against the reference's
    SciPy calls, every row."""
    import oracle
    monkeypatch.delenv("EB2_G3_MIN")
    monkeypatch.delenv("EB2_G3_CMI")
    rng = np.random.default_rng(0)
    n = 500_000
    cov = np.array([[1.0, 0.5, 0.2, 0.1], [0.5, 1.0, 0.3, 0.0], [0.2, 0.3, 1.0, -0.4], [0.1, 0.0, -0.4, 1.0]])
    d4 = rng.multivariate_normal(np.zeros(4), cov, size=n)
    want = oracle.knn_entropy(d4, 5, backend="scipy")
    v, got = nat.entropy(nat.pack_coords([d4]), 5, details=True)
    assert nat.last_pipeline() == 2
    assert np.array_equal(got["dist"], want["dist"]) and close(v, want["value"])
    t = rng.standard_t(2, size=(200_000, 3))                  # heavy tails at the smallest size the grid takes
    want = oracle.knn_entropy(t, 3, backend="scipy")
    v, got = nat.entropy(nat.pack_coords([t]), 3, details=True)
    assert nat.last_pipeline() == 2
    assert np.array_equal(got["dist"], want["dist"]) and close(v, want["value"])
    for x in (d4[:150_000], rng.normal(size=(200_000, 5))):
        nat.entropy(nat.pack_coords([x]), 3)
        assert nat.last_pipeline() == 0
    z = d4[:200_000, :2]
    nat.cmi(nat.pack_coords([d4[:200_000, 2], d4[:200_000, 3], z]), 3)
    assert nat.last_pipeline() == 0
