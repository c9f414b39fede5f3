"""Parity of the CUDA path with the reference, through the C ABI, on a real B200.

Bar (BASELINE.json north_star): k-th-neighbour distances and every neighbour count BIT-EXACT,
final MI / entropy within 1e-10 absolute.  Checked three ways: (1) against the fixtures captured
from the unmodified reference, (2) against the CPU oracle on fresh seeded inputs at sizes it
finishes in seconds, (3) at BASELINE.json's full sizes through size-independent properties."""
import warnings

import numpy as np
import pytest

from ennemi_b200 import _native as nat

pytestmark = pytest.mark.gpu

TOL = 1e-10                       # absolute tolerance on the final estimate (north_star)
MODES = [0, nat.FLAG_NO_PRUNE, nat.FLAG_BRUTE_COUNT, nat.FLAG_NO_PRUNE | nat.FLAG_BRUTE_COUNT]


def close(a, b):
    return a == b or (np.isnan(a) and np.isnan(b)) or abs(a - b) <= TOL


def classes(y):
    labels, inv = np.unique(y, return_inverse=True)
    return np.ascontiguousarray(inv, dtype=np.int32), len(labels)


def run_gpu(name, c, flags):
    k = int(c["k"])
    if name.startswith("ksg"):
        return nat.ksg_mi(nat.pack_coords([c["x"], c["y"]]), k, flags=flags, details=True)
    if name.startswith("cmi"):
        return nat.cmi(nat.pack_coords([c["x"], c["y"], c["z"]]), k, flags=flags, details=True)
    if name.startswith("ross"):
        cls, ncls = classes(c["y"])
        return nat.ross_mi(nat.pack_coords([c["x"]]), cls, ncls, k, flags=flags, details=True)
    if name.startswith("cross"):
        cls, ncls = classes(c["y"])
        return nat.ross_cmi(nat.pack_coords([c["x"], c["z"]]), cls, ncls, k, flags=flags, details=True)
    if name.startswith("ent"):
        return nat.entropy(nat.pack_coords([c["x"]]), k, flags=flags, details=True)
    raise AssertionError(name)


@pytest.mark.parametrize("flags", MODES)
def test_golden_fixtures_bit_exact(golden_estimators, flags):
    """Every fixture case (all five estimators; k from 1 to n-1; duplicates -> -inf / nan;
    classes smaller than k; string labels) in every kernel mode."""
    for name, c in golden_estimators.items():
        if name == "psi":
            continue
        value, parts = run_gpu(name, c, flags)
        for key, arr in parts.items():
            assert np.array_equal(arr, c[key]), (name, key, int(np.sum(arr != c[key])))
        assert close(value, float(c["value"])), (name, value, float(c["value"]))


def test_device_psi_matches_reference_formula(golden_estimators):
    c = golden_estimators["psi"]
    out = nat.psi(c["n"])
    assert np.max(np.abs(out - c["value"])) <= 4e-15 * np.max(np.abs(c["value"]))
    assert nat.psi(np.array([0, 1, 2]))[0] == np.inf
    assert nat.psi(np.array([1]))[0] == -0.5772156649015331


def _oracle_case(kind, n, k, seed, backend):
    import oracle
    rng = np.random.default_rng(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if kind == "ksg":
            d = rng.multivariate_normal([0, 0], [[1, 0.6], [0.6, 1]], size=n)
            return {"x": d[:, 0], "y": d[:, 1], "k": k}, oracle.ksg_mi(d[:, 0], d[:, 1], k, backend=backend)
        if kind == "cmi":
            z = rng.normal(size=(n, 3)); x = rng.normal(size=n) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=n)
            return {"x": x, "y": y, "z": z, "k": k}, oracle.conditional_mi(x, y, z, k, backend=backend)
        if kind == "ross":
            yd = rng.integers(0, 16, n); x = rng.normal(size=n) + 0.25 * yd
            return {"x": x, "y": yd, "k": k}, oracle.semidiscrete_mi(x, yd, k, backend=backend)
        if kind == "cross":
            yd = rng.integers(0, 5, n); z = rng.normal(size=(n, 2)); x = rng.normal(size=n) + 0.5 * yd + z[:, 0]
            return {"x": x, "y": yd, "z": z, "k": k}, oracle.conditional_semidiscrete_mi(x, yd, z, k, backend=backend)
        if kind == "ent":
            cov = np.array([[1.0, 0.5, 0.6, -0.2], [0.5, 1.0, 0.7, -0.5], [0.6, 0.7, 2.0, -0.1], [-0.2, -0.5, -0.1, 0.5]])
            x = rng.multivariate_normal([0, 0, 0, 0], cov, size=n)
            return {"x": x, "k": k}, oracle.knn_entropy(x, k, backend=backend)
    raise AssertionError(kind)


@pytest.mark.parametrize("kind,n,k", [
    ("ksg", 10_000, 3), ("ksg", 30_000, 1), ("ksg", 20_000, 12), ("ksg", 4_097, 100), ("ksg", 513, 3),
    ("cmi", 20_000, 3), ("cmi", 6_000, 20), ("ross", 50_000, 5), ("ross", 3_000, 40),
    ("cross", 20_000, 3), ("ent", 30_000, 5), ("ent", 5_000, 60),
])
def test_against_oracle_on_fresh_inputs(kind, n, k):
    """BASELINE.json config shapes at oracle-sized N (C all-pairs backend: independent of SciPy),
    including tile-boundary sizes and k beyond the register top-k (heap variant)."""
    c, want = _oracle_case(kind, n, k, seed=n + k, backend="c")
    name = {"ksg": "ksg", "cmi": "cmi", "ross": "ross", "cross": "cross", "ent": "ent"}[kind]
    for flags in (0, nat.FLAG_NO_PRUNE):
        value, parts = run_gpu(name, c, flags)
        for key, arr in parts.items():
            assert np.array_equal(arr, want[key]), (kind, key, flags, int(np.sum(arr != want[key])))
        assert close(value, want["value"]), (kind, flags, value, want["value"])


def test_config1_reference_shape():
    """BASELINE.json configs[0]: bivariate Gaussian rho=0.6, N=10,000, k=3 (the reference's CPU case);
    SciPy backend = the reference's own calls."""
    import oracle
    rng = np.random.default_rng(0)
    d = rng.multivariate_normal([0, 0], [[1, 0.6], [0.6, 1]], size=10_000)
    want = oracle.ksg_mi(d[:, 1], d[:, 0], 3, backend="scipy")
    value, parts = nat.ksg_mi(nat.pack_coords([d[:, 1], d[:, 0]]), 3, details=True)
    assert np.array_equal(parts["eps"], want["eps"]) and np.array_equal(parts["nx"], want["nx"])
    assert np.array_equal(parts["ny"], want["ny"]) and close(value, want["value"])


def test_primitives_and_class_restriction():
    import oracle
    rng = np.random.default_rng(11)
    pts = rng.normal(size=(3_000, 3))
    cls = rng.integers(0, 7, 3_000).astype(np.int32)
    coords = nat.pack_coords([pts])
    assert np.array_equal(nat.kth_distance(coords, 4), oracle.kth_distance(pts, 4, backend="c"))
    rad = np.abs(rng.normal(size=3_000)) * 0.5
    rad[:5] = [np.inf, -1.0, 0.0, 1e-300, 10.0]
    assert np.array_equal(nat.ball_count(coords, rad), oracle.ball_count(pts, rad, backend="c"))
    # within-class variants against per-class oracle calls
    want_k = np.empty(3_000); want_c = np.empty(3_000, dtype=np.int64)
    for g in range(7):
        sel = cls == g
        want_k[sel] = oracle.kth_distance(pts[sel], 4, backend="c")
        want_c[sel] = oracle.ball_count(pts[sel], rad[sel], backend="c")
    assert np.array_equal(nat.kth_distance(coords, 4, cls=cls, ncls=7), want_k)
    assert np.array_equal(nat.ball_count(coords, rad, cls=cls, ncls=7, within_class=True), want_c)
    # 1-D search path and 1-D all-pairs path agree with the oracle
    c1 = nat.pack_coords([pts[:, 0]])
    for flags in (0, nat.FLAG_BRUTE_COUNT, nat.FLAG_NO_PRUNE):
        assert np.array_equal(nat.ball_count(c1, rad, flags=flags), oracle.ball_count(pts[:, :1], rad, backend="c"))


def test_edge_cases():
    # n just above k; exact duplicates everywhere; a single class; dimension limit; non-finite input
    x = np.array([0.0, 1.0, 3.0, 7.0]); y = np.array([1.0, 0.0, 2.0, 5.0])
    import oracle
    want = oracle.ksg_mi(x, y, 3)
    value, parts = nat.ksg_mi(nat.pack_coords([x, y]), 3, details=True)
    assert np.array_equal(parts["eps"], want["eps"]) and np.array_equal(parts["nx"], want["nx"]) and close(value, want["value"])
    ones = np.ones(100)
    assert nat.ksg_mi(nat.pack_coords([ones, ones]), 3) == -np.inf
    assert nat.entropy(nat.pack_coords([ones]), 3) == -np.inf
    assert np.isnan(nat.cmi(nat.pack_coords([ones, ones, ones]), 3))
    xs = np.random.default_rng(0).normal(size=200)
    cls = np.zeros(200, dtype=np.int32)
    v1, p1 = nat.ross_mi(nat.pack_coords([xs]), cls, 1, 3, details=True)
    w1 = oracle.semidiscrete_mi(xs, cls, 3)
    assert np.array_equal(p1["eps"], w1["eps"]) and np.array_equal(p1["n_full"], w1["n_full"]) and close(v1, w1["value"])
    with pytest.raises(ValueError, match="data must be finite"):
        nat.ksg_mi(nat.pack_coords([np.array([0.0, np.inf, 1.0, 2.0, 3.0]), np.arange(5.0)]), 2)
    with pytest.raises(NotImplementedError):
        nat.entropy(np.zeros((33, 50)), 3)
    big = np.random.default_rng(1).normal(size=(600, 12))
    assert np.array_equal(nat.entropy(nat.pack_coords([big]), 3, details=True)[1]["dist"], oracle.kth_distance(big, 3, backend="c"))


def test_full_size_properties_config2():
    """BASELINE.json configs[1] at full size (N = 10^6, k = 3), where the oracle is too slow for
    the whole set: (a) pruned and brute-force kernels agree bit for bit on eps and counts,
    (b) a random sample of rows agrees with the oracle's all-pairs restatement run on just those
    query rows, (c) row shards sum to the whole, (d) run-to-run bitwise determinism,
    (e) count symmetry: sum_i n_x(i) equals the number of ordered pairs within radius either way."""
    import oracle
    n = 1_000_000
    rng = np.random.default_rng(0)
    d = rng.multivariate_normal([0, 0], [[1, 0.6], [0.6, 1]], size=n)
    coords = nat.pack_coords([d[:, 0], d[:, 1]])
    v_fast, fast = nat.ksg_mi(coords, 3, details=True)
    v_brute, brute = nat.ksg_mi(coords, 3, flags=nat.FLAG_NO_PRUNE, details=True)
    for key in ("eps", "nx", "ny"):
        assert np.array_equal(fast[key], brute[key]), key
    assert close(v_fast, v_brute)
    assert abs(v_fast - (-0.5 * np.log(1 - 0.36))) < 0.01                     # analytic MI of the Gaussian
    rows = rng.choice(n, 2_000, replace=False)
    pts = np.column_stack((d[:, 0], d[:, 1]))
    eps_want = oracle.kth_distance(pts, 3, query=pts[rows], backend="c")
    assert np.array_equal(fast["eps"][rows], eps_want)
    assert np.array_equal(fast["nx"][rows], oracle.ball_count(pts[:, :1], eps_want - 1e-12, query=pts[rows, :1], backend="c"))
    assert np.array_equal(fast["ny"][rows], oracle.ball_count(pts[:, 1:], eps_want - 1e-12, query=pts[rows, 1:], backend="c"))
    parts = [nat.ksg_mi_rows(coords.ctypes.data, n, 3, lo, hi) for lo, hi in ((0, 300_000), (300_000, 650_001), (650_001, n))]
    total = np.sum(parts, axis=0)
    assert total[nat.P_ROWS] == n
    assert close(nat.ksg_mi_finish(total, n, 3), v_fast)
    assert nat.ksg_mi(coords, 3) == v_fast                                      # bitwise repeatable


def test_full_size_config2_every_row_vs_reference_calls():
    """BASELINE.json configs[1] (N = 10^6, k = 3): EVERY eps / n_x / n_y against the reference's own SciPy
    calls (oracle "scipy" backend = cKDTree.query + query_ball_point on all rows,
    /root/reference/ennemi/_entropy_estimators.py:100-110), bit for bit; value within 1e-10."""
    import oracle
    n = 1_000_000
    rng = np.random.default_rng(0)
    d = rng.multivariate_normal([0, 0], [[1, 0.6], [0.6, 1]], size=n)
    x, y = np.ascontiguousarray(d[:, 0]), np.ascontiguousarray(d[:, 1])
    want = oracle.ksg_mi(x, y, 3, backend="scipy")
    value, parts = nat.ksg_mi(nat.pack_coords([x, y]), 3, details=True)
    assert int(np.count_nonzero(parts["eps"] != want["eps"])) == 0
    assert int(np.count_nonzero(parts["nx"] != want["nx"])) == 0
    assert int(np.count_nonzero(parts["ny"] != want["ny"])) == 0
    assert close(value, want["value"])


def test_full_size_other_configs():
    """configs[2] one lag (N=200,000, 3-D condition), configs[4] Ross (N=500,000, 16 classes, k=5)
    and 4-D entropy (N=500,000, k=5) against the SciPy backend = the reference's own calls."""
    import oracle
    rng = np.random.default_rng(0)
    n = 200_000
    z = rng.normal(size=(n, 3)); x = rng.normal(size=n) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=n)
    want = oracle.conditional_mi(x, y, z, 3, backend="scipy")
    value, parts = nat.cmi(nat.pack_coords([x, y, z]), 3, details=True)
    for key in ("eps", "nxz", "nyz", "nz"):
        assert np.array_equal(parts[key], want[key]), key
    assert close(value, want["value"])
    n = 500_000
    yd = rng.integers(0, 16, n); xc = rng.normal(size=n) + 0.25 * yd
    want = oracle.semidiscrete_mi(xc, yd, 5, backend="scipy")
    cls, ncls = classes(yd)
    value, parts = nat.ross_mi(nat.pack_coords([xc]), cls, ncls, 5, details=True)
    assert np.array_equal(parts["eps"], want["eps"]) and np.array_equal(parts["n_full"], want["n_full"])
    assert close(value, want["value"])
    cov = np.array([[1.0, 0.5, 0.6, -0.2], [0.5, 1.0, 0.7, -0.5], [0.6, 0.7, 2.0, -0.1], [-0.2, -0.5, -0.1, 0.5]])
    x4 = rng.multivariate_normal([0, 0, 0, 0], cov, size=n)
    want = oracle.knn_entropy(x4, 5, backend="scipy")
    value, parts = nat.entropy(nat.pack_coords([x4]), 5, details=True)
    assert np.array_equal(parts["dist"], want["dist"]) and close(value, want["value"])



def same(got, want, keys, tag, gpu_again=None, oracle_again=None):
    """Bit-equality of the per-row outputs; on a mismatch the message says where, and whether a repeat of the GPU call or
    of the oracle call gives the same answer again (an unstable side shows up as a non-zero count)."""
    for key in keys:
        bad = np.flatnonzero(np.asarray(got[key]) != np.asarray(want[key]))
        if bad.size:
            rep_gpu = [int(np.sum(gpu_again()[key] != want[key])) for _ in range(3)] if gpu_again else None
            rep_orc = int(np.sum(oracle_again()[key] != want[key])) if oracle_again else None
            raise AssertionError((tag, key, int(bad.size), bad[:8].tolist(), np.asarray(got[key])[bad[:8]].tolist(),
                                  np.asarray(want[key])[bad[:8]].tolist(), "gpu repeats vs want", rep_gpu, "oracle repeat vs want", rep_orc))


@pytest.mark.parametrize("n", [2, 5, 16, 17, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2049])
def test_tile_and_chunk_boundary_sizes(n):
    """Sizes around the 16-slot padding, the 256/512-row tiles and the 512-slot chunks, for the 2-D
    (two-level), 1-D (one-level) and 5-D paths, pruned and brute force."""
    import oracle
    rng = np.random.default_rng(n)
    k = min(3, n - 1)
    x = rng.normal(size=n); y = 0.5 * x + rng.normal(size=n)
    want = oracle.ksg_mi(x, y, k, backend="c")
    for flags in (0, nat.FLAG_NO_PRUNE):
        value, parts = nat.ksg_mi(nat.pack_coords([x, y]), k, flags=flags, details=True)
        same(parts, want, ("eps", "nx", "ny"), ("ksg", n, flags), lambda: nat.ksg_mi(nat.pack_coords([x, y]), k, flags=flags, details=True)[1],
             lambda: oracle.ksg_mi(x, y, k, backend="c"))
        assert close(value, want["value"])
    assert np.array_equal(nat.entropy(nat.pack_coords([x]), k, details=True)[1]["dist"], oracle.kth_distance(x, k, backend="c"))
    if n > 4:
        z = rng.normal(size=(n, 3))
        wc = oracle.conditional_mi(x, y, z, k, backend="c")
        vc, pc = nat.cmi(nat.pack_coords([x, y, z]), k, details=True)
        same(pc, wc, ("eps", "nxz", "nyz", "nz"), ("cmi", n), lambda: nat.cmi(nat.pack_coords([x, y, z]), k, details=True)[1],
             lambda: oracle.conditional_mi(x, y, z, k, backend="c"))
        assert close(vc, wc["value"])


@pytest.mark.parametrize("kind", ["cauchy", "outlier", "x_ties", "clusters", "lattice", "sorted_input"])
def test_hard_distributions(kind):
    """Data that stress the pruning: heavy tails and a far outlier (wide windows, deferred stragglers), ties
    in the sort coordinate, tight clusters, a lattice with many exact ties in both coordinates, presorted rows."""
    import oracle
    rng = np.random.default_rng(len(kind))
    n = 6_000
    if kind == "cauchy":
        x = rng.standard_cauchy(n); y = x + rng.standard_cauchy(n)
    elif kind == "outlier":
        x = rng.normal(size=n); y = rng.normal(size=n); x[7] = 1e6; y[11] = -1e7
    elif kind == "x_ties":
        x = rng.integers(0, 20, n).astype(float); y = rng.normal(size=n) + 0.1 * x
    elif kind == "clusters":
        c = rng.integers(0, 5, n); x = c * 100.0 + rng.normal(size=n) * 1e-3; y = c * -50.0 + rng.normal(size=n) * 1e-3
    elif kind == "lattice":
        x = rng.integers(0, 40, n).astype(float); y = rng.integers(0, 40, n).astype(float)
    else:
        x = np.sort(rng.normal(size=n)); y = np.sort(rng.normal(size=n))[::-1].copy()
    for k in (1, 3, 20):
        want = oracle.ksg_mi(x, y, k, backend="c")
        value, parts = nat.ksg_mi(nat.pack_coords([x, y]), k, details=True)
        for key in ("eps", "nx", "ny"):
            assert np.array_equal(parts[key], want[key]), (kind, k, key, int(np.sum(parts[key] != want[key])))
        assert close(value, want["value"]), (kind, k, value, want["value"])
    z = np.column_stack((y, x * 0.5 + rng.normal(size=n)))
    wc = oracle.conditional_mi(x, y, z, 3, backend="c")
    vc, pc = nat.cmi(nat.pack_coords([x, y, z]), 3, details=True)
    same(pc, wc, ("eps", "nxz", "nyz", "nz"), ("cmi", kind), lambda: nat.cmi(nat.pack_coords([x, y, z]), 3, details=True)[1],
         lambda: oracle.conditional_mi(x, y, z, 3, backend="c"))
    assert close(vc, wc["value"]), kind


def test_ragged_classes():
    """Ross estimators with classes of size 1, k, k+1 and one big class; class ids in arbitrary row order."""
    import oracle
    rng = np.random.default_rng(3)
    sizes = [1, 3, 4, 5, 16, 17, 600, 2000]
    y = np.repeat(np.arange(len(sizes)), sizes); rng.shuffle(y)
    n = len(y)
    x = rng.normal(size=n) + 0.3 * y; z = rng.normal(size=(n, 2))
    cls, ncls = classes(y)
    for k in (3, 4):
        w = oracle.semidiscrete_mi(x, y, k, backend="c")
        v, p = nat.ross_mi(nat.pack_coords([x]), cls, ncls, k, details=True)
        assert np.array_equal(p["eps"], w["eps"]) and np.array_equal(p["n_full"], w["n_full"]) and close(v, w["value"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            wc = oracle.conditional_semidiscrete_mi(x, y, z, k, backend="c")
        vc, pc = nat.ross_cmi(nat.pack_coords([x, z]), cls, ncls, k, details=True)
        assert all(np.array_equal(pc[key], wc[key]) for key in ("eps", "nxz", "nyz", "nz")) and close(vc, wc["value"])


@pytest.mark.parametrize("d,n,k", [(13, 3_000, 3), (20, 2_500, 5), (32, 1_111, 12)])
def test_wide_spaces_generic_kernels(d, n, k):
    """Joint spaces wider than the 12 dimensions the specialised kernels cover go through the generic
    run-time-dimension kernels: entropy in d dimensions, CMI with a (d-2)-dimensional condition and
    conditional Ross with a (d-1)-dimensional condition, all against the C oracle."""
    import oracle
    rng = np.random.default_rng(100 * d + k)
    mix = rng.normal(size=(d, d)) / np.sqrt(d) + np.eye(d)
    x = rng.normal(size=(n, d)) @ mix
    want = oracle.knn_entropy(x, k, backend="c")
    value, parts = nat.entropy(nat.pack_coords([x]), k, details=True)
    assert np.array_equal(parts["dist"], want["dist"]) and close(value, want["value"])

    z = x[:, 2:]
    a = rng.normal(size=n) + z[:, 0]
    b = rng.normal(size=n) + 0.7 * a - z[:, 1]
    want = oracle.conditional_mi(a, b, z, k, backend="c")
    for flags in (0, nat.FLAG_NO_PRUNE):
        value, parts = run_gpu("cmi", {"x": a, "y": b, "z": z, "k": k}, flags)
        for key, arr in parts.items():
            assert np.array_equal(arr, want[key]), (d, key, flags)
        assert close(value, want["value"])

    yd = rng.integers(0, 3, n)
    z1 = x[:, 1:]
    a = rng.normal(size=n) + yd + z1[:, 0]
    want = oracle.conditional_semidiscrete_mi(a, yd, z1, k, backend="c")
    value, parts = run_gpu("cross", {"x": a, "y": yd, "z": z1, "k": k}, 0)
    for key, arr in parts.items():
        assert np.array_equal(arr, want[key]), (d, key)
    assert close(value, want["value"])


@pytest.mark.parametrize("dist", ["gauss", "student_t2", "ties_plus_noise"])
def test_search_variants_agree_bit_for_bit(dist, monkeypatch):
    """Every way the library can lay out and walk the two-level search gives the same eps and counts, bit for bit,
    as the brute-force kernel: per-lane window walk (default) vs the all-lanes window scan (EB2_LANE_SCAN=0) vs
    a hybrid (walk only narrow warp windows), layout by partition (default at this size) vs per-chunk sort
    (EB2_CELL_SORT), quarter tiles at the edge chunks on / off, straggler deferral off / default / aggressive
    (one-warp and whole-CTA leftover paths).  Sizes sit above the partition threshold (300,000 rows)."""
    rng = np.random.default_rng(17)
    n = 320_001
    if dist == "gauss":
        d = rng.multivariate_normal([0, 0], [[1, 0.8], [0.8, 1]], size=n)
    elif dist == "student_t2":
        d = rng.standard_t(2, size=(n, 2))
    else:       # what ennemi's preprocessing makes of integer-valued data: clusters 1e-10 wide
        d = rng.integers(0, 50, size=(n, 2)).astype(float) + rng.normal(0, 1e-10, size=(n, 2))
    coords = nat.pack_coords([d[:, 0], d[:, 1]])
    v_ref, ref = nat.ksg_mi(coords, 3, flags=nat.FLAG_NO_PRUNE | nat.FLAG_BRUTE_COUNT, details=True)
    variants = [{}, {"EB2_LANE_SCAN": "0"}, {"EB2_LANE_SCAN": "96"}, {"EB2_CELL_SORT": "1"}, {"EB2_EDGE_CHUNKS": "0"},
                {"EB2_EDGE_CHUNKS": "5"}, {"EB2_DEFER": "0"}, {"EB2_DEFER": "64"}, {"EB2_DEFER": "64", "EB2_LANE_SCAN": "0"},
                {"EB2_PARTITION_MIN": "1000000000"}, {"EB2_NO_CELLS": "1"}]
    for env in variants:
        with monkeypatch.context() as m:
            for key, val in env.items():
                m.setenv(key, val)
            value, got = nat.ksg_mi(coords, 3, details=True)
        for key in ("eps", "nx", "ny"):
            assert np.array_equal(got[key], ref[key]), (dist, env, key)
        assert close(value, v_ref) or (np.isinf(value) and value == v_ref), (dist, env)


def test_walk_variants_agree_in_wider_spaces(monkeypatch):
    """3-D and 5-D spaces (k = 3 and k = 6: both register list lengths): per-lane walk vs all-lanes scan vs brute force."""
    rng = np.random.default_rng(23)
    for dims, n, k in ((3, 120_000, 3), (5, 60_000, 6), (4, 50_000, 7)):
        x = rng.normal(size=(n, dims)) @ rng.normal(size=(dims, dims))
        coords = nat.pack_coords([x])
        _, ref = nat.entropy(coords, k, flags=nat.FLAG_NO_PRUNE, details=True)
        for env in ({}, {"EB2_LANE_SCAN": "0"}, {"EB2_LANE_DIM": "2"}, {"EB2_DEFER": "0"}, {"EB2_DEFER": "100"}):
            with monkeypatch.context() as m:
                for key, val in env.items():
                    m.setenv(key, val)
                _, got = nat.entropy(coords, k, details=True)
            assert np.array_equal(got["dist"], ref["dist"]), (dims, env)


def test_count_work_splitting_is_exact(monkeypatch):
    """The fused Frenzel-Pompe / Ross counts with the chunk range of every tile dealt to 1, 2, 3 or 8 CTAs that add their
    counts (EB2_COUNT_SPLIT; the default picks 1 to 4 from the number of tiles): the same three counts, bit for bit."""
    rng = np.random.default_rng(31)
    n = 70_000
    z = rng.standard_t(3, size=(n, 3)); x = rng.normal(size=n) + z[:, 0]; y = 0.5 * x + z[:, 1] + rng.normal(size=n)
    co = nat.pack_coords([x, y, z])
    brute = nat.FLAG_NO_PRUNE | nat.FLAG_BRUTE_COUNT
    v_ref, ref = nat.cmi(co, 3, flags=brute, details=True)
    yd = rng.integers(0, 7, n)
    cls, ncls = classes(yd)
    cr = nat.pack_coords([x, z[:, :2]])
    vr_ref, rref = nat.ross_cmi(cr, cls, ncls, 3, flags=brute, details=True)
    for split in ("1", "2", "3", "8", None):
        with monkeypatch.context() as m:
            if split is not None:
                m.setenv("EB2_COUNT_SPLIT", split)
            v, got = nat.cmi(co, 3, details=True)
            vr, rgot = nat.ross_cmi(cr, cls, ncls, 3, details=True)
        for key in ("eps", "nxz", "nyz", "nz"):
            assert np.array_equal(got[key], ref[key]), (split, key)
            assert np.array_equal(rgot[key], rref[key]), (split, "ross", key)
        assert close(v, v_ref) and close(vr, vr_ref), split
