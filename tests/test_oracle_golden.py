"""The oracle is only trustworthy if it reproduces the reference: every backend (numpy all-pairs,
C all-pairs, SciPy cKDTree) against the fixtures that ``oracle/make_golden.py`` captured from the
unmodified reference — intermediate arrays and values, bit for bit."""
import warnings

import numpy as np
import pytest

import oracle


def run_case(name, c, backend):
    k = int(c["k"])
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        if name.startswith("ksg"):
            return oracle.ksg_mi(c["x"], c["y"], k, backend=backend)
        if name.startswith("cmi"):
            return oracle.conditional_mi(c["x"], c["y"], c["z"], k, backend=backend)
        if name.startswith("ross"):
            return oracle.semidiscrete_mi(c["x"], c["y"], k, backend=backend)
        if name.startswith("cross"):
            return oracle.conditional_semidiscrete_mi(c["x"], c["y"], c["z"], k, backend=backend)
        if name.startswith("ent"):
            return oracle.knn_entropy(c["x"], k, backend=backend)
    raise AssertionError(name)


@pytest.mark.parametrize("backend", oracle.BACKENDS)
def test_oracle_matches_reference_bit_for_bit(golden_estimators, backend):
    checked = 0
    for name, c in golden_estimators.items():
        if name == "psi":
            continue
        if backend == "brute" and len(c["x"]) > 2000:
            continue
        got = run_case(name, c, backend)
        for key, val in got.items():
            if key == "value":
                assert val == float(c["value"]) or (np.isnan(val) and np.isnan(float(c["value"]))), (name, val)
            else:
                assert np.array_equal(np.asarray(val), c[key]), (name, key)
        checked += 1
    assert checked >= 20


def test_oracle_psi(golden_estimators):
    c = golden_estimators["psi"]
    assert np.array_equal(oracle.psi(c["n"]), c["value"])
    assert np.isinf(oracle.psi(np.array([3, 0, 5]))) and np.isinf(c["zero"])
    # SURVEY.md Appendix A.8 spot values
    assert oracle.psi(np.array([2]))[0] == 0.4227726765916914
    assert oracle.psi(np.array([10]))[0] == 2.251752589025792


def test_oracle_edge_semantics():
    # fewer than k+1 candidates -> inf; radius < 0 -> 0; inclusive radius; non-finite rejected
    pts = np.array([[0.0], [1.0], [3.0]])
    for be in oracle.BACKENDS:
        assert np.all(np.isinf(oracle.kth_distance(pts, 3, backend=be)))
        assert np.array_equal(oracle.kth_distance(pts, 1, backend=be), [1.0, 1.0, 2.0])
        assert np.array_equal(oracle.ball_count(pts, np.array([-1e-12, 1.0, np.inf]), backend=be), [0, 2, 3])
    with pytest.raises(ValueError, match="data must be finite"):
        oracle.ksg_mi(np.array([0.0, np.nan, 1.0, 2.0]), np.arange(4.0), 1)
