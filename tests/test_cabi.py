"""The C-ABI library: builds, loads, exports every symbol the header declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from ennemi_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ennemi_b200.h")).read()
    return sorted(set(re.findall(r"EB2_API\s+[\w\s\*]+?\b(eb2_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    assert set(syms) == set(_native.EXPORTS)
    for must in ("eb2_ksg_mi", "eb2_cmi", "eb2_ross_mi", "eb2_ross_cmi", "eb2_entropy", "eb2_psi"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_native.LIB_PATH), "build the library first (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), name
    lib.eb2_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.eb2_version()


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device():
    if _native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    coords = np.zeros((2, 16))
    with pytest.raises(RuntimeError, match="ennemi_b200"):
        _native.ksg_mi(coords, 3)
    import ennemi_b200
    with pytest.raises(RuntimeError):
        ennemi_b200.estimate_mi(np.arange(20.0), np.arange(20.0) ** 2, preprocess=False)


def test_argument_errors_do_not_need_a_device():
    lib = _native.load()
    val = ctypes.c_double()
    assert lib.eb2_ksg_mi(0, None, 10, 3, 0, ctypes.byref(val), None, None, None) == _native.ERR_ARG
    assert b"NULL" in lib.eb2_last_error()
    z = np.zeros((40, 8))        # 40 dimensions > EB2_MAX_DIM (32)
    assert lib.eb2_entropy(0, z.ctypes.data, 8, 40, 3, 0, ctypes.byref(val), None) == _native.ERR_UNSUPPORTED
    part = np.zeros(_native.P_LEN)
    part[_native.P_SUM] = 10.0
    v = _native.ksg_mi_finish(part, 5, 1)     # psi(5) + psi(1) - 10/5, host-only arithmetic
    import oracle
    assert abs(v - (oracle.psi(np.array([5]))[0] + oracle.psi(np.array([1]))[0] - 2.0)) < 1e-14
    part[_native.P_ZERO_A] = 1
    assert _native.ksg_mi_finish(part, 5, 1) == -np.inf
    part[_native.P_ZERO_C] = 1
    assert np.isnan(_native.cmi_finish(part, 5, 1))
    # the bivariate pipeline's exact sum: 128-bit fixed point (2^-48 units) in four 32-bit limbs, the last one signed;
    # limbs of several shards add as doubles without rounding, whatever the number of shards
    def limbs(total):
        t = total & ((1 << 128) - 1)
        out = [(t >> (32 * i)) & 0xffffffff for i in range(4)]
        out[3] -= (1 << 32) if out[3] >= (1 << 31) else 0
        return out
    want = -1234.567890123
    pieces = [int(round(x * 2 ** 48)) for x in (want * 0.25, want * 0.5, want * 0.125, want * 0.125)]
    fixed = np.zeros(_native.P_LEN)
    for q in pieces:
        fixed[8:12] += limbs(q)
        fixed[12] += 1
    fixed[_native.P_SUM] = 777.0              # ignored when the limbs are present
    v = _native.ksg_mi_finish(fixed, 5, 1)
    exact = sum(pieces) / 2 ** 48
    assert abs(v - (oracle.psi(np.array([5]))[0] + oracle.psi(np.array([1]))[0] - exact / 5)) < 1e-12
