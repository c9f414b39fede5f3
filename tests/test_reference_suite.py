"""Runs the REFERENCE'S OWN test-suite (unit, pandas, integration: 165 tests) against
``ennemi_b200``'s host logic, with the estimator seam routed to the CPU oracle.

Only possible where ``/root/reference`` is mounted (the build container); skipped elsewhere.  The
reference's files are executed where they lie — nothing is copied into this repository."""
import os
import subprocess
import sys
import textwrap

import pytest

REF_TESTS = "/root/reference/tests"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = textwrap.dedent("""
    import sys
    sys.dont_write_bytecode = True
    sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
    import ennemi_b200
    from ennemi_b200 import _native, _devices, _estimators
    from conftest import OracleBackend
    fake = OracleBackend()
    for name in OracleBackend.PATCHED:
        setattr(_native, name, getattr(fake, name))
    _native.device_count = lambda: 1
    _devices.visible = lambda: [0]
    sys.modules["ennemi"] = ennemi_b200                       # `from ennemi import estimate_mi, ...`
    sys.modules["ennemi._entropy_estimators"] = _estimators   # `from ennemi._entropy_estimators import _psi, ...`
    import pytest
    sys.exit(pytest.main(["-q", "-x", "-p", "no:cacheprovider", "--rootdir=/tmp"] + sys.argv[1:]))
""")


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference tree not mounted")
@pytest.mark.parametrize("suite,min_rows", [("unit", None), ("pandas", None), ("integration", None),
                                            ("unit", "0"), ("pandas", "0")])
def test_reference_suite_passes_on_our_host_logic(suite, min_rows):
    """min_rows="0" forces the device-column path (cached columns + device-side rescaling, emulated
    by the oracle backend) for every eligible call, so the reference's tests cover it as well."""
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    if min_rows is not None:
        env["ENNEMI_B200_COLUMNS_MIN_ROWS"] = min_rows
    proc = subprocess.run([sys.executable, "-c", DRIVER.format(root=ROOT), os.path.join(REF_TESTS, suite)],
                          capture_output=True, text=True, env=env, cwd="/tmp", timeout=900)
    tail = (proc.stdout + proc.stderr)[-2000:]
    assert proc.returncode == 0, tail
    assert " passed" in proc.stdout and "failed" not in proc.stdout, tail
