"""CPU oracle for the ennemi k-NN hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``ennemi_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, and only as the checker / the timed CPU baseline.

Parity status: **pinned against the reference run in the build container** — the fixtures
under ``tests/golden/`` were produced by ``oracle/make_golden.py`` importing the unmodified
reference from ``/root/reference`` (ennemi 1.5.0, SciPy 1.18.1), and
``tests/test_oracle_golden.py`` checks every oracle backend against them bit for bit.
The reference's own test-suite holds no bit-level vectors for this path (all its
assertions are statistical tolerances); the 8-digit outputs printed in
its docs are included in the fixtures as known-answer tests.
"""
from .estimators import (  # noqa: F401
    psi, ksg_mi, conditional_mi, semidiscrete_mi, conditional_semidiscrete_mi, knn_entropy,
    kth_distance, ball_count, BACKENDS,
)
