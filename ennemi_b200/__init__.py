"""ennemi_b200 — B200-native k-nearest-neighbour mutual information (drop-in for ``ennemi``).

>>> from ennemi_b200 import estimate_mi, estimate_entropy, pairwise_mi, normalize_mi

The estimators run on NVIDIA B200 GPUs (sm_100a) through ``libennemi_b200.so``; there is no CPU
fallback.
"""
from .api import (estimate_entropy, estimate_corr, estimate_mi, normalize_mi, pairwise_corr, pairwise_mi)

__all__ = ["estimate_entropy", "estimate_corr", "estimate_mi", "normalize_mi", "pairwise_corr", "pairwise_mi"]
__version__ = "0.1.0"
