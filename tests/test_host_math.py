"""Host-side floating-point pieces against the numpy / scipy calls the reference makes."""
import math

import numpy as np
import pytest
from scipy.stats import binom

from phaser_b200 import pipeline


@pytest.mark.parametrize("seed", range(12))
def test_percentile_from_histogram_equals_numpy(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 5000))
    vals = rng.integers(-300, 400, size=n) if seed % 3 else rng.integers(120, 151, size=n)
    hist = np.bincount(vals + 32768, minlength=65536).astype(np.int64)
    for q in (0.05, 0.01, 0.5, 0.333, 0.999, 1.0):
        assert pipeline.percentile_from_histogram(hist, q) == float(np.percentile(vals.tolist(), q * 100))


@pytest.mark.parametrize("noise", [0.0, 1e-4, 0.0018488389291524921, 0.01, 0.04])
@pytest.mark.parametrize("thr", [0.01, 0.05])
def test_critical_values_equal_direct_test(noise, thr):
    p = 1 - ((6 * noise) + (10 * math.pow(noise, 2)))
    ks = pipeline.critical_values(300, noise, thr)
    for n in list(range(1, 80)) + [150, 299, 300]:
        k = np.arange(0, n + 1)
        drop_direct = binom.cdf(k, n, p) < thr
        drop_table = k < ks[n]
        assert (drop_direct == drop_table).all(), (n, noise)
