import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but torch.cuda.is_available() is False")
    return torch.device("cuda:0")


def make_graph_arrays(m, n, z, seed, dup=False, sort=False):
    """Random COO (int64 row, col; float32 val) for an m x n matrix with ~z entries."""
    rng = np.random.default_rng(seed)
    row = rng.integers(0, max(m, 1), size=z)
    col = rng.integers(0, max(n, 1), size=z)
    if not dup and z:
        key = np.unique(row * n + col)
        rng.shuffle(key)
        row, col = key // n, key % n
    val = rng.uniform(-1, 1, size=row.shape[0]).astype(np.float32)
    if sort:
        o = np.lexsort((col, row))
        row, col, val = row[o], col[o], val[o]
    return row.astype(np.int64), col.astype(np.int64), val


# Known deviations of the 16-bit storage modes from the north-star logit bar (every entry within 2e-2 of the row norm 10,
# status agreement >= 99.9 %), stated in ONE place and printed by bench.py as `parity.*.known_deviation`.  With
# random-initialised weights a handful of rows have a raw logit vector 10-100x shorter than typical and F.normalize
# (reference arch.py:134-135) amplifies their rounding error by that factor.  fp32 (the default mode) has no deviation.
KNOWN_DEVIATION = {
    "bf16": dict(max_entry=1.5e-1, frac_within_bar=0.99, status_agreement=0.98, frobenius=2e-2),     # bar: 2e-2 / 1.0 / 0.999
    "fp16": dict(max_entry=5e-2, frac_within_bar=0.9999, status_agreement=0.999, frobenius=2.5e-3),  # only the worst entry deviates
}
LOGIT_BAR_16BIT = 2e-2
