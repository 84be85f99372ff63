import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_genome():
    """Concatenated ASCII genome of the scan fixture + offsets/lengths (chrom index = CHROM-1)."""
    z = golden("scan")
    s1, s2 = z["seq_chr1"], z["seq_chr2"]
    off2 = ((len(s1) + 127) // 128) * 128
    seq = np.full(off2 + len(s2), ord("N"), dtype=np.uint8)
    seq[:len(s1)] = s1
    seq[off2:] = s2
    return seq, np.array([0, off2], dtype=np.int64), np.array([len(s1), len(s2)], dtype=np.int64)


def assert_pvals_close(got, want, tol=1e-6):
    """|delta log10 p| <= tol; both-NaN and both-(sub)zero count as equal."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), "NaN pattern differs at %s" % np.flatnonzero(nan_g != nan_w)[:10]
    tiny = 2.3e-308
    both_tiny = (got < tiny) & (want < tiny)
    m = ~nan_g & ~both_tiny
    assert np.all(got[m] > 0) and np.all(want[m] > 0), "zero vs non-zero p-value"
    d = np.abs(np.log10(got[m]) - np.log10(want[m]))
    worst = np.argmax(d) if d.size else 0
    assert d.size == 0 or d.max() <= tol, "max |dlog10 p| = %.3g at %d (got %r want %r)" % (
        d.max(), np.flatnonzero(m)[worst], got[m][worst], want[m][worst])


@pytest.fixture(scope="session")
def oracle():
    from oracle import dig_oracle
    dig_oracle.build()
    return dig_oracle
