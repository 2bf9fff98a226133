"""CPU: the numpy restatement of the dropout mask generator (oracle/dropout_oracle.py) — threshold arithmetic and the
statistical quality the kernels rely on (the bit-exact comparison with the device generator is in test_gpu_dropout.py)."""
import numpy as np
import pytest

from oracle import dropout_oracle as D


def test_threshold_and_scale():
    assert D.drop_threshold(0.0) == 0
    assert D.drop_threshold(0.1) == 26 and D.drop_threshold(0.25) == 64 and D.drop_threshold(0.5) == 128
    assert D.drop_threshold(1e-6) == 1 and D.drop_threshold(0.9999) == 255
    assert D.survivor_scale(0.5) == 2.0
    assert D.survivor_scale(0.1) == pytest.approx(256 / 230)


def test_mask_statistics():
    p, seed = 0.1, 0x1234567887654321
    m = D.keep_mask(2, 2, 256, 2048, p, seed).astype(np.float64)
    keep = 1 - D.drop_threshold(p) / 256
    n = m.size
    assert abs(m.mean() - keep) < 5 * (keep * (1 - keep) / n) ** 0.5
    assert np.abs(m.mean(-1) - keep).max() < 0.05 and np.abs(m.mean(-2) - keep).max() < 0.12
    c = m - keep
    var = keep * (1 - keep)
    for dq, dk in ((0, 1), (1, 0), (1, 1), (0, 2), (2, 0), (3, 5), (0, 64)):
        a = c[:, :, : 256 - dq, : 2048 - dk]
        b = c[:, :, dq:, dk:]
        assert abs((a * b).mean()) < 0.004 * var, (dq, dk)
    # heads, batch rows and adjacent seeds are uncorrelated; the mask is a pure function of its arguments
    assert abs((c[0, 0] * c[0, 1]).mean()) < 0.004 * var and abs((c[0, 0] * c[1, 0]).mean()) < 0.004 * var
    other = D.keep_mask(1, 1, 256, 2048, p, seed + 1).astype(np.float64) - keep
    assert abs((c[0, 0] * other[0, 0]).mean()) < 0.004 * var
    assert np.array_equal(D.keep_mask(1, 2, 64, 100, p, seed), D.keep_mask(1, 2, 64, 100, p, seed))
    assert D.keep_mask(1, 1, 8, 8, 0.0, 5).all()
