"""CPU checks of the error-free split behind the int8 scoring path (gpk_ozaki.cuh: oz_exponent / oz_digits), through its
numpy restatement in tools/ozaki_study.py: digits fit an int8, the 7-digit sum is within half a unit of the 56th bit of
the scaled value, the exponent rule keeps every first digit in range (also just below a power of two), and the 28
digit-pair products of two split operands reproduce an fp64 GEMM to ~1e-16 relative."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ozaki_study as Z                                       # noqa: E402


def _reconstruct(Q, e, bits=8.0):
    return sum(q * np.exp2(-bits * (s + 1)) for s, q in enumerate(Q)) * np.exp2(e)


def test_digits_are_int8_and_exact_to_56_bits():
    rng = np.random.RandomState(0)
    A = rng.randn(64, 300) * np.exp(rng.uniform(-20, 5, (64, 1)))
    A[3, :] = 0.0                                             # an all-zero row: exponent 0, digits 0
    A[5, 7] = np.abs(A[5]).max() * 3.0                        # one dominant entry
    Q, e = Z.split256(A, 7, axis=1)
    for q in Q:
        assert q.min() >= -128 and q.max() <= 127 and np.all(q == np.rint(q))
    err = np.abs(_reconstruct(Q, e) - A)
    assert np.all(err <= 0.5 * np.exp2(e - 56.0) * (1 + 1e-12))
    assert not np.any(_reconstruct(Q, e)[3])


def test_exponent_rule_at_the_edge_of_the_digit_interval():
    # largest mantissas just below 1: rint(v 2^56) + 0x80..80 must stay below 2^56, i.e. |v| < 127.49 / 256 after scaling
    vals = np.array([[1.0 - 2.0 ** -53, 0.3], [0.99609375, -0.2], [0.9960937, 0.1], [0.5, -0.5], [31.874931782080985, 1.0],
                     [-(1.0 - 2.0 ** -30), 0.25], [2.0 ** -1040, 2.0 ** -1045]])
    Q, e = Z.split256(vals, 7, axis=1)
    scaled = np.abs(vals) / np.exp2(e)
    assert np.all(scaled < 127.49 / 256.0)
    assert np.all(scaled.max(axis=1)[vals.max(axis=1) != 0] >= 0.12)          # no more than two bits of headroom are given away
    for q in Q:
        assert q.min() >= -128 and q.max() <= 127
    np.testing.assert_allclose(_reconstruct(Q, e), vals, rtol=0, atol=float(np.max(np.exp2(e - 56.0))))


def test_28_digit_pair_products_reproduce_the_fp64_product():
    rng = np.random.RandomState(1)
    n, k, m = 96, 512, 40
    P = np.tril(rng.randn(n, k) * np.exp(-0.01 * np.abs(np.arange(n)[:, None] - np.arange(k)[None, :])))
    Kt = rng.rand(k, m)
    V, pairs = Z.ozaki_matmul(P, Kt, 7, amp=1.0, base=256)
    assert pairs == 28
    ref = (P.astype(np.longdouble) @ Kt.astype(np.longdouble)).astype(np.float64)
    scale = np.abs(P).max(axis=1, keepdims=True) * k
    assert np.max(np.abs(V - ref) / scale) < 2e-16
    V6, pairs6 = Z.ozaki_matmul(P, Kt, 6, amp=1.0, base=256)
    assert pairs6 == 21 and np.max(np.abs(V6 - ref) / scale) > np.max(np.abs(V - ref) / scale)
