#!/usr/bin/env python
"""Numerical study for VERDICT item 8: could the variance contraction V = L^-1 K*^T run on the int8 tensor pipe
(tcgen05 kind::i8, TMEM accumulators) through an Ozaki-style error-free split and still meet sigma^2 <= 1e-10?

Emulation (exact): operand rows / columns are scaled by a power of two and cut into S int8 digits, every slice-pair
product is an exact integer GEMM (emulated with fp64 BLAS on integer-valued matrices: K * 128^2 < 2^53), and the pair
products are combined in fp64 with their scales, least significant level first, keeping the pairs with s + t < S (the
usual triangle).  Two digit schemes: "base128" = 7-bit digits by truncation (|q| <= 127; the first version of the
kernel, S = 8, 36 pairs) and "base256" = BALANCED base-256 digits (-128 .. 127, 8 bits per int8; what gpk_ozaki.cuh
ships: S = 7, 28 pairs).  Reported: the posterior-variance error against an 80-bit reference, in the units of the
parity tests (|dvar| / max(var, 1e-6 k**)), for the plain fp64 path and for both schemes over a range of S, plus the
int8 products per fp64 product that S implies.

CPU only (numpy); run: python tools/ozaki_study.py [N] [M]
"""
import sys
import os
import numpy as np
import scipy.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import robo_oracle as O   # noqa: E402  (study tool, not product code)


def split(A, S, axis, global_max=None):
    """A ~ sum_s Q_s * 2^(e - 7 (s+1)) with integer |Q_s| <= 127; e = per-row (axis=1) or per-column (axis=0) exponent,
    or ONE exponent from global_max (what the kernel does for K*: 0 < k <= amp)."""
    mx = np.max(np.abs(A), axis=axis, keepdims=True) if global_max is None else np.full((1, 1), float(global_max))
    e = np.where(mx > 0, np.ceil(np.log2(np.where(mx > 0, mx, 1.0))) + 1, 0.0)        # |A| / 2^e < 1/2
    r = A / np.exp2(e)
    Q = []
    for _ in range(S):
        r = r * 128.0
        q = np.trunc(r)
        r = r - q
        Q.append(q)
    return Q, e


def split256(A, S, axis, global_max=None):
    """A ~ sum_s Q_s * 2^(e - 8 (s+1)) with balanced digits -128 <= Q_s <= 127 (oz_exponent / oz_digit of gpk_ozaki.cuh):
    remainders stay in [-128/255, 127/255), so S digits carry 8 S bits."""
    mx = np.max(np.abs(A), axis=axis, keepdims=True) if global_max is None else np.full((1, 1), float(global_max))
    m, ex = np.frexp(mx)
    e = np.where(mx > 0, ex + 1 + (m * 128.0 >= 127.49), 0).astype(float)
    r = A / np.exp2(e)
    Q = []
    if S <= 7:
        # the kernel's way (oz_digits): X = rint(v 256^S), the bytes of X + 0x80..80 are the digits + 128
        X = np.rint(r * 256.0 ** S).astype(np.int64)
        Y = X + sum(128 << (8 * j) for j in range(S))
        for s_ in range(S):
            Q.append((((Y >> (8 * (S - 1 - s_))) & 0xFF) - 128).astype(np.float64))
        return Q, e
    for _ in range(S):                                   # more than 56 bits: same digits by floating-point peeling
        r = r * 256.0
        q = np.clip(np.floor(r + 128.0 / 255.0), -128, 127)
        r = r - q
        Q.append(q)
    return Q, e


def ozaki_matmul(P, Kt, S, amp=None, base=128):
    """P (n x k) @ Kt (k x m) from S x S slices, pairs with s + t <= S - 1 (0-based)."""
    bits = 7.0 if base == 128 else 8.0
    sp = split if base == 128 else split256
    QP, eP = sp(P, S, axis=1)
    QK, eK = sp(Kt, S, axis=0, global_max=amp)
    out = np.zeros((P.shape[0], Kt.shape[1]))
    pairs = 0
    for lvl in range(S - 1, -1, -1):                  # least significant level first, then upwards
        acc = np.zeros_like(out)
        for s in range(lvl + 1):
            t = lvl - s
            acc += QP[s] @ QK[t]                      # exact: integers < 2^53
            pairs += 1
        assert np.abs(acc).max() < 2.0 ** 31             # what the int32 TMEM accumulator of a level must hold
        out += acc * np.exp2(-bits * (lvl + 2))
    return out * np.exp2(eP) * np.exp2(eK), pairs


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    D = 16
    X, y, Xs, theta, noise = O.synthetic_problem(N, D, M)
    for label, th, nz in (("C2 hyper-parameters (metric D/4, noise 1e-3)", theta, noise),
                          ("ill-conditioned: metric x4, noise 1e-6 (cond ~1e9+)", theta + np.r_[0, np.full(D, np.log(4.0))], 1e-6)):
        kernel = O.make_kernel("matern52", D, th)
        K = kernel.get_value(X)
        K[np.diag_indices_from(K)] += nz + 1.25e-12
        L = spla.cholesky(K, lower=True)
        P = spla.solve_triangular(L, np.eye(N), lower=True)
        Ks = kernel.get_value(Xs, X)
        amp = float(np.exp(th[0]))
        # 80-bit reference of sum_i V_i^2 with V = P Ks^T (P itself is the fp64 factor's inverse: both paths share it)
        Pl, Kl = P.astype(np.longdouble), Ks.T.astype(np.longdouble)
        Vl = np.empty((N, M), dtype=np.longdouble)
        for j in range(M):
            Vl[:, j] = (Pl * Kl[:, j][None, :]).sum(axis=1)
        var_ref = (amp - (Vl * Vl).sum(axis=0)).astype(np.float64)
        den = np.maximum(np.abs(var_ref), 1e-6 * amp)
        V64 = P @ Ks.T
        err64 = np.max(np.abs((amp - np.einsum("ij,ij->j", V64, V64)) - var_ref) / den)
        print("== %s: N=%d, M=%d, cond(K) ~ %.1e, max|P| = %.1e, var in [%.2e, %.2e]"
              % (label, N, M, np.linalg.cond(K), np.abs(P).max(), var_ref.min(), var_ref.max()))
        print("   fp64 (DMMA-equivalent) path: scaled variance error %.2e" % err64)
        for base, rng in ((128, range(6, 11)), (256, range(5, 10))):
            for S in rng:
                V, pairs = ozaki_matmul(P, Ks.T, S, amp, base)
                err = np.max(np.abs((amp - np.einsum("ij,ij->j", V, V)) - var_ref) / den)
                print("   base%d S = %2d slices/operand: %3d int8 GEMMs per fp64 GEMM, scaled variance error %.2e  %s"
                      % (base, S, pairs, err, "<= 1e-10" if err <= 1e-10 else ""))


if __name__ == "__main__":
    main()
