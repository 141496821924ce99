"""Minimax fit of the erf-GELU gate used by the GEGLU epilogue (gcd_b200/csrc/common.cuh::geglu_pair).

Phi(g) - 1/2 = g * Q(g^2) on |g| <= A with a polynomial Q of degree `deg` in u = g^2, weights chosen so that the maximum ABSOLUTE
error of Phi is minimised (iteratively re-weighted least squares on a dense grid); prints the error of the float32 Horner evaluation.
    python tools/fit_gelu.py            # table for A in {3.5, 4}, deg 4..7; the kernel uses A = 4, deg = 6 (1.05e-4)
"""
from math import erf, sqrt

import numpy as np
from numpy.polynomial import chebyshev as C
from numpy.polynomial import polynomial as P


def fit(deg, A, n=20001, iters=200):
    g = np.linspace(1e-4, A, n)
    u = g * g
    target = (np.array([0.5 * (1 + erf(x / sqrt(2))) for x in g]) - 0.5) / g
    V = C.chebvander(2 * u / (A * A) - 1, deg)
    wt = np.ones_like(g)
    for _ in range(iters):
        coef, *_ = np.linalg.lstsq(V * (g * wt)[:, None], target * g * wt, rcond=None)
        err = (V @ coef - target) * g
        wt = wt * (1 + 0.5 * np.abs(err) / np.abs(err).max())
        wt /= wt.mean()
    xu, pu = np.array([-1.0, 2.0 / (A * A)]), np.zeros(1)
    for k, c in enumerate(C.cheb2poly(coef)):
        pu = P.polyadd(pu, c * P.polypow(xu, k))
    uf, q = u.astype(np.float32), np.full(n, np.float32(pu[-1]))
    for c in pu[-2::-1]:
        q = q * uf + np.float32(c)
    e32 = np.abs(np.float32(0.5) + g.astype(np.float32) * q - (target * g + 0.5)).max()
    return e32, 1 - 0.5 * (1 + erf(A / sqrt(2))), pu


if __name__ == "__main__":
    for A in (3.5, 4.0):
        for deg in (4, 5, 6, 7):
            e32, tail, pu = fit(deg, A)
            print(f"A={A} deg={deg}: max |dPhi| (fp32 Horner) {e32:.2e}, clamped tail {tail:.2e}")
            if (A, deg) == (4.0, 6):
                print("   Q coefficients, u^0..u^6:", ", ".join(f"{c:.9e}" for c in pu))
