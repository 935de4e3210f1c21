"""Quadratic interpolating spline as a linear map (host side, fp64).

The reference evaluates ``jax_cosmo.scipy.interpolate.InterpolatedUnivariateSpline(x, y, k=2)``
inside ``MBDPI.node2u`` / ``u2node`` (dial_mpc/core/dial_core.py:91-101).  That spline is
linear in ``y``, so the planner only ever needs the two constant matrices
``M_n2u [Hs+1, Hn+1]`` and ``M_u2n [Hn+1, Hs+1]``.  They are built here by B-spline
collocation with the FITPACK knot placement for even degree (interior knots at the
midpoints between data sites, none in the first and last interval = "not-a-knot").
"""
from __future__ import annotations

import numpy as np


def _knots(x: np.ndarray) -> np.ndarray:
    m = len(x)
    if m < 3:
        raise ValueError("quadratic spline needs at least 3 points")
    interior = 0.5 * (x[1:m - 2] + x[2:m - 1])
    return np.concatenate([[x[0]] * 3, interior, [x[-1]] * 3])


def _basis(t: np.ndarray, xq: np.ndarray, k: int = 2) -> np.ndarray:
    """All B-spline basis functions of degree k on knots t at points xq (Cox-de Boor).
    Points outside [t[k], t[-k-1]] are evaluated with the end polynomial pieces
    (extrapolation), as FITPACK's ``splev`` does with ext=0."""
    n = len(t) - k - 1
    xq = np.asarray(xq, dtype=np.float64)
    # interval index i with t[i] <= x < t[i+1], clamped to the valid spans
    span = np.clip(np.searchsorted(t, xq, side="right") - 1, k, n - 1)
    B = np.zeros((len(xq), n))
    for q, (xv, i) in enumerate(zip(xq, span)):
        N = np.zeros(k + 1)
        N[0] = 1.0
        left = np.zeros(k + 1)
        right = np.zeros(k + 1)
        for j in range(1, k + 1):
            left[j] = xv - t[i + 1 - j]
            right[j] = t[i + j] - xv
            saved = 0.0
            for r in range(j):
                tmp = N[r] / (right[r + 1] + left[j - r])
                N[r] = saved + right[r + 1] * tmp
                saved = left[j - r] * tmp
            N[j] = saved
        B[q, i - k:i + 1] = N
    return B


def interp_matrix(x_from, x_to) -> np.ndarray:
    """Matrix M with  spline(x_to) = M @ y(x_from)  (shape [len(x_to), len(x_from)])."""
    x_from = np.asarray(x_from, dtype=np.float64)
    t = _knots(x_from)
    A = _basis(t, x_from)
    E = _basis(t, np.asarray(x_to, dtype=np.float64))
    return E @ np.linalg.inv(A)
