"""Shared helpers for the parity tests (CUDA path vs oracle)."""
from __future__ import annotations

import numpy as np

TOL = {"float": 1e-5, "double": 1e-12}          # north_star: normwise, justified by FMA contraction


def normwise(a: np.ndarray, b: np.ndarray) -> float:
    """max|a-b| / max|b|  (pointwise-relative explodes at zero crossings)."""
    b64 = b.astype(np.float64)
    den = float(np.max(np.abs(b64))) if b.size else 0.0
    if den == 0.0:
        den = 1.0
    return float(np.max(np.abs(a.astype(np.float64) - b64))) / den if b.size else 0.0


def rotate(rot: int, cur: list, idxs: list):
    """The reference driver's pointer rotation (laplacian.c:299-300, wave13pt.c:919-920)."""
    if rot == 2:
        cur[0], cur[1] = cur[1], cur[0]
        idxs[0], idxs[1] = idxs[1], idxs[0]
    elif rot == 3:
        cur[:] = [cur[1], cur[2], cur[0]]
        idxs[:] = [idxs[1], idxs[2], idxs[0]]


def uxx1_bound(scalars, arrays, nx, ny, ns, real: str, nt: int) -> np.ndarray:
    """Condition-aware per-point error bound for uxx1: the update divides by
    d = 0.25*(4 nearly cancelling d1 terms).  With kappa = sum|d1 terms| / |sum d1 terms| and
    S = |u0| + (dth/|d|) * sum|c*(term)| the admissible error is  tol * nt * (1 + kappa) * S."""
    c1, c2 = scalars
    u0, u1, d1, xx, xy, xz = [a.astype(np.float64).reshape(ns, ny, nx) for a in arrays]
    k, j, i = slice(2, ns - 1), slice(2, ny - 1), slice(2, nx - 1)

    def sh(a, dk=0, dj=0, di=0):
        return a[2 + dk:ns - 1 + dk, 2 + dj:ny - 1 + dj, 2 + di:nx - 1 + di]

    terms = [sh(d1), sh(d1, dj=-1), sh(d1, dk=-1), sh(d1, dk=-1, dj=-1)]
    ssum = np.abs(sum(terms))
    kappa = sum(np.abs(t) for t in terms) / np.maximum(ssum, 1e-300)
    d = 0.25 * ssum
    dth = 1.0 / nx
    mags = (abs(c1) * (np.abs(sh(xx)) + np.abs(sh(xx, di=-1))) + abs(c2) * (np.abs(sh(xx, di=1)) + np.abs(sh(xx, di=-2))) +
            abs(c1) * (np.abs(sh(xy)) + np.abs(sh(xy, dj=-1))) + abs(c2) * (np.abs(sh(xy, dj=1)) + np.abs(sh(xy, dj=-2))) +
            abs(c1) * (np.abs(sh(xz)) + np.abs(sh(xz, dk=-1))) + abs(c2) * (np.abs(sh(xz, dk=1)) + np.abs(sh(xz, dk=-2))))
    S = np.maximum(np.abs(sh(u0)), np.abs(sh(u1))) + (dth / np.maximum(d, 1e-300)) * mags
    bound = np.zeros((ns, ny, nx))
    bound[k, j, i] = TOL[real] * max(nt, 1) * (1.0 + kappa) * S
    return bound.reshape(-1)
