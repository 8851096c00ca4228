"""Host mirror of `initgrid` (/root/reference/src/initgrid.f90:14-165): the stretched z grid
that the tridiagonal coefficients a, b, c are built from.  Host arithmetic run once at start-up,
exactly as in the reference (where it stays Fortran)."""
from __future__ import annotations

import numpy as np


def gridpoint(gtype: int, alpha: float, z0: float) -> float:
    """The tanh clusterings of src/initgrid.f90:105-165 (1 two-end, 2 lower, 3 upper, 4 middle)."""
    if not alpha > np.finfo(np.float64).eps:
        return z0
    if gtype == 2:
        return 1.0 * (1.0 + np.tanh((z0 - 1.0) * alpha) / np.tanh(alpha / 1.0))
    if gtype == 3:
        return 1.0 - 1.0 * (1.0 + np.tanh((1.0 - z0 - 1.0) * alpha) / np.tanh(alpha / 1.0))
    if gtype == 4:
        if z0 <= 0.5:
            return 0.5 * (1.0 - 1.0 + np.tanh(2.0 * alpha * (z0 - 0.0)) / np.tanh(alpha))
        return 0.5 * (1.0 + 1.0 + np.tanh(2.0 * alpha * (z0 - 1.0)) / np.tanh(alpha))
    return 0.5 * (1.0 + np.tanh((z0 - 0.5) * alpha) / np.tanh(alpha / 2.0))


def initgrid(gtype: int, n: int, gr: float, lz: float, is_periodic: bool = False, dtype=np.float64):
    """-> (dzc, dzf), each indexed 0..n+1 (src/initgrid.f90:43-98)."""
    zf = np.zeros(n + 2)
    for k in range(1, n + 1):
        zf[k] = gridpoint(gtype, gr, k / (1.0 * n))
    zf[1:n + 1] *= lz
    dzf = np.zeros(n + 2)
    dzc = np.zeros(n + 2)
    if abs(gr) < np.finfo(dtype).eps:
        dzf[:] = lz / (1.0 * n)
        dzc[:] = lz / (1.0 * n)
    else:
        dzf[1:n + 1] = zf[1:n + 1] - zf[0:n]
        if not is_periodic:
            dzf[0], dzf[n + 1] = dzf[1], dzf[n]
        else:
            dzf[0], dzf[n + 1] = dzf[n], dzf[1]
        dzc[0:n + 1] = 0.5 * (dzf[0:n + 1] + dzf[1:n + 2])
        dzc[n + 1] = dzc[n] if not is_periodic else dzc[1]
    return dzc.astype(dtype), dzf.astype(dtype)
