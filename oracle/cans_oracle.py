"""CPU restatement of the CaNS FFT-based Poisson/Helmholtz solver path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product path (`cans_b200`) never does.

Parity status: PINNED by the reference's only stored result.  The reference cannot be built in
this image (no gfortran / MPI / FFTW3) and holds no stored single-solve vectors, but its regression
vector `tests/lid_driven_cavity/data_ldc_re1000.txt` (64 FP64 values after 1500 time steps = 4500 calls of
`solver`, compared at rtol = 1e-7 by `tests/lid_driven_cavity/test.py:8`) is reproduced by
`oracle/ldc_replay.py` -- the explicit CaNS time loop restated around THIS file's `solver` -- to
3.6e-8 relative, i.e. to the 8 printed digits (`tests/test_ldc_golden.py`; the same replay with the CUDA
solver: `tests/test_gpu_ldc.py`).  That pins R2HC/HC2R, REDFT10/01 and gaussel (Neumann-Neumann, with the
singular-pivot pin).  The kinds the golden case does not reach (RODFT*, REDFT00/11, periodic z, stretched z,
Helmholtz) are pinned by the reference's in-binary self-test properties (`src/sanity.f90:206-409`:
post-correction divergence < `small`, Helmholtz residual < `small`; `tests/test_oracle.py`) and by
brute-force evaluation of FFTW's published r2r definitions.

Third-party arithmetic: FFTW3 (system package, version unpinned by the
reference: `.github/actions/build/scripts/install-GNU.sh:5`).  Its r2r kinds
are restated with `scipy.fft` (pocketfft), norm=None, which implement the
same published definitions (FFTW manual, "1d Real-even DFTs (DCTs)" /
"1d Real-odd DFTs (DSTs)" / "The Halfcomplex-format DFT").

Array convention: every 3-D field is a C-ordered numpy array indexed [k, j, i],
which is the memory layout of the Fortran array p(i,j,k) (x fastest).  Fields
with halos have one ghost cell on every side: shape (n3+2, n2+2, n1+2).
`bc` arguments are 2-D lists bc[idir][ibound] (idir 0..2 = x,y,z), i.e. the
transpose of the Fortran `cbc(0:1,3)`.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as _sfft

# FFTW r2r kind constants, src/fftw.f90:66-85
FFTW_R2HC, FFTW_HC2R = 0, 1
FFTW_REDFT00, FFTW_REDFT01, FFTW_REDFT10, FFTW_REDFT11 = 3, 4, 5, 6
FFTW_RODFT00, FFTW_RODFT01, FFTW_RODFT10, FFTW_RODFT11 = 7, 8, 9, 10


def small(dtype=np.float64):
    """`small` of src/param.f90:18 = epsilon*10**(precision/2)."""
    fi = np.finfo(dtype)
    return float(fi.eps) * 10 ** (fi.precision // 2)


# --------------------------------------------------------------------------
# grid, src/initgrid.f90:14-118
# --------------------------------------------------------------------------
def _gridpoint(gtype, alpha, z0):
    """src/initgrid.f90:105-165 (the four tanh clusterings)."""
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


def initgrid(gtype, n, gr, lz, is_periodic=False, dtype=np.float64):
    """src/initgrid.f90:14-99.  Returns dzc, dzf, zc, zf, each indexed 0..n+1."""
    zf = np.zeros(n + 2)
    for k in range(1, n + 1):
        zf[k] = _gridpoint(gtype, gr, (k - 0.0) / (1.0 * n))
    zf[1:n + 1] *= lz
    dzf = np.zeros(n + 2)
    dzc = np.zeros(n + 2)
    if abs(gr) < np.finfo(dtype).eps:
        dzf[:] = lz / (1.0 * n)
        dzc[:] = lz / (1.0 * n)
    else:
        for k in range(1, n + 1):
            dzf[k] = zf[k] - zf[k - 1]
        if not is_periodic:
            dzf[0], dzf[n + 1] = dzf[1], dzf[n]
        else:
            dzf[0], dzf[n + 1] = dzf[n], dzf[1]
        for k in range(0, n + 1):
            dzc[k] = 0.5 * (dzf[k] + dzf[k + 1])
        dzc[n + 1] = dzc[n] if not is_periodic else dzc[1]
    zc = np.zeros(n + 2)
    zc[0] = -dzc[0] / 2.0
    zf[0] = 0.0
    for k in range(1, n + 2):
        zc[k] = zc[k - 1] + dzc[k - 1]
        zf[k] = zf[k - 1] + dzf[k]
    return dzc.astype(dtype), dzf.astype(dtype), zc.astype(dtype), zf.astype(dtype)


# --------------------------------------------------------------------------
# initsolver pieces, src/initsolver.f90
# --------------------------------------------------------------------------
def eigenvalues(n, bc, c_or_f, dtype=np.float64):
    """src/initsolver.f90:85-144 (CPU build: halfcomplex order for 'PP')."""
    pi = np.arccos(dtype(-1.0))
    lam = np.zeros(n, dtype=dtype)
    key = bc[0] + bc[1]
    two, one = dtype(2.0), dtype(1.0)
    for l in range(1, n + 1):
        if key == 'PP':
            lam[l - 1] = -two * (one - np.cos(dtype(2 * (l - 1)) * pi / dtype(1.0 * n)))
        elif key == 'NN':
            lam[l - 1] = -two * (one - np.cos(dtype(l - 1) * pi / dtype(1.0 * n)))
        elif key == 'DD':
            if c_or_f == 'c':
                lam[l - 1] = -two * (one - np.cos(dtype(l) * pi / dtype(1.0 * n)))
            else:
                lam[l - 1] = -two * (one - np.cos(dtype(l) * pi / dtype(1.0 * n))) if l < n else 0.0
        else:  # ND, DN
            lam[l - 1] = -two * (one - np.cos(dtype(2 * l - 1) * pi / dtype(2.0 * n)))
    return lam


def tridmatrix(bc, n, dzci, dzfi, c_or_f, dtype=np.float64):
    """src/initsolver.f90:146-187.  dzci, dzfi indexed 0..n+1."""
    a = np.zeros(n, dtype=dtype)
    c = np.zeros(n, dtype=dtype)
    for k in range(1, n + 1):
        if c_or_f == 'c':
            a[k - 1] = dzfi[k] * dzci[k - 1]
            c[k - 1] = dzfi[k] * dzci[k]
        else:
            a[k - 1] = dzfi[k] * dzci[k]
            c[k - 1] = dzfi[k + 1] * dzci[k]
    b = -(a + c)
    factor = [{'P': 0.0, 'D': -1.0, 'N': 1.0}[bc[i]] for i in (0, 1)]
    if c_or_f == 'c':
        b[0] = b[0] + dtype(factor[0]) * a[0]
        b[n - 1] = b[n - 1] + dtype(factor[1]) * c[n - 1]
    else:
        if bc[0] == 'N':
            b[0] = b[0] + dtype(factor[0]) * a[0]
        if bc[1] == 'N':
            b[n - 1] = b[n - 1] + dtype(factor[1]) * c[n - 1]
    return a, b, c


def bc_rhs(cbc, bc, dlc, dlf, c_or_f):
    """src/initsolver.f90:189-232.  Returns the two scalars rhs[0], rhs[1]
    (the reference broadcasts them over a boundary plane)."""
    factor = [0.0, 0.0]
    for ib in (0, 1):
        sgn = 1.0 if ib == 0 else -1.0
        if cbc[ib] == 'P':
            factor[ib] = 0.0
        elif cbc[ib] == 'D':
            factor[ib] = -2.0 * bc[ib] if c_or_f == 'c' else -bc[ib]
        else:
            factor[ib] = sgn * (dlc[ib] if c_or_f == 'c' else dlf[ib]) * bc[ib]
    return [factor[ib] / dlc[ib] / dlf[ib] for ib in (0, 1)]


def find_fft(bc, c_or_f):
    """src/fft.f90:260-313 -> (kind_fwd, kind_bwd, norm(2))."""
    key = bc[0] + bc[1]
    if key == 'PP':
        return FFTW_R2HC, FFTW_HC2R, (1.0, 0.0)
    if c_or_f == 'c':
        return {'NN': (FFTW_REDFT10, FFTW_REDFT01, (2.0, 0.0)),
                'DD': (FFTW_RODFT10, FFTW_RODFT01, (2.0, 0.0)),
                'ND': (FFTW_REDFT11, FFTW_REDFT11, (2.0, 0.0)),
                'DN': (FFTW_RODFT11, FFTW_RODFT11, (2.0, 0.0))}[key]
    return {'NN': (FFTW_REDFT00, FFTW_REDFT00, (2.0, -1.0)),
            'DD': (FFTW_RODFT00, FFTW_RODFT00, (2.0, 1.0)),
            'ND': (FFTW_REDFT10, FFTW_REDFT01, (2.0, 0.0)),
            'DN': (FFTW_RODFT01, FFTW_RODFT10, (2.0, 0.0))}[key]


class R2RPlan:
    """What one `fftw_plan_guru_r2r` of src/fft.f90:83-97,148-162 holds:
    transform kind, logical length `n - i` and the axis it runs along."""

    def __init__(self, kind, n, axis):
        self.kind, self.n, self.axis = kind, n, axis


def fftini(ng, bcxy, c_or_f, dtype=np.float64):
    """src/fft.f90:25-209 (CPU branch).  Returns arrplan[dir][fwd|bwd], normfft.
    arrplan[0] = (fwd_x, bwd_x), arrplan[1] = (fwd_y, bwd_y)."""
    arrplan = []
    normfft = dtype(1.0)
    for idir in (0, 1):
        kf, kb, norm = find_fft(bcxy[idir], c_or_f[idir])
        ii = 1 if (bcxy[idir][0] + bcxy[idir][1] == 'DD' and c_or_f[idir] == 'f') else 0
        axis = 2 - idir  # arrays are [k, j, i]
        arrplan.append((R2RPlan(kf, ng[idir] - ii, axis), R2RPlan(kb, ng[idir] - ii, axis)))
        normfft = normfft * dtype(norm[0]) * dtype(ng[idir] + norm[1] - ii)
    return arrplan, dtype(1.0) / normfft


def r2r_1d(x, kind, axis=-1, workers=1):
    """One unnormalised FFTW r2r transform along `axis` (FFTW manual §4.8)."""
    n = x.shape[axis]
    if kind == FFTW_R2HC:
        X = _sfft.rfft(x, axis=axis, workers=workers)
        Xm = np.moveaxis(X, axis, -1)
        out = np.empty(Xm.shape[:-1] + (n,), dtype=x.dtype)
        out[..., :n // 2 + 1] = Xm.real
        nim = (n + 1) // 2 - 1
        if nim > 0:
            out[..., n - nim:] = Xm.imag[..., nim:0:-1]
        return np.moveaxis(out, -1, axis)
    if kind == FFTW_HC2R:
        xm = np.moveaxis(x, axis, -1)
        ctype = np.complex64 if x.dtype == np.float32 else np.complex128
        X = np.zeros(xm.shape[:-1] + (n // 2 + 1,), dtype=ctype)
        X.real = xm[..., :n // 2 + 1]
        nim = (n + 1) // 2 - 1
        if nim > 0:
            X.imag[..., 1:nim + 1] = xm[..., n - 1:n - nim - 1:-1]
        out = _sfft.irfft(X, n=n, axis=-1, workers=workers) * x.dtype.type(n)
        return np.moveaxis(out.astype(x.dtype, copy=False), -1, axis)
    table = {FFTW_REDFT00: (_sfft.dct, 1), FFTW_REDFT10: (_sfft.dct, 2),
             FFTW_REDFT01: (_sfft.dct, 3), FFTW_REDFT11: (_sfft.dct, 4),
             FFTW_RODFT00: (_sfft.dst, 1), FFTW_RODFT10: (_sfft.dst, 2),
             FFTW_RODFT01: (_sfft.dst, 3), FFTW_RODFT11: (_sfft.dst, 4)}
    fn, typ = table[kind]
    return fn(x, type=typ, axis=axis, norm=None, workers=workers).astype(x.dtype, copy=False)


def fft(plan, arr, workers=1):
    """src/fft.f90:247-258: in-place r2r on the first `plan.n` points of
    `plan.axis`; a face-centred Dirichlet line leaves its last point alone."""
    sl = [slice(None)] * arr.ndim
    sl[plan.axis] = slice(0, plan.n)
    sl = tuple(sl)
    arr[sl] = r2r_1d(np.ascontiguousarray(arr[sl]), plan.kind, axis=plan.axis, workers=workers)


def initsolver(ng, dli, dzci_g, dzfi_g, cbc, bc, c_or_f, dtype=np.float64):
    """src/initsolver.f90:15-83 on one rank (lo_z = 1, hi_z = ng).

    Returns dict(lambdaxy[j,i], a, b, c, arrplan, normfft, rhsbx, rhsby, rhsbz)
    with rhsb* = [lower, upper] scalars."""
    dli = [dtype(v) for v in dli]
    lx = eigenvalues(ng[0], cbc[0], c_or_f[0], dtype) * dli[0] ** 2
    ly = eigenvalues(ng[1], cbc[1], c_or_f[1], dtype) * dli[1] ** 2
    lambdaxy = (lx[None, :] + ly[:, None]).astype(dtype)
    a, b, c = tridmatrix(cbc[2], ng[2], dzci_g, dzfi_g, c_or_f[2], dtype)
    dl = [dtype(1.0) / v for v in dli]
    dzc_g = dtype(1.0) / dzci_g
    dzf_g = dtype(1.0) / dzfi_g
    n3 = ng[2]
    rhsbx = bc_rhs(cbc[0], bc[0], [dl[0], dl[0]], [dl[0], dl[0]], c_or_f[0])
    rhsby = bc_rhs(cbc[1], bc[1], [dl[1], dl[1]], [dl[1], dl[1]], c_or_f[1])
    if c_or_f[2] == 'c':
        rhsbz = bc_rhs(cbc[2], bc[2], [dzc_g[0], dzc_g[n3]], [dzf_g[1], dzf_g[n3]], 'c')
    else:
        rhsbz = bc_rhs(cbc[2], bc[2], [dzc_g[1], dzc_g[n3 - 1]], [dzf_g[1], dzf_g[n3]], 'f')
    arrplan, normfft = fftini(ng, [cbc[0], cbc[1]], c_or_f[:2], dtype)
    return dict(lambdaxy=lambdaxy, a=a, b=b, c=c, arrplan=arrplan, normfft=normfft,
                rhsbx=rhsbx, rhsby=rhsby, rhsbz=rhsbz)


# --------------------------------------------------------------------------
# gaussel, src/solver.f90:114-307 -- exact operation order, no contraction
# --------------------------------------------------------------------------
def gaussel(n, a, b, c, is_periodic, norm, p, lambdaxy=None):
    """Thomas solve along axis 0 of p[k, j, i] for rows 0..n-1 (Fortran 1..n),
    in place.  numpy evaluates every elementwise op separately, so the rounding
    sequence equals gfortran -O3 without FMA contraction on x86-64."""
    dt = p.dtype.type
    eps = dt(np.finfo(p.dtype).eps)
    norm = dt(norm)
    nn = n - 1 if is_periodic else n
    if nn < 1:
        raise ValueError("gaussel: empty system")
    lam = lambdaxy if lambdaxy is not None else np.zeros(p.shape[1:], dtype=p.dtype)
    d = np.empty((nn,) + p.shape[1:], dtype=p.dtype)
    one = dt(1.0)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        z = one / (b[0] + lam)
        d[0] = c[0] * z
        p[0] = p[0] * norm * z
        for k in range(1, nn):
            bl = b[k] + lam
            ad = a[k] * d[k - 1]
            den = bl - ad
            z = one / den
            dk = c[k] * z
            pk = (p[k] * norm - a[k] * p[k - 1]) * z
            if k == nn - 1 and lambdaxy is not None:
                tol = eps * np.maximum(np.abs(bl), np.abs(ad))
                pin = np.abs(den) <= tol
                dk = np.where(pin, dt(0.0), dk)
                pk = np.where(pin, dt(0.0), pk)
            d[k] = dk
            p[k] = pk
        for k in range(nn - 2, -1, -1):
            p[k] = p[k] - d[k] * p[k + 1]
        if is_periodic:
            p2 = np.zeros((nn,) + p.shape[1:], dtype=p.dtype)
            p2[0] = -a[0]
            p2[nn - 1] = p2[nn - 1] - c[nn - 1]
            z = one / (b[0] + lam)
            d[0] = c[0] * z
            p2[0] = p2[0] * z
            for k in range(1, nn):
                z = one / (b[k] + lam - a[k] * d[k - 1])
                d[k] = c[k] * z
                p2[k] = (p2[k] - a[k] * p2[k - 1]) * z
            for k in range(nn - 2, -1, -1):
                p2[k] = p2[k] - d[k] * p2[k + 1]
            corr = c[nn] * p2[0] + a[nn] * p2[nn - 1]
            den = ((b[nn] + lam) + c[nn] * p2[0]) + a[nn] * p2[nn - 1]  # left-to-right, :277
            val = (p[nn] * norm - c[nn] * p[0] - a[nn] * p[nn - 1]) / den
            if lambdaxy is not None:
                tol = eps * np.maximum(np.abs(b[nn] + lam), np.abs(corr))
                val = np.where(np.abs(den) <= tol, dt(0.0), val)
            p[nn] = val
            for k in range(nn):
                p[k] = p[k] + p2[k] * p[nn]
    return p


# --------------------------------------------------------------------------
# gaussel_dtdma, src/solver.f90:309-517 -- distributed TDMA (is_poisson_dtdma), P ranks emulated in one process
# --------------------------------------------------------------------------
def gaussel_dtdma(starts, n, a, b, c, is_periodic, norm, p, lambdaxy=None):
    """The reference's distributed tridiagonal solve, restated for a z decomposition with split starts `starts`
    (len P + 1, rank r owns the global rows starts[r] .. starts[r+1]-1; rows >= n -- the face-centred Dirichlet
    plane -- are dropped from the last rank, src/solver.f90:77,91).  `p[k, j, i]` holds rows 0..n-1 of the GLOBAL
    system and is solved in place; `a, b, c` are the global coefficient arrays (every rank sees its own slice,
    src/initsolver.f90:60-65).  Same expression order as the Fortran, every elementwise operation rounded
    separately.  Steps per rank (:351-391): eliminate the inner rows so that they only couple to the rank's
    first and last row; gather the 2 P boundary rows (transpose_y_to_z of aa_y, cc_y, pp_y, :429-440); solve the
    reduced system (:449-461, periodic closure :462-490); hand the boundary values back and update the inner
    rows (:499-510).  Used as the oracle of a future DTDMA kernel; pinned by the dense solve in tests/."""
    dt = p.dtype.type
    one = dt(1.0)
    norm = dt(norm)
    P = len(starts) - 1
    lam = lambdaxy if lambdaxy is not None else np.zeros(p.shape[1:], dtype=p.dtype)
    ranges = [(starts[r], min(starts[r + 1], n)) for r in range(P)]
    if any(k1 - k0 < 3 for k0, k1 in ranges):
        raise ValueError("gaussel_dtdma: every rank needs at least 3 rows")
    AA, CC = [], []
    red_a = np.empty((2 * P,) + p.shape[1:], dtype=p.dtype)
    red_c = np.empty_like(red_a)
    red_p = np.empty_like(red_a)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        for r, (k0, k1) in enumerate(ranges):
            nl = k1 - k0
            al, bl, cl = a[k0:k1], b[k0:k1], c[k0:k1]
            pl = p[k0:k1]
            aa = np.empty((nl,) + p.shape[1:], dtype=p.dtype)
            cc = np.empty_like(aa)
            for k in (0, 1):
                zz = one / (bl[k] + lam)
                aa[k] = al[k] * zz
                cc[k] = cl[k] * zz
                pl[k] = pl[k] * norm * zz
            for k in range(2, nl):        # elimination of lower diagonals
                z = one / ((bl[k] + lam) - al[k] * cc[k - 1])
                pl[k] = (pl[k] * norm - al[k] * pl[k - 1]) * z
                aa[k] = -al[k] * aa[k - 1] * z
                cc[k] = cl[k] * z
            for k in range(nl - 3, 0, -1):  # elimination of upper diagonals
                pl[k] = pl[k] - cc[k] * pl[k + 1]
                aa[k] = aa[k] - cc[k] * aa[k + 1]
                cc[k] = -cc[k] * cc[k + 1]
            z = one / (one - aa[1] * cc[0])
            pl[0] = (pl[0] - cc[0] * pl[1]) * z
            aa[0] = aa[0] * z
            cc[0] = -cc[0] * cc[1] * z
            AA.append(aa)
            CC.append(cc)
            red_a[2 * r], red_a[2 * r + 1] = aa[0], aa[nl - 1]
            red_c[2 * r], red_c[2 * r + 1] = cc[0], cc[nl - 1]
            red_p[2 * r], red_p[2 * r + 1] = pl[0], pl[nl - 1]
        # reduced system: aa_z x_{k-1} + x_k + cc_z x_{k+1} = pp_z
        nn = 2 * P
        cc_z0 = red_c.copy()
        if is_periodic:
            nn -= 1
        for k in range(1, nn):
            z = one / (one - red_a[k] * red_c[k - 1])
            red_p[k] = (red_p[k] - red_a[k] * red_p[k - 1]) * z
            red_c[k] = red_c[k] * z
        for k in range(nn - 2, -1, -1):
            red_p[k] = red_p[k] - red_c[k] * red_p[k + 1]
        if is_periodic:
            cz = cc_z0
            p2 = np.zeros((nn,) + p.shape[1:], dtype=p.dtype)
            p2[0] = -red_a[0]
            p2[nn - 1] = p2[nn - 1] - cz[nn - 1]
            for k in range(1, nn):
                z = one / (one - red_a[k] * cz[k - 1])
                p2[k] = (p2[k] - red_a[k] * p2[k - 1]) * z
                cz[k] = cz[k] * z
            for k in range(nn - 2, -1, -1):
                p2[k] = p2[k] - cz[k] * p2[k + 1]
            red_p[nn] = (red_p[nn] - cz[nn] * red_p[0] - red_a[nn] * red_p[nn - 1]) / \
                        (one + cz[nn] * p2[0] + red_a[nn] * p2[nn - 1])
            for k in range(nn):
                red_p[k] = red_p[k] + p2[k] * red_p[nn]
        # back to the ranks: boundary values, then the inner rows
        for r, (k0, k1) in enumerate(ranges):
            nl = k1 - k0
            pl = p[k0:k1]
            pl[0] = red_p[2 * r]
            pl[nl - 1] = red_p[2 * r + 1]
            for k in range(1, nl - 1):
                pl[k] = pl[k] - AA[r][k] * pl[0] - CC[r][k] * pl[nl - 1]
    return p


# --------------------------------------------------------------------------
# solver / solve_helmholtz, src/solver.f90:17-112, src/solve_helmholtz.f90:28-75
# --------------------------------------------------------------------------
def solver(n, ng, arrplan, normfft, lambdaxy, a, b, c, bc, c_or_f, p, workers=1, stages=None):
    """In-place solve on the interior of the haloed p[k,j,i] (one rank, so all
    four 2DECOMP transposes are copies, transpose_x_to_y.f90:27-36).
    `stages`, if a dict, receives copies of the field after each stage."""
    n1, n2, n3 = n
    px = np.ascontiguousarray(p[1:n3 + 1, 1:n2 + 1, 1:n1 + 1])
    fft(arrplan[0][0], px, workers)
    if stages is not None:
        stages['fwd_x'] = px.copy()
    fft(arrplan[1][0], px, workers)
    if stages is not None:
        stages['fwd_y'] = px.copy()
    q = 1 if (c_or_f[2] == 'f' and bc[2][1] == 'D') else 0
    is_periodic_z = bc[2][0] + bc[2][1] == 'PP'
    gaussel(n3 - q, a, b, c, is_periodic_z, normfft, px, lambdaxy)
    if stages is not None:
        stages['gaussel'] = px.copy()
    fft(arrplan[1][1], px, workers)
    if stages is not None:
        stages['bwd_y'] = px.copy()
    fft(arrplan[0][1], px, workers)
    p[1:n3 + 1, 1:n2 + 1, 1:n1 + 1] = px
    return p


def updt_rhs_b(c_or_f, cbc, n, rhsbx, rhsby, rhsbz, p, alpha=None):
    """src/bound.f90:514-598 on one rank (is_bound all true)."""
    dt = p.dtype.type
    norm = dt(1.0) if alpha is None else dt(alpha)
    q = [1 if (c_or_f[d] == 'f' and cbc[d][1] == 'D') else 0 for d in range(3)]
    n1, n2, n3 = n
    if rhsbx is not None:
        p[1:n3 + 1, 1:n2 + 1, 1] += dt(rhsbx[0]) * norm
        p[1:n3 + 1, 1:n2 + 1, n1 - q[0]] += dt(rhsbx[1]) * norm
    if rhsby is not None:
        p[1:n3 + 1, 1, 1:n1 + 1] += dt(rhsby[0]) * norm
        p[1:n3 + 1, n2 - q[1], 1:n1 + 1] += dt(rhsby[1]) * norm
    if rhsbz is not None:
        p[1, 1:n2 + 1, 1:n1 + 1] += dt(rhsbz[0]) * norm
        p[n3 - q[2], 1:n2 + 1, 1:n1 + 1] += dt(rhsbz[1]) * norm


def solve_helmholtz(n, ng, arrplan, normfft, alpha, lambdaxy, a, b, c, rhsbx, rhsby, rhsbz,
                    cbc, c_or_f, p, workers=1):
    """src/solve_helmholtz.f90:28-75: p/alpha + lap(p) = rhs."""
    dt = p.dtype.type
    updt_rhs_b(c_or_f, cbc, n, rhsbx, rhsby, rhsbz, p, alpha)
    alphai = dt(alpha) ** (-1)
    bb = b + alphai
    return solver(n, ng, arrplan, dt(normfft) * alphai, lambdaxy, a, bb, c, cbc, c_or_f, p, workers)


# --------------------------------------------------------------------------
# the steps either side of the path + acceptance operators
# --------------------------------------------------------------------------
def fillps(n, dli, dzfi, dti, u, v, w, p):
    """src/fillps.f90:13-51."""
    n1, n2, n3 = n
    dt = p.dtype.type
    dti = dt(dti)
    dtidxi = dti * dt(dli[0])
    dtidyi = dti * dt(dli[1])
    K, J, I = slice(1, n3 + 1), slice(1, n2 + 1), slice(1, n1 + 1)
    p[K, J, I] = ((w[K, J, I] - w[0:n3, J, I]) * dti * dzfi[1:n3 + 1, None, None]
                  + (v[K, J, I] - v[K, 0:n2, I]) * dtidyi
                  + (u[K, J, I] - u[K, J, 0:n1]) * dtidxi)


def correc(n, dli, dzci, dt_, p, u, v, w):
    """src/correc.f90:13-60."""
    n1, n2, n3 = n
    ty = p.dtype.type
    fi = ty(dt_) * ty(dli[0])
    fj = ty(dt_) * ty(dli[1])
    u[:, :, 0:n1 + 1] -= fi * (p[:, :, 1:n1 + 2] - p[:, :, 0:n1 + 1])
    v[:, 0:n2 + 1, :] -= fj * (p[:, 1:n2 + 2, :] - p[:, 0:n2 + 1, :])
    w[0:n3 + 1, :, :] -= (ty(dt_) * dzci[0:n3 + 1])[:, None, None] * (p[1:n3 + 2, :, :] - p[0:n3 + 1, :, :])


def _axis_of(idir):
    return 2 - idir


def set_bc(ctype, ibound, idir, centered, rvalue, dr, p):
    """src/bound.f90:182-470 for nh = 1."""
    ax = _axis_of(idir)
    n = p.shape[ax] - 2
    ty = p.dtype.type

    def sl(i):
        s = [slice(None)] * 3
        s[ax] = i
        return tuple(s)

    factor = ty(rvalue)
    sgn = ty(1.0)
    if ctype == 'D' and centered:
        factor = ty(2.0) * factor
        sgn = ty(-1.0)
    if ctype == 'N':
        factor = -ty(dr) * factor if ibound == 0 else ty(dr) * factor
        sgn = ty(1.0)
    if ctype == 'P':
        if ibound == 0:
            p[sl(0)] = p[sl(n)]
        else:
            p[sl(n + 1)] = p[sl(1)]
    elif centered:
        if ibound == 0:
            p[sl(0)] = factor + sgn * p[sl(1)]
        else:
            p[sl(n + 1)] = factor + sgn * p[sl(n)]
    elif ctype == 'D':
        if ibound == 0:
            p[sl(0)] = factor
        else:
            p[sl(n + 1)] = p[sl(n - 1)]
            p[sl(n)] = factor
    elif ctype == 'N':
        if ibound == 0:
            p[sl(0)] = factor + p[sl(1)]
        else:
            p[sl(n + 1)] = p[sl(n)]
            p[sl(n)] = factor + p[sl(n - 1)]


def boundp(cbc, n, bc, dl, dzc, p):
    """src/bound.f90:124-180 on one rank (halo exchange == periodic wrap)."""
    drs = [(dl[0], dl[0]), (dl[1], dl[1]), (dzc[0], dzc[n[2]])]
    for idir in range(3):
        for ib in (0, 1):
            set_bc(cbc[idir][ib], ib, idir, True, bc[idir][ib], drs[idir][ib], p)


def bounduvw(cbcvel, n, bcvel, dl, dzc, dzf, u, v, w, keep_norm_values=False):
    """src/bound.f90:14-122 on one rank.  cbcvel[ivel][idir][ibound]."""
    vel = [u, v, w]
    drs_c = [(dl[0], dl[0]), (dl[1], dl[1]), (dzc[0], dzc[n[2]])]
    drs_n = [(dl[0], dl[0]), (dl[1], dl[1]), (dzf[0], dzf[n[2]])]
    for idir in range(3):
        for ib in (0, 1):
            for ivel in range(3):
                normal = ivel == idir
                ct = cbcvel[ivel][idir][ib]
                if normal:
                    is_pp = cbcvel[ivel][idir][0] + cbcvel[ivel][idir][1] == 'PP'
                    if keep_norm_values and not is_pp:
                        continue
                    set_bc(ct, ib, idir, False, bcvel[ivel][idir][ib], drs_n[idir][ib], vel[ivel])
                else:
                    set_bc(ct, ib, idir, True, bcvel[ivel][idir][ib], drs_c[idir][ib], vel[ivel])


def chkdiv(n, l, dli, dzfi, u, v, w):
    """src/chkdiv.f90:15-54 -> (divtot, divmax)."""
    n1, n2, n3 = n
    K, J, I = slice(1, n3 + 1), slice(1, n2 + 1), slice(1, n1 + 1)
    ty = u.dtype.type
    dxi, dyi = ty(dli[0]), ty(dli[1])
    dz = dzfi[1:n3 + 1, None, None]
    div = ((w[K, J, I] - w[0:n3, J, I]) * dz + (v[K, J, I] - v[K, 0:n2, I]) * dyi
           + (u[K, J, I] - u[K, J, 0:n1]) * dxi)
    divmax = float(np.max(np.abs(div)))
    divtot = float(np.sum(np.abs(div) / (dxi * dyi * dz))) / float(l[0] * l[1] * l[2])
    return divtot, divmax


def chk_helmholtz(n, l, dli, dzci, dzfi, alpha, fp, fpp, bc, c_or_f):
    """src/debug.f90:16-90 -> (restot, resmax)."""
    n1, n2, n3 = n
    ty = fpp.dtype.type
    q = [1 if (bc[d][1] != 'P' and c_or_f[d] == 'f') else 0 for d in range(3)]
    K = slice(1, n3 + 1 - q[2])
    J = slice(1, n2 + 1 - q[1])
    I = slice(1, n1 + 1 - q[0])

    def sh(s, o):
        return slice(s.start + o, s.stop + o)
    dxi, dyi = ty(dli[0]), ty(dli[1])
    c0 = fpp[K, J, I]
    lapx = (fpp[K, J, sh(I, 1)] - ty(2.0) * c0 + fpp[K, J, sh(I, -1)]) * dxi ** 2
    lapy = (fpp[K, sh(J, 1), I] - ty(2.0) * c0 + fpp[K, sh(J, -1), I]) * dyi ** 2
    kk = np.arange(K.start, K.stop)
    if c_or_f[2] == 'c':
        lapz = ((fpp[sh(K, 1), J, I] - c0) * dzci[kk][:, None, None]
                - (c0 - fpp[sh(K, -1), J, I]) * dzci[kk - 1][:, None, None]) * dzfi[kk][:, None, None]
        vol = dzfi[kk][:, None, None]
    else:
        lapz = ((fpp[sh(K, 1), J, I] - c0) * dzfi[kk + 1][:, None, None]
                - (c0 - fpp[sh(K, -1), J, I]) * dzfi[kk][:, None, None]) * dzci[kk][:, None, None]
        vol = dzci[kk][:, None, None]
    val = (c0 + (ty(1.0) / ty(alpha)) * (lapx + lapy + lapz)) * ty(alpha)
    res = np.abs(val - fp[K, J, I])
    return float(np.sum(res / (dxi * dyi * vol))) / float(l[0] * l[1] * l[2]), float(np.max(res))


def chk_poisson(n, l, dli, dzci, dzfi, fp, fpp):
    """src/debug.f90:91-132 -> (restot, resmax)."""
    n1, n2, n3 = n
    ty = fpp.dtype.type
    K, J, I = slice(1, n3 + 1), slice(1, n2 + 1), slice(1, n1 + 1)
    dxi, dyi = ty(dli[0]), ty(dli[1])
    c0 = fpp[K, J, I]
    kk = np.arange(1, n3 + 1)
    val = ((fpp[K, J, 2:n1 + 2] - ty(2.0) * c0 + fpp[K, J, 0:n1]) * dxi ** 2
           + (fpp[K, 2:n2 + 2, I] - ty(2.0) * c0 + fpp[K, 0:n2, I]) * dyi ** 2
           + ((fpp[2:n3 + 2, J, I] - c0) * dzci[kk][:, None, None]
              - (c0 - fpp[0:n3, J, I]) * dzci[kk - 1][:, None, None]) * dzfi[kk][:, None, None])
    res = np.abs(val - fp[K, J, I])
    return float(np.sum(res / (dxi * dyi * dzfi[kk][:, None, None]))) / float(l[0] * l[1] * l[2]), \
        float(np.max(res))


# --------------------------------------------------------------------------
# synthetic inputs shared by tests and bench
# --------------------------------------------------------------------------
def hash_field(ng, seed, dtype=np.float64, lo=(0, 0, 0), n=None):
    """Counter-based uniform(-1,1) field indexed by the GLOBAL (i,j,k): any
    decomposition sees the same numbers.  splitmix64 finaliser; the CUDA twin
    is `cansb200_fill_hash` (cans_b200/csrc/aux_kernels.cu)."""
    n = n or ng
    k = np.arange(lo[2], lo[2] + n[2], dtype=np.uint64)[:, None, None]
    j = np.arange(lo[1], lo[1] + n[1], dtype=np.uint64)[None, :, None]
    i = np.arange(lo[0], lo[0] + n[0], dtype=np.uint64)[None, None, :]
    with np.errstate(over='ignore'):
        idx = (k * np.uint64(ng[1]) + j) * np.uint64(ng[0]) + i
        z = idx + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u01 = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (2.0 * u01 - 1.0).astype(dtype)


def make_case(ng, l, cbc, c_or_f=('c', 'c', 'c'), gr=0.0, gtype=1, bc=None, dtype=np.float64):
    """Everything `initsolver` needs/returns for one rank, in one dict."""
    bc = bc or [[0.0, 0.0]] * 3
    is_pz = cbc[2][0] + cbc[2][1] == 'PP'
    dzc, dzf, zc, zf = initgrid(gtype, ng[2], gr, l[2], is_pz, dtype)
    dli = [dtype(ng[0] / l[0]), dtype(ng[1] / l[1]), dtype(ng[2] / l[2])]
    dzci, dzfi = dtype(1.0) / dzc, dtype(1.0) / dzf
    s = initsolver(ng, dli, dzci, dzfi, cbc, bc, c_or_f, dtype)
    s.update(ng=list(ng), l=list(l), cbc=cbc, c_or_f=list(c_or_f), dli=dli, dzc=dzc, dzf=dzf,
             dzci=dzci, dzfi=dzfi, dtype=dtype, bc=bc)
    return s


# --------------------------------------------------------------------------
# C restatement of gaussel (oracle/gaussel_c.c) -- used by the CPU baseline
# --------------------------------------------------------------------------
_GAUSSEL_C = None


def _load_gaussel_c():
    global _GAUSSEL_C
    if _GAUSSEL_C is None:
        import ctypes as C
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libgaussel_c.so")
        if not os.path.exists(path):
            _GAUSSEL_C = False
        else:
            lib = C.CDLL(path)
            dp = C.POINTER(C.c_double)
            lib.gaussel_c.restype = C.c_int
            lib.gaussel_c.argtypes = [C.c_long, C.c_long, C.c_long, dp, dp, dp, C.c_int, C.c_double, dp, dp]
            _GAUSSEL_C = lib
    return _GAUSSEL_C


def gaussel_c(n, a, b, c, is_periodic, norm, p, lambdaxy):
    """FP64 only; p[k,j,i] C-contiguous.  Bit-identical to `gaussel` (tests/test_oracle.py)."""
    import ctypes as C
    lib = _load_gaussel_c()
    if not lib:
        raise RuntimeError("oracle/_build/libgaussel_c.so missing: run __graft_entry__.build()")
    assert p.dtype == np.float64 and p.flags["C_CONTIGUOUS"]
    dp = C.POINTER(C.c_double)
    f = lambda x: np.ascontiguousarray(x, dtype=np.float64).ctypes.data_as(dp)
    rc = lib.gaussel_c(p.shape[2], p.shape[1], n, f(a), f(b), f(c), int(is_periodic), float(norm), p.ctypes.data_as(dp),
                       f(lambdaxy))
    if rc:
        raise RuntimeError(f"gaussel_c failed: {rc}")
    return p


def solver_gaussel_z(n, ng, hi, a, b, c, bcz, c_or_f, norm, p):
    """src/solver.f90:547-616 on one rank (z not decomposed): the lambda-less `gaussel` on the interior of the
    haloed p[k,j,i], in place (`call gaussel(n(1),n(2),n(3)-q,1,a,b,c,is_periodic_z,norm,p)`, :598)."""
    n1, n2, n3 = n
    q = 1 if (c_or_f[2] == 'f' and bcz[1] == 'D' and hi[2] == ng[2]) else 0
    pz = np.ascontiguousarray(p[1:n3 + 1, 1:n2 + 1, 1:n1 + 1])
    gaussel(n3 - q, a, b, c, bcz[0] + bcz[1] == 'PP', norm, pz, None)
    p[1:n3 + 1, 1:n2 + 1, 1:n1 + 1] = pz
    return p


def solver_fast(n, ng, arrplan, normfft, lambdaxy, a, b, c, bc, c_or_f, p, workers=-1):
    """`solver` with threaded pocketfft and the OpenMP C gaussel: the CPU baseline
    ("kind": "port").  FP64.  Same stages, same order as `solver`."""
    n1, n2, n3 = n
    px = np.ascontiguousarray(p[1:n3 + 1, 1:n2 + 1, 1:n1 + 1])
    fft(arrplan[0][0], px, workers)
    fft(arrplan[1][0], px, workers)
    q = 1 if (c_or_f[2] == 'f' and bc[2][1] == 'D') else 0
    gaussel_c(n3 - q, a, b, c, bc[2][0] + bc[2][1] == 'PP', normfft, px, lambdaxy)
    fft(arrplan[1][1], px, workers)
    fft(arrplan[0][1], px, workers)
    p[1:n3 + 1, 1:n2 + 1, 1:n1 + 1] = px
    return p
