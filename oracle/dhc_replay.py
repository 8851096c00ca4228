"""Replay of the reference's second regression case, the differentially heated cavity, through the oracle's solver.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as cans_oracle.py).

`tests/differentially_heated_cavity/` of the reference holds a known answer: the Nusselt number at the cold wall after
10 000 time steps of `input.nml` (Ra = 1e6, Pr = 0.71, 128 x 2 x 128), `nusselt_ref = 8.8252` with rtol = atol = 1e-2
(`tests/differentially_heated_cavity/test.py:19-20`).  It is reached through the Navier-Stokes loop with one transported
scalar and Boussinesq buoyancy, explicit diffusion, one rank.  On the solver path it pins what the lid-driven cavity does
not: **REDFT10 / REDFT01 along x** (n = 128, Neumann-Neumann pressure in x) next to R2HC (n = 2) in y and the Neumann-
Neumann Thomas solve (n = 128) in z, through 30 000 solves.  A weak pin (1 %), but a reference-held one.

Restated on top of ldc_replay.py (momentum terms, chkdt) and cans_oracle.py (boundary conditions, fillps, correc):

    main loop            src/main.f90:416-419,429-505 (scalar step before the momentum step, :441-454)
    rk  (explicit)       src/rk.f90:24-186            (buoyancy: :163-177, the branch without _LOOP_UNSWITCHING)
    rk_scal (explicit)   src/rk.f90:335-491
    scal                 src/scal.f90:41-101          (advection + diffusion of the scalar)
    initscal 'dhc'       src/initflow.f90:327-337     (linear profile between the two x walls)
    Nusselt number       tests/differentially_heated_cavity/test.py:5-13

`run_dhc(solve=...)` takes the Poisson solve as a callable like `ldc_replay.run_ldc`.
"""
from __future__ import annotations

import numpy as np

from . import cans_oracle as O
from .ldc_replay import RKCOEFF, _sh, chkdt, mom_xyz_ad

# tests/differentially_heated_cavity/input.nml
DHC = dict(
    ng=[128, 2, 128], l=[1.0, 0.015625, 1.0], gtype=1, gr=0.0, cfl=0.95, dtmax=1.0e9, visci=1186.78165819, nstep=10000,
    icheck=10,
    cbcvel=[[["D", "D"], ["P", "P"], ["D", "D"]]] * 3, bcvel=[[[0.0, 0.0]] * 3] * 3,
    cbcpre=[["N", "N"], ["P", "P"], ["N", "N"]], bcpre=[[0.0, 0.0]] * 3,
    gacc=[0.0, 0.0, -1.0], beta=1.0, alphai=842.614977318,
    cbcscal=[["D", "D"], ["P", "P"], ["N", "N"]], bcscal=[[-0.5, 0.5], [0.0, 0.0], [0.0, 0.0]],
)
NUSSELT_REF = 8.8252   # tests/differentially_heated_cavity/test.py:19


def scal(n, dxi, dyi, dzci, dzfi, alpha, u, v, w, s):
    """src/scal.f90:61-101, explicit-diffusion branch -> dsdt on the interior."""
    n3 = n[2]
    K = slice(1, n3 + 1)
    col = lambda a: a[:, None, None]
    dzci_k, dzci_km, dzfi_k = col(dzci[K]), col(dzci[0:n3]), col(dzfi[K])
    sh = lambda f, di=0, dj=0, dk=0: _sh(f, n, di, dj, dk)
    s_c = sh(s)
    usim = 0.5 * (sh(s, -1) + s_c) * sh(u, -1)
    usip = 0.5 * (sh(s, 1) + s_c) * sh(u)
    vsjm = 0.5 * (sh(s, 0, -1) + s_c) * sh(v, 0, -1)
    vsjp = 0.5 * (sh(s, 0, 1) + s_c) * sh(v)
    wskm = 0.5 * (sh(s, 0, 0, -1) + s_c) * sh(w, 0, 0, -1)
    wskp = 0.5 * (sh(s, 0, 0, 1) + s_c) * sh(w)
    dsdxp, dsdxm = (sh(s, 1) - s_c) * dxi, (s_c - sh(s, -1)) * dxi
    dsdyp, dsdym = (sh(s, 0, 1) - s_c) * dyi, (s_c - sh(s, 0, -1)) * dyi
    dsdzp, dsdzm = (sh(s, 0, 0, 1) - s_c) * dzci_k, (s_c - sh(s, 0, 0, -1)) * dzci_km
    dsdt = dxi * (-usip + usim) + dyi * (-vsjp + vsjm) + dzfi_k * (-wskp + wskm)
    dsdtd_xy = (dsdxp - dsdxm) * alpha * dxi + (dsdyp - dsdym) * alpha * dyi
    dsdtd_z = (dsdzp - dsdzm) * alpha * dzfi_k
    return dsdt + dsdtd_xy + dsdtd_z


def nusselt(s, dl, dzf, tw=-0.5, l=1.0):
    """tests/differentially_heated_cavity/test.py:5-13: data[0,0,:] = s(1,1,1:nz), dx = 2 xp[0], dz = 2 zp[0]."""
    col = s[1:-1, 1, 1]
    return float((1.0 / l) * np.sum(((col - tw) / dl[0]) * dzf[1]) * l / (-tw))


def oracle_solve(cs):
    """Poisson solve = the oracle's restatement of src/solver.f90."""
    def solve(pp):
        O.solver(cs["ng"], cs["ng"], cs["arrplan"], cs["normfft"], cs["lambdaxy"], cs["a"], cs["b"], cs["c"], cs["cbc"],
                 cs["c_or_f"], pp)
    return solve


def run_dhc(solve=None, nstep=None, cfg=None, return_state=False):
    """The shipped differentially-heated-cavity case.  `solve(pp)` solves the Poisson equation in place on the haloed
    pp[k,j,i] (default: the oracle).  Returns the Nusselt number test.py computes after `nstep` steps."""
    cfg = dict(DHC, **(cfg or {}))
    ng, l = cfg["ng"], cfg["l"]
    nstep = cfg["nstep"] if nstep is None else nstep
    n = ng
    cs = O.make_case(ng, l, cfg["cbcpre"], gr=cfg["gr"], gtype=cfg["gtype"], bc=cfg["bcpre"])
    if solve is None:
        solve = oracle_solve(cs)
    dl = [l[d] / (1.0 * ng[d]) for d in range(3)]
    dli = [dl[d] ** (-1) for d in range(3)]
    visc = cfg["visci"] ** (-1)
    alpha = cfg["alphai"] ** (-1)                       # s%alpha; alpha_max = 1 / minval(alphai) (src/param.f90:308-310)
    dzc, dzf, dzci, dzfi = cs["dzc"], cs["dzf"], cs["dzci"], cs["dzfi"]
    shp = (n[2] + 2, n[1] + 2, n[0] + 2)
    u, v, w, p, pp, s = (np.zeros(shp) for _ in range(6))   # inivel = 'zer'
    xx = (np.arange(1, n[0] + 1) - 0.5) * dl[0] / l[0]       # initscal 'dhc'
    s[1:-1, 1:-1, 1:-1] = ((1.0 - xx) * cfg["bcscal"][0][0] + xx * cfg["bcscal"][0][1])[None, None, :]
    dudtrko, dvdtrko, dwdtrko, dsdtrko = (np.zeros((n[2], n[1], n[0])) for _ in range(4))
    cbcvel, bcvel, cbcpre, bcpre = cfg["cbcvel"], cfg["bcvel"], cfg["cbcpre"], cfg["bcpre"]
    cbcs, bcs = cfg["cbcscal"], cfg["bcscal"]
    gacc, beta = cfg["gacc"], cfg["beta"]
    O.bounduvw(cbcvel, n, bcvel, dl, dzc, dzf, u, v, w, False)
    O.boundp(cbcpre, n, bcpre, dl, dzc, p)
    O.boundp(cbcs, n, bcs, dl, dzc, s)
    dt = min(cfg["cfl"] * chkdt(n, dl, dzci, dzfi, visc, alpha, u, v, w), cfg["dtmax"])
    I = (slice(1, n[2] + 1), slice(1, n[1] + 1), slice(1, n[0] + 1))
    sh = lambda f, di=0, dj=0, dk=0: _sh(f, n, di, dj, dk)
    dzci_k = dzci[1:n[2] + 1][:, None, None]
    divmax = 0.0
    for istep in range(1, nstep + 1):
        for irk in range(3):
            rkpar = RKCOEFF[irk]
            dtrk = (rkpar[0] + rkpar[1]) * dt
            dtrki = dtrk ** (-1)
            f1, f2 = rkpar[0] * dt, rkpar[1] * dt
            f12 = f1 + f2
            # rk_scal (src/rk.f90:335-491), then the scalar's boundary conditions (src/main.f90:441-454)
            dsdtrk = scal(n, dli[0], dli[1], dzci, dzfi, alpha, u, v, w, s)
            s[I] = s[I] + f1 * dsdtrk + f2 * dsdtrko + f12 * 0.0
            dsdtrko[...] = dsdtrk
            O.boundp(cbcs, n, bcs, dl, dzc, s)
            # rk (src/rk.f90:24-186) with the Boussinesq term
            dudtrk, dvdtrk, dwdtrk = mom_xyz_ad(n, dli[0], dli[1], dzci, dzfi, visc, u, v, w)
            u[I] = u[I] + f1 * dudtrk + f2 * dudtrko + f12 * (0.0 - dli[0] * (sh(p, 1) - sh(p)))
            v[I] = v[I] + f1 * dvdtrk + f2 * dvdtrko + f12 * (0.0 - dli[1] * (sh(p, 0, 1) - sh(p)))
            w[I] = w[I] + f1 * dwdtrk + f2 * dwdtrko + f12 * (0.0 - dzci_k * (sh(p, 0, 0, 1) - sh(p)))
            if gacc[0] != 0.0:
                u[I] = u[I] - f12 * gacc[0] * beta * 0.5 * (sh(s, 1) + sh(s))
            if gacc[1] != 0.0:
                v[I] = v[I] - f12 * gacc[1] * beta * 0.5 * (sh(s, 0, 1) + sh(s))
            if gacc[2] != 0.0:
                w[I] = w[I] - f12 * gacc[2] * beta * 0.5 * (sh(s, 0, 0, 1) + sh(s))
            dudtrko[...] = dudtrk
            dvdtrko[...] = dvdtrk
            dwdtrko[...] = dwdtrk
            O.bounduvw(cbcvel, n, bcvel, dl, dzc, dzf, u, v, w, False)
            O.fillps(n, dli, dzfi, dtrki, u, v, w, pp)
            solve(pp)                                    # rhsbp = 0: every pressure boundary value is zero
            O.boundp(cbcpre, n, bcpre, dl, dzc, pp)
            O.correc(n, dli, dzci, dtrk, pp, u, v, w)
            O.bounduvw(cbcvel, n, bcvel, dl, dzc, dzf, u, v, w, True)
            p[I] = p[I] + pp[I]                          # updatep, explicit branch
            O.boundp(cbcpre, n, bcpre, dl, dzc, p)
        if cfg["icheck"] > 0 and istep % cfg["icheck"] == 0:
            dt = min(cfg["cfl"] * chkdt(n, dl, dzci, dzfi, visc, alpha, u, v, w), cfg["dtmax"])
            _, dm = O.chkdiv(n, l, dli, dzfi, u, v, w)
            divmax = max(divmax, dm)
    nu = nusselt(s, dl, dzf)
    if return_state:
        return nu, dict(u=u, v=v, w=w, p=p, s=s, divmax=divmax, dt=dt)
    return nu
