"""Replay of the reference's only golden vector through the oracle's solver.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as cans_oracle.py).

The reference pins its solver with ONE stored result: `tests/lid_driven_cavity/
data_ldc_re1000.txt` (64 values of v along z at x = 1, y = ny/2 + 1 after 1500 time steps of
`tests/lid_driven_cavity/input.nml`, compared with rtol = 1e-7 in `tests/lid_driven_cavity/test.py:8`).
That vector is reached only through the full Navier-Stokes loop, so this file restates the part of
CaNS's time loop that the shipped input exercises (explicit diffusion, no forcing, no scalars, one rank):

    main loop            src/main.f90:416-419,429-505   (dt from chkdt every `icheck` steps)
    rk  (explicit)       src/rk.f90:24-119,168-196,261-273
    mom_xyz_ad           src/mom.f90:771-1019           (advection "VV" form + diffusion, one kernel)
    chkdt                src/chkdt.f90:16-77
    updatep (explicit)   src/updatep.f90:75-86
    bounduvw / boundp    src/bound.f90:17-180           (via cans_oracle.bounduvw / boundp)
    fillps / correc      src/fillps.f90, src/correc.f90 (via cans_oracle)
    solver               src/solver.f90:17-112          (the function under test; injectable)

`run_ldc(solve=...)` takes the Poisson solve as a callable, so the same replay pins
  * the oracle (`cans_oracle.solver`)  -> tests/test_ldc_golden.py (CPU), and
  * the CUDA path through the C ABI   -> tests/test_gpu_ldc.py (GPU, 4500 solves of the 2x64x64 grid).
A copy of the 64 golden numbers lives in tests/golden/ldc_re1000_ref.txt (data, not code; the reference
tree does not exist on the GPU box).
"""
from __future__ import annotations

import numpy as np

from . import cans_oracle as O

# src/param.f90:20-22
RKCOEFF = np.array([[32.0 / 60.0, 0.0], [25.0 / 60.0, -17.0 / 60.0], [45.0 / 60.0, -25.0 / 60.0]])

# tests/lid_driven_cavity/input.nml
LDC = dict(
    ng=[2, 64, 64], l=[0.03125, 1.0, 1.0], gtype=1, gr=0.0, cfl=0.95, dtmax=1.0e5, visci=1000.0, nstep=1500, icheck=10,
    # cbcvel[ivel][idir][ibound], bcvel likewise (transpose of the Fortran cbcvel(0:1,1:3,ivel))
    cbcvel=[[["P", "P"], ["D", "D"], ["D", "D"]]] * 3,
    bcvel=[[[0.0, 0.0], [0.0, 0.0], [0.0, 0.0]], [[0.0, 0.0], [0.0, 0.0], [0.0, 1.0]], [[0.0, 0.0], [0.0, 0.0], [0.0, 0.0]]],
    cbcpre=[["P", "P"], ["N", "N"], ["N", "N"]], bcpre=[[0.0, 0.0]] * 3,
)


def _sh(f, n, di=0, dj=0, dk=0):
    """interior of the haloed field shifted by (di, dj, dk): f(i+di, j+dj, k+dk), i=1..n1 etc."""
    n1, n2, n3 = n
    return f[1 + dk:n3 + 1 + dk, 1 + dj:n2 + 1 + dj, 1 + di:n1 + 1 + di]


def mom_xyz_ad(n, dxi, dyi, dzci, dzfi, visc, u, v, w):
    """src/mom.f90:771-1019, explicit-diffusion branch (:1008-1015): returns (dudt, dvdt, dwdt) on the interior."""
    n3 = n[2]
    K = slice(1, n3 + 1)
    col = lambda a: a[:, None, None]
    dzci_k, dzci_km = col(dzci[K]), col(dzci[0:n3])
    dzfi_k, dzfi_kp = col(dzfi[K]), col(dzfi[2:n3 + 2])
    rdzm = col(dzci[K] / dzfi[K])            # :804-807
    rdzp = col(dzci[K] / dzfi[2:n3 + 2])
    s = lambda f, di=0, dj=0, dk=0: _sh(f, n, di, dj, dk)
    u_ccm, u_pcm, u_cmc, u_pmc = s(u, 0, 0, -1), s(u, 1, 0, -1), s(u, 0, -1, 0), s(u, 1, -1, 0)
    u_mcc, u_ccc, u_pcc, u_mpc, u_cpc = s(u, -1), s(u), s(u, 1), s(u, -1, 1, 0), s(u, 0, 1, 0)
    u_mcp, u_ccp = s(u, -1, 0, 1), s(u, 0, 0, 1)
    v_ccm, v_cpm, v_cmc, v_pmc = s(v, 0, 0, -1), s(v, 0, 1, -1), s(v, 0, -1, 0), s(v, 1, -1, 0)
    v_mcc, v_ccc, v_pcc, v_cpc = s(v, -1), s(v), s(v, 1), s(v, 0, 1, 0)
    v_cmp, v_ccp = s(v, 0, -1, 1), s(v, 0, 0, 1)
    w_ccm, w_pcm, w_cpm, w_cmc = s(w, 0, 0, -1), s(w, 1, 0, -1), s(w, 0, 1, -1), s(w, 0, -1, 0)
    w_mcc, w_ccc, w_pcc, w_cpc, w_ccp = s(w, -1), s(w), s(w, 1), s(w, 0, 1, 0), s(w, 0, 0, 1)
    # x momentum (:893-923)
    dudxp, dudxm = (u_pcc - u_ccc) * dxi, (u_ccc - u_mcc) * dxi
    dudyp, dudym = (u_cpc - u_ccc) * dyi, (u_ccc - u_cmc) * dyi
    dudzp, dudzm = (u_ccp - u_ccc) * dzci_k, (u_ccc - u_ccm) * dzci_km
    uuip, uuim = 0.25 * (u_ccc + u_pcc) * u_pcc, 0.25 * (u_ccc + u_mcc) * u_mcc
    vujp, vujm = 0.25 * (v_ccc + v_pcc) * u_cpc, 0.25 * (v_cmc + v_pmc) * u_cmc
    wukp, wukm = 0.25 * (w_ccc + w_pcc) * u_ccp, 0.25 * (w_ccm + w_pcm) * u_ccm
    dudtd_xy = visc * (dudxp - dudxm) * dxi + visc * (dudyp - dudym) * dyi
    dudtd_z = visc * (dudzp - dudzm) * dzfi_k
    dudt = -(uuip - uuim) * dxi - (vujp - vujm) * dyi - (wukp - wukm) * dzfi_k
    # y momentum (:927-957)
    dvdxp, dvdxm = (v_pcc - v_ccc) * dxi, (v_ccc - v_mcc) * dxi
    dvdyp, dvdym = (v_cpc - v_ccc) * dyi, (v_ccc - v_cmc) * dyi
    dvdzp, dvdzm = (v_ccp - v_ccc) * dzci_k, (v_ccc - v_ccm) * dzci_km
    uvip, uvim = 0.25 * (u_ccc + u_cpc) * v_pcc, 0.25 * (u_mcc + u_mpc) * v_mcc
    vvjp, vvjm = 0.25 * (v_ccc + v_cpc) * v_cpc, 0.25 * (v_ccc + v_cmc) * v_cmc
    wvkp, wvkm = 0.25 * (w_ccc + w_cpc) * v_ccp, 0.25 * (w_ccm + w_cpm) * v_ccm
    dvdtd_xy = visc * (dvdxp - dvdxm) * dxi + visc * (dvdyp - dvdym) * dyi
    dvdtd_z = visc * (dvdzp - dvdzm) * dzfi_k
    dvdt = -(uvip - uvim) * dxi - (vvjp - vvjm) * dyi - (wvkp - wvkm) * dzfi_k
    # z momentum (:961-991)
    dwdxp, dwdxm = (w_pcc - w_ccc) * dxi, (w_ccc - w_mcc) * dxi
    dwdyp, dwdym = (w_cpc - w_ccc) * dyi, (w_ccc - w_cmc) * dyi
    dwdzp, dwdzm = (w_ccp - w_ccc) * dzfi_kp, (w_ccc - w_ccm) * dzfi_k
    uwip, uwim = 0.25 * (rdzm * u_ccc + rdzp * u_ccp) * w_pcc, 0.25 * (rdzm * u_mcc + rdzp * u_mcp) * w_mcc
    vwjp, vwjm = 0.25 * (rdzm * v_ccc + rdzp * v_ccp) * w_cpc, 0.25 * (rdzm * v_cmc + rdzp * v_cmp) * w_cmc
    wwkp, wwkm = 0.25 * (w_ccc + w_ccp) * w_ccp, 0.25 * (w_ccc + w_ccm) * w_ccm
    dwdtd_xy = visc * (dwdxp - dwdxm) * dxi + visc * (dwdyp - dwdym) * dyi
    dwdtd_z = visc * (dwdzp - dwdzm) * dzci_k
    dwdt = -(uwip - uwim) * dxi - (vwjp - vwjm) * dyi - (wwkp - wwkm) * dzci_k
    return dudt + dudtd_xy + dudtd_z, dvdt + dvdtd_xy + dvdtd_z, dwdt + dwdtd_xy + dwdtd_z


def chkdt(n, dl, dzci, dzfi, visc, alpha, u, v, w):
    """src/chkdt.f90:16-77 -> dtmax (explicit diffusion)."""
    n3 = n[2]
    dxi, dyi = 1.0 / dl[0], 1.0 / dl[1]
    dlmin = min(min(dl[0], dl[1]), float(np.min(1.0 / dzfi)))
    s = lambda f, di=0, dj=0, dk=0: _sh(f, n, di, dj, dk)
    dzf_k = dzfi[1:n3 + 1][:, None, None]
    dzc_k = dzci[1:n3 + 1][:, None, None]
    ux = np.abs(s(u))
    vx = 0.25 * np.abs(s(v) + s(v, 0, -1) + s(v, 1) + s(v, 1, -1))
    wx = 0.25 * np.abs(s(w) + s(w, 0, 0, -1) + s(w, 1) + s(w, 1, 0, -1))
    dtix = ux * dxi + vx * dyi + wx * dzf_k
    uy = 0.25 * np.abs(s(u) + s(u, 0, 1) + s(u, -1, 1) + s(u, -1))
    vy = np.abs(s(v))
    wy = 0.25 * np.abs(s(w) + s(w, 0, 1) + s(w, 0, 1, -1) + s(w, 0, 0, -1))
    dtiy = uy * dxi + vy * dyi + wy * dzf_k
    uz = 0.25 * np.abs(s(u) + s(u, -1) + s(u, -1, 0, 1) + s(u, 0, 0, 1))
    vz = 0.25 * np.abs(s(v) + s(v, 0, -1) + s(v, 0, -1, 1) + s(v, 0, 0, 1))
    wz = np.abs(s(w))
    dtiz = uz * dxi + vz * dyi + wz * dzc_k
    dti = max(0.0, float(dtix.max()), float(dtiy.max()), float(dtiz.max()))
    if dti < np.finfo(np.float64).eps:
        dti = 1.0
    return min(1.65 / 12.0 / max(visc, alpha) * dlmin ** 2, np.sqrt(3.0) / dti)


def rk_explicit(rkpar, n, dli, dzci, dzfi, dt, visc, p, dudtrko, dvdtrko, dwdtrko, u, v, w):
    """src/rk.f90:24-273, is_impdiff = F, no forcing, no buoyancy."""
    n1, n2, n3 = n
    factor1, factor2 = rkpar[0] * dt, rkpar[1] * dt
    factor12 = factor1 + factor2
    dudtrk, dvdtrk, dwdtrk = mom_xyz_ad(n, dli[0], dli[1], dzci, dzfi, visc, u, v, w)
    I = (slice(1, n3 + 1), slice(1, n2 + 1), slice(1, n1 + 1))
    s = lambda f, di=0, dj=0, dk=0: _sh(f, n, di, dj, dk)
    dzci_k = dzci[1:n3 + 1][:, None, None]
    u[I] = u[I] + factor1 * dudtrk + factor2 * dudtrko + factor12 * (0.0 - dli[0] * (s(p, 1) - s(p)))
    v[I] = v[I] + factor1 * dvdtrk + factor2 * dvdtrko + factor12 * (0.0 - dli[1] * (s(p, 0, 1) - s(p)))
    w[I] = w[I] + factor1 * dwdtrk + factor2 * dwdtrko + factor12 * (0.0 - dzci_k * (s(p, 0, 0, 1) - s(p)))
    dudtrko[...] = dudtrk
    dvdtrko[...] = dvdtrk
    dwdtrko[...] = dwdtrk


def oracle_solve(cs):
    """Poisson solve = the oracle's restatement of src/solver.f90."""
    def solve(pp):
        O.solver(cs["ng"], cs["ng"], cs["arrplan"], cs["normfft"], cs["lambdaxy"], cs["a"], cs["b"], cs["c"], cs["cbc"],
                 cs["c_or_f"], pp)
    return solve


def run_ldc(solve=None, nstep=None, cfg=None, return_state=False):
    """The shipped lid-driven-cavity case.  `solve(pp)` solves the Poisson equation in place on the haloed pp[k,j,i]
    (default: the oracle).  Returns v(1, ny/2+1, 1:nz) after `nstep` steps, i.e. what test.py:5-8 extracts."""
    cfg = dict(LDC, **(cfg or {}))
    ng, l = cfg["ng"], cfg["l"]
    nstep = cfg["nstep"] if nstep is None else nstep
    n = ng
    cs = O.make_case(ng, l, cfg["cbcpre"], gr=cfg["gr"], gtype=cfg["gtype"], bc=cfg["bcpre"])
    if solve is None:
        solve = oracle_solve(cs)
    dl = [l[d] / (1.0 * ng[d]) for d in range(3)]       # src/param.f90:196-198
    dli = [dl[d] ** (-1) for d in range(3)]
    visc = cfg["visci"] ** (-1)
    alpha_max = 1.0 / np.finfo(np.float64).max          # nscal = 0: minval of an empty array (src/param.f90:308-310)
    dzc, dzf, dzci, dzfi = cs["dzc"], cs["dzf"], cs["dzci"], cs["dzfi"]
    shp = (n[2] + 2, n[1] + 2, n[0] + 2)
    u, v, w, p, pp = (np.zeros(shp) for _ in range(5))   # inivel = 'zer'
    dudtrko, dvdtrko, dwdtrko = (np.zeros((n[2], n[1], n[0])) for _ in range(3))
    cbcvel, bcvel, cbcpre, bcpre = cfg["cbcvel"], cfg["bcvel"], cfg["cbcpre"], cfg["bcpre"]
    O.bounduvw(cbcvel, n, bcvel, dl, dzc, dzf, u, v, w, False)
    O.boundp(cbcpre, n, bcpre, dl, dzc, p)
    dt = min(cfg["cfl"] * chkdt(n, dl, dzci, dzfi, visc, alpha_max, u, v, w), cfg["dtmax"])
    divmax = 0.0
    for istep in range(1, nstep + 1):
        for irk in range(3):
            dtrk = (RKCOEFF[irk, 0] + RKCOEFF[irk, 1]) * dt
            dtrki = dtrk ** (-1)
            rk_explicit(RKCOEFF[irk], n, dli, dzci, dzfi, dt, visc, p, dudtrko, dvdtrko, dwdtrko, u, v, w)
            O.bounduvw(cbcvel, n, bcvel, dl, dzc, dzf, u, v, w, False)
            O.fillps(n, dli, dzfi, dtrki, u, v, w, pp)
            # updt_rhs_b: every pressure BC value is zero in this case, so rhsb* = 0 (src/initsolver.f90:189-232)
            solve(pp)
            O.boundp(cbcpre, n, bcpre, dl, dzc, pp)
            O.correc(n, dli, dzci, dtrk, pp, u, v, w)
            O.bounduvw(cbcvel, n, bcvel, dl, dzc, dzf, u, v, w, True)
            I = (slice(1, n[2] + 1), slice(1, n[1] + 1), slice(1, n[0] + 1))
            p[I] = p[I] + pp[I]                          # updatep, explicit branch
            O.boundp(cbcpre, n, bcpre, dl, dzc, p)
        if cfg["icheck"] > 0 and istep % cfg["icheck"] == 0:
            dt = min(cfg["cfl"] * chkdt(n, dl, dzci, dzfi, visc, alpha_max, u, v, w), cfg["dtmax"])
            _, dm = O.chkdiv(n, l, dli, dzfi, u, v, w)
            divmax = max(divmax, dm)
    islice = n[2] // 2
    out = v[1:n[2] + 1, 1 + islice, 1].copy()           # data[0, islice, :] of the (i, j, k) array in test.py
    if return_state:
        return out, dict(u=u, v=v, w=w, p=p, divmax=divmax, dt=dt, zc=np.cumsum(dzf[1:n[2] + 1]) - 0.5 * dzf[1:n[2] + 1])
    return out
