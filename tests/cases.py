"""Shared case table for the parity tests (CPU and GPU)."""
import numpy as np

from oracle import cans_oracle as O

P, N, D = ["P", "P"], ["N", "N"], ["D", "D"]
ND, DN = ["N", "D"], ["D", "N"]
C3 = ["c", "c", "c"]

# name -> (ng, l, cbc, c_or_f, gr, dtype)
CASES = {
    # BASELINE.json configs[0]: tests/lid_driven_cavity/input.nml as shipped
    "C1_ldc_2x64x64": ([2, 64, 64], [0.03125, 1.0, 1.0], [P, N, N], C3, 0.0, np.float64),
    # scaled-down twins of configs[1..3]
    "C2s_triperiodic": ([32, 16, 24], [6.2832, 6.2832, 6.2832], [P, P, P], C3, 0.0, np.float64),
    "C3s_channel": ([32, 16, 40], [12.0, 6.0, 2.0], [P, P, N], C3, 2.0, np.float64),
    "C4s_duct": ([16, 24, 48], [12.0, 2.0, 2.0], [P, N, N], C3, 1.5, np.float64),
    # edge cases: odd / prime sizes (direct transform path), tiny z, Dirichlet pressure
    "odd_sizes": ([9, 14, 11], [1.0, 1.0, 1.0], [P, N, N], C3, 1.0, np.float64),
    "prime_17x34": ([34, 17, 6], [1.0, 1.0, 1.0], [N, P, N], C3, 0.0, np.float64),
    "dirichlet_xyz": ([16, 12, 20], [1.0, 1.0, 1.0], [D, D, D], C3, 1.0, np.float64),
    "mixed_nd_dn": ([12, 10, 16], [1.0, 1.0, 1.0], [ND, DN, D], C3, 0.5, np.float64),
    "periodic_z_odd": ([8, 6, 37], [1.0, 1.0, 1.0], [N, P, P], C3, 0.0, np.float64),
    "tiny_z": ([8, 8, 2], [1.0, 1.0, 1.0], [P, P, N], C3, 0.0, np.float64),
    "nz_gt_512": ([8, 8, 600], [1.0, 1.0, 2.0], [P, P, N], C3, 2.0, np.float64),
    # 8-column pipelined Thomas tiles (512 < nz <= 1024) and the sequential fallback beyond
    "nz_768_duct": ([16, 12, 768], [1.0, 1.0, 2.0], [P, N, N], C3, 1.5, np.float64),
    "nz_1024": ([10, 8, 1024], [1.0, 1.0, 2.0], [P, P, N], C3, 2.0, np.float64),
    "nz_1000_periodic": ([8, 6, 1000], [1.0, 1.0, 1.0], [P, P, P], C3, 0.0, np.float64),
    "nz_gt_1024": ([8, 8, 1100], [1.0, 1.0, 2.0], [P, P, D], C3, 1.0, np.float64),
    # shapes that take the TMA tile path of the pipelined Thomas kernel (full chunks, >= 16 even columns);
    # 40 columns = 2.5 tiles (out-of-range columns), 257 periodic rows = 256-row system + closure row
    "nz_512_tma": ([48, 6, 512], [1.0, 1.0, 2.0], [P, P, N], C3, 2.0, np.float64),
    "nz_257_periodic_tma": ([40, 4, 257], [1.0, 1.0, 1.0], [P, P, P], C3, 0.0, np.float64),
    "nz_1024_tma_cluster": ([32, 4, 1024], [1.0, 1.0, 2.0], [P, P, N], C3, 2.0, np.float64),
    "fp32_nz_512_tma": ([32, 4, 512], [1.0, 1.0, 2.0], [P, P, N], C3, 2.0, np.float32),
    # deduplicated pivot cache (periodic x with whole tile pairs and / or periodic y): mirrored tiles, y pairs, the
    # periodic-z closure arrays (p2, den) under deduplication, 8-column tiles, the CTA-pair kernel, FP32 (32-column tiles)
    "dedup_xy_64x32x48": ([64, 32, 48], [6.0, 3.0, 2.0], [P, P, N], C3, 2.0, np.float64),
    "dedup_x_128x24x40": ([128, 24, 40], [12.0, 2.0, 2.0], [P, N, N], C3, 1.5, np.float64),
    "dedup_y_40x32x36": ([40, 32, 36], [1.0, 1.0, 1.0], [N, P, D], C3, 1.0, np.float64),
    "dedup_xyz_periodic_64x16x65": ([64, 16, 65], [6.2832] * 3, [P, P, P], C3, 0.0, np.float64),
    "dedup_tma_96x8x256": ([96, 8, 256], [6.0, 3.0, 2.0], [P, P, N], C3, 2.0, np.float64),
    "dedup_periodic_tma_64x6x257": ([64, 6, 257], [1.0, 1.0, 1.0], [P, P, P], C3, 0.0, np.float64),
    "dedup_cluster_64x8x1024": ([64, 8, 1024], [6.0, 3.0, 2.0], [P, P, N], C3, 2.0, np.float64),
    "dedup_8col_periodic_64x4x700": ([64, 4, 700], [1.0, 1.0, 1.0], [P, P, P], C3, 0.0, np.float64),
    "fp32_dedup_128x16x256": ([128, 16, 256], [6.0, 3.0, 2.0], [P, P, N], C3, 2.0, np.float32),
    # implicit-diffusion (Helmholtz) operators: face-centred in one direction, Dirichlet walls
    "helm_u_face_x": ([16, 12, 20], [1.0, 1.0, 1.0], [D, D, D], ["f", "c", "c"], 1.0, np.float64),
    "helm_v_face_y": ([16, 12, 20], [1.0, 1.0, 1.0], [P, D, D], ["c", "f", "c"], 1.0, np.float64),
    "helm_w_face_z": ([16, 12, 20], [1.0, 1.0, 1.0], [P, P, D], ["c", "c", "f"], 1.0, np.float64),
    "helm_w_face_z_nn": ([16, 12, 20], [1.0, 1.0, 1.0], [N, N, N], ["f", "f", "f"], 1.0, np.float64),
    # single precision build (-D_SINGLE_PRECISION)
    "fp32_channel": ([32, 16, 40], [12.0, 6.0, 2.0], [P, P, N], C3, 2.0, np.float32),
    "fp32_ldc": ([2, 64, 64], [0.03125, 1.0, 1.0], [P, N, N], C3, 0.0, np.float32),
}
# cases of the distributed (z-slab) solve: name -> (ng, l, cbc, c_or_f, gr, dtype, helmholtz)
DIST_CASES = {
    "chan_64x64x64": ([64, 64, 64], [6.0, 3.0, 2.0], [P, P, N], C3, 2.0, np.float64, False),
    "tgv_64x128x64": ([64, 128, 64], [6.2832] * 3, [P, P, P], C3, 0.0, np.float64, False),
    "duct_128x64x96": ([128, 64, 96], [6.0, 2.0, 2.0], [P, N, N], C3, 1.5, np.float64, False),
    "uneven_64x64x70": ([64, 64, 70], [1.0, 1.0, 1.0], [N, P, ["D", "D"]], C3, 1.0, np.float64, False),
    "helm_w_64x64x64": ([64, 64, 64], [1.0, 1.0, 1.0], [P, P, ["D", "D"]], ["c", "c", "f"], 1.0, np.float64, True),
    "tma_32x64x256": ([32, 64, 256], [6.0, 3.0, 2.0], [P, P, N], C3, 2.0, np.float64, False),   # full chunks: TMA tile path
    "cluster_32x64x1024": ([32, 64, 1024], [6.0, 3.0, 2.0], [P, P, N], C3, 2.0, np.float64, False),   # CTA-pair Thomas + peer stores
    "fp32_64x64x64": ([64, 64, 64], [6.0, 3.0, 2.0], [P, P, N], C3, 2.0, np.float32, False),
    # y lengths / kinds the two-for-one kernels do not serve (odd, prime, small, REDFT00 / RODFT11): generic engine + row copies
    "geny_64x30x48": ([64, 30, 48], [6.0, 3.0, 2.0], [P, P, N], C3, 2.0, np.float64, False),
    "geny_40x17x36": ([40, 17, 36], [1.0, 1.0, 1.0], [N, P, ["D", "D"]], C3, 1.0, np.float64, False),
    "geny_face_48x33x40": ([48, 33, 40], [1.0, 1.0, 1.0], [P, N, ["D", "D"]], ["c", "f", "c"], 1.5, np.float64, False),
    "geny_mixed_32x20x24": ([32, 20, 24], [1.0, 1.0, 1.0], [P, ["D", "N"], ["D", "D"]], C3, 0.5, np.float64, False),
    # chunk lengths that are not powers of two (m = 3 and m = 5 rows per thread of the pipelined tridiagonal kernel)
    "m3_32x64x160": ([32, 64, 160], [6.0, 3.0, 2.0], [P, P, N], C3, 2.0, np.float64, False),
    "m5_32x64x300": ([32, 64, 300], [1.0, 1.0, 2.0], [P, P, D], C3, 1.0, np.float64, False),
    # regular operator with 256-row slabs on 2 ranks: full chunks of the pipelined kernel (TMA tiles) in distributed-TDMA mode
    "reg_32x64x512": ([32, 64, 512], [1.0, 1.0, 2.0], [P, P, D], C3, 1.0, np.float64, False),
}

HELMHOLTZ = {"helm_u_face_x", "helm_v_face_y", "helm_w_face_z", "helm_w_face_z_nn"}
ALPHA = -0.0123  # alpha = -0.5 * visc * dt_rk < 0 (src/main.f90:456)


def build_case(name):
    ng, l, cbc, cf, gr, dt = CASES[name]
    cs = O.make_case(ng, l, cbc, c_or_f=cf, gr=gr, dtype=dt)
    return cs


def make_rhs(cs, seed=123):
    """Seeded RHS on the haloed grid; made compatible when the operator is singular."""
    ng, dt = cs["ng"], cs["dtype"]
    rhs = O.hash_field(ng, seed, np.float64)
    singular = all(b[0] in "PN" and b[1] in "PN" for b in cs["cbc"])
    if singular:
        w = cs["dzf"][1:-1].astype(np.float64)[:, None, None]
        rhs = rhs - (rhs * w).sum() / (w.sum() * ng[0] * ng[1])
    p = np.zeros((ng[2] + 2, ng[1] + 2, ng[0] + 2), dtype=dt)
    p[1:-1, 1:-1, 1:-1] = rhs.astype(dt)
    return p


def oracle_solve(name, cs, p, helmholtz=False):
    ng = cs["ng"]
    ref = p.copy()
    if helmholtz:
        O.solve_helmholtz(ng, ng, cs["arrplan"], cs["normfft"], ALPHA, cs["lambdaxy"], cs["a"], cs["b"], cs["c"],
                          None, None, None, cs["cbc"], cs["c_or_f"], ref)
    else:
        O.solver(ng, ng, cs["arrplan"], cs["normfft"], cs["lambdaxy"], cs["a"], cs["b"], cs["c"], cs["cbc"], cs["c_or_f"], ref)
    return ref


def is_singular(cs):
    return all(b[0] in "PN" and b[1] in "PN" for b in cs["cbc"])


def null_mode_is_pinned(cs, helmholtz=False):
    """Does the reference's singular-pivot test (src/solver.f90:151-164 / :276-283) fire for the
    lambda = 0 column?  On uniform grids the last pivot of that column is exactly 0 (or within
    eps*max(...)) and the reference pins p = 0 there.  On stretched grids rounding leaves a pivot a
    few ulp ABOVE the tolerance, the test does not fire, and the reference divides the O(eps)
    compatibility residual of the right-hand side by an O(eps) pivot: the constant (null-space)
    component of ITS OWN answer is then rounding noise of the FFT library's summation order.  Parity
    is defined modulo that constant in this case (pressure only enters the code through its
    gradient: src/correc.f90:33-59)."""
    if helmholtz or not is_singular(cs):
        return True
    a, b, c = cs["a"], cs["b"], cs["c"]
    dt = a.dtype.type
    eps = np.finfo(a.dtype).eps
    lam = cs["lambdaxy"].flat[0]
    per = cs["cbc"][2] == P
    nn = len(a) - 1 if per else len(a)
    with np.errstate(all="ignore"):
        d = c[0] * (dt(1) / (b[0] + lam))
        for k in range(1, nn):
            bl, ad = b[k] + lam, a[k] * d
            den = bl - ad
            if k == nn - 1:
                return bool(abs(den) <= eps * max(abs(bl), abs(ad)))
            d = c[k] * (dt(1) / den)
    return True


def rel_l2_mod_const(got, ref):
    """relative L2 of (got - ref) after projecting out the constant mode"""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    diff = got - ref
    diff = diff - diff.mean()
    return float(np.linalg.norm(diff) / max(np.linalg.norm(ref - ref.mean()), 1e-300))


def parity_error(cs, got, ref, helmholtz=False):
    """the parity metric of the full solve: strict relative L2 whenever the reference's answer is
    well defined, modulo the constant mode when it is not (see null_mode_is_pinned)."""
    if null_mode_is_pinned(cs, helmholtz):
        return rel_l2(got, ref)
    return rel_l2_mod_const(got, ref)


def rel_l2(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))
