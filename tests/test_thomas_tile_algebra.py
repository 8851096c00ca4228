"""CPU test of the tile bookkeeping of the pipelined tridiagonal kernel (cans_b200/csrc/thomas_kernels.cuh), restated in Python:

  * `decode` (tile number -> first y row, x tile) is a bijection onto the tiles of the launch for every combination of the
    pivot-cache deduplication flags (dx, dy) and of the tall-tile factor jb, and consecutive tile numbers share their stored
    pivot tile (what makes neighbouring CTAs of the persistent grid hit L2 instead of HBM);
  * `zux` / `zuy` (column / row -> stored column / row of the deduplicated cache) stay inside the stored extents, map whole
    128-byte tiles onto whole tiles, and pair exactly the columns whose eigenvalues agree (split order in x, halfcomplex in y);
  * the seam rule of tall tiles (a_0 := 0, c_{nn-1} := 0 per system) reproduces jb independent Thomas solves.
No GPU is involved: this guards the index algebra the GPU parity tests exercise on a handful of shapes only."""
import numpy as np
import pytest

COLS = 16


def decode(t, tiles_x, ny, dx, dy, jb, grouped):
    if not grouped:
        tj, ti = divmod(t, tiles_x)
        return tj * jb, ti
    mx = my = 0
    if dx:
        mx, t = t & 1, t >> 1
    if dy:
        my, t = t & 1, t >> 1
    ntu = tiles_x // 2 if dx else tiles_x
    qy, tu = divmod(t, ntu)
    ti = tu + tiles_x // 2 if mx else tu
    if not dy:
        tj = qy * jb
    else:
        tj = (ny // 2 if my else 0) if qy == 0 else (ny - qy if my else qy)
    return tj, ti


def zux(i, nx, nxu, dx):
    return i - nx // 2 if (dx and i >= nxu) else i


def zuy(j, ny, dy):
    return ny - j if (dy and 2 * j > ny) else j


@pytest.mark.parametrize("nx,ny", [(64, 8), (96, 6), (128, 32), (1024, 512), (32, 16)])
@pytest.mark.parametrize("dx,dy,jb", [(0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1), (0, 0, 4), (1, 0, 2)])
def test_decode_is_a_bijection_and_groups_share_pivots(nx, ny, dx, dy, jb):
    tiles_x = nx // COLS
    if dx and (nx % (2 * COLS) or nx < 4 * COLS):
        pytest.skip("the host only deduplicates x for whole tile pairs")
    if ny % jb or (dy and jb > 1):
        pytest.skip("combination the host never builds")
    nxu = nx // 2 + COLS if dx else nx
    grouped = bool(dx or dy) and tiles_x % 2 == 0 and (not dy or ny % 2 == 0)
    ntiles = tiles_x * (ny // jb)
    seen = set()
    stored = []
    for t in range(ntiles):
        tj, ti = decode(t, tiles_x, ny, dx, dy, jb, grouped)
        assert 0 <= ti < tiles_x and 0 <= tj < ny and tj % jb == 0
        seen.add((tj, ti))
        x0 = ti * COLS
        zx0, ju = zux(x0, nx, nxu, dx), zuy(tj, ny, dy)
        assert zx0 % COLS == 0 and 0 <= zx0 and zx0 + COLS <= nxu          # whole stored tiles, inside the cache
        assert 0 <= ju <= (ny // 2 if dy else ny - 1)
        assert all(zux(x0 + c, nx, nxu, dx) == zx0 + c for c in range(COLS))   # same column order: no reversal in the kernel
        stored.append((ju, zx0))
    assert len(seen) == ntiles
    if grouped:   # the members of a group (consecutive tile numbers) read the same stored pivot tile
        g = (2 if dx else 1) * (2 if dy else 1)
        shared = sum(len(set(stored[k:k + g])) == 1 for k in range(0, ntiles, g))
        # all groups share, except those holding the self-paired modes (x tile 0 / nx/2, y rows 0 / ny/2)
        special = (ny // jb if dx else 0) + (tiles_x if dy else 0)
        assert shared >= ntiles // g - special


@pytest.mark.parametrize("n", [64, 96, 1024])
def test_stored_columns_pair_equal_eigenvalues(n):
    """split order in x: position p < n/2 + 16 is stored, position p >= n/2 + 16 uses p - n/2; the two must carry the same mode"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from cans_b200.solver import eigenvalues
    hc = eigenvalues(n, "PP", "c")
    split_to_hc = lambda pos: pos if 2 * pos <= n else n - (pos - n // 2)
    lam_split = np.array([hc[split_to_hc(p)] for p in range(n)])
    nxu = n // 2 + COLS
    for p in range(n):
        u = zux(p, n, nxu, 1)
        assert abs(lam_split[p] - lam_split[u]) <= 1e-10 * max(abs(lam_split[p]), 1e-300) or lam_split[p] == lam_split[u]
    for j in range(n):   # halfcomplex rows: j > n/2 uses n - j
        u = zuy(j, n, 1)
        assert abs(hc[j] - hc[u]) <= 1e-10 * max(abs(hc[j]), 1e-300) or hc[j] == hc[u]


@pytest.mark.parametrize("nn,jb", [(5, 3), (16, 4), (40, 8), (128, 4)])
def test_tall_tile_seams_decouple_the_systems(nn, jb):
    """one Thomas solve over jb * nn rows with a'[k=0] = 0 and c'[k=nn-1] = 0 equals jb separate solves"""
    rng = np.random.default_rng(nn * 7 + jb)
    a, c = rng.uniform(0.5, 1.5, nn), rng.uniform(0.5, 1.5, nn)
    b = -(a + c) - rng.uniform(0.1, 1.0, nn)
    rhs = rng.uniform(-1, 1, jb * nn)

    def thomas(a, b, c, r):
        n = len(r)
        d, p = np.zeros(n), np.zeros(n)
        z = 1.0 / b[0]
        d[0], p[0] = c[0] * z, r[0] * z
        for k in range(1, n):
            z = 1.0 / (b[k] - a[k] * d[k - 1])
            d[k] = c[k] * z
            p[k] = (r[k] - a[k] * p[k - 1]) * z
        for k in range(n - 2, -1, -1):
            p[k] -= d[k] * p[k + 1]
        return p

    ref = np.concatenate([thomas(a, b, c, rhs[s * nn:(s + 1) * nn]) for s in range(jb)])
    at, bt, ct = np.tile(a, jb), np.tile(b, jb), np.tile(c, jb)
    kk = np.arange(jb * nn) % nn
    at[kk == 0] = 0.0
    ct[kk == nn - 1] = 0.0
    got = thomas(at, bt, ct, rhs)
    assert np.allclose(got, ref, rtol=1e-13, atol=1e-14)
