"""CPU model of how the 2D half-sweep kernel keeps the periodic y-halo rows itself (csrc/pf_sor.cu, YIMG 1 / 2),
against the reference's order of operations (ibm_2d_uniform_omp_cpu.f90:318-392: refresh ALL halo rows and copy
p_old before EACH half-sweep).  Random coefficients, even and odd n, bit for bit."""
import numpy as np
import pytest


def _coeffs(rng, n, m):
    c = {k: rng.uniform(0.5, 1.5, (n + 2, m + 2)) for k in ("ae", "aw", "an", "as", "bb")}
    c["ap"] = -(c["ae"] + c["aw"] + c["an"] + c["as"]) - rng.uniform(0.0, 0.1, (n + 2, m + 2))
    return c


def _update(c, p, om):
    I = (slice(1, -1), slice(1, -1))
    return ((c["bb"][I] - c["ae"][I] * p[1:-1, 2:] - c["aw"][I] * p[1:-1, :-2] - c["an"][I] * p[2:, 1:-1]
             - c["as"][I] * p[:-2, 1:-1]) / c["ap"][I] * om + p[I] * (1.0 - om))


def _masks(n, m):
    jj, ii = np.meshgrid(np.arange(1, n + 1), np.arange(1, m + 1), indexing="ij")
    return {1: ((ii + jj) % 2) == 1, 0: ((ii + jj) % 2) == 0}


def reference_scheme(c, p, iters, om):
    n, m = p.shape[0] - 2, p.shape[1] - 2
    mk = _masks(n, m)
    for _ in range(iters):
        for colour in (1, 0):                       # first colour (i+j) odd (:339-344)
            p[0, 1:m + 1] = p[n, 1:m + 1]           # :323-330
            p[n + 1, 1:m + 1] = p[1, 1:m + 1]
            po = p.copy()                           # :332-337
            new = _update(c, po, om)
            p[1:-1, 1:-1][mk[colour]] = new[mk[colour]]
    p[0, 1:m + 1] = p[n, 1:m + 1]
    p[n + 1, 1:m + 1] = p[1, 1:m + 1]
    return p


def kernel_scheme(c, p, iters, om):
    """one refresh up front; then every half-sweep of colour X updates its cells IN PLACE from the stored halos and
    refreshes the halo cells of colour X: from the cells it just wrote (n even: the image has the same colour) or from
    the other colour's rows 1 and n, which this launch does not touch (n odd)"""
    n, m = p.shape[0] - 2, p.shape[1] - 2
    mk = _masks(n, m)
    i = np.arange(1, m + 1)
    p[0, 1:m + 1] = p[n, 1:m + 1]
    p[n + 1, 1:m + 1] = p[1, 1:m + 1]
    for _ in range(iters):
        for colour in (1, 0):
            if n % 2 == 1:                          # YIMG 2: may run anywhere inside the launch -- do it FIRST here
                lo = i[((i + 0) % 2) == colour]     # halo cells (i, 0) / (i, n+1) of this colour
                hi = i[((i + n + 1) % 2) == colour]
                p[0, lo] = p[n, lo]
                p[n + 1, hi] = p[1, hi]
            new = _update(c, p, om)                 # in place: neighbours all have the other colour
            p[1:-1, 1:-1][mk[colour]] = new[mk[colour]]
            if n % 2 == 0:                          # YIMG 1: image stored with the cell
                lo = i[((i + n) % 2) == colour]
                hi = i[((i + 1) % 2) == colour]
                p[0, lo] = p[n, lo]
                p[n + 1, hi] = p[1, hi]
    p[0, 1:m + 1] = p[n, 1:m + 1]
    p[n + 1, 1:m + 1] = p[1, 1:m + 1]
    return p


@pytest.mark.parametrize("n,m", [(6, 7), (5, 8), (7, 7), (2, 3), (3, 2), (12, 9), (11, 10)])
def test_self_kept_halo_rows_equal_refresh_before_every_half_sweep(n, m):
    rng = np.random.default_rng(10 * n + m)
    c = _coeffs(rng, n, m)
    p0 = rng.standard_normal((n + 2, m + 2))
    a = reference_scheme(c, p0.copy(), 6, 1.7)
    b = kernel_scheme(c, p0.copy(), 6, 1.7)
    assert np.array_equal(a, b)


def test_the_odd_case_really_needs_the_stored_halo():
    """sanity of the model: with odd n, reading the live image instead of the stored halo changes the result"""
    n, m = 5, 6
    rng = np.random.default_rng(1)
    c = _coeffs(rng, n, m)
    p0 = rng.standard_normal((n + 2, m + 2))
    a = reference_scheme(c, p0.copy(), 3, 1.7)
    p = p0.copy()
    mk = _masks(n, m)
    for _ in range(3):
        for colour in (1, 0):
            for j, i in zip(*np.nonzero(mk[colour])):     # sequential in-place sweep with a live periodic wrap
                j1, i1 = j + 1, i + 1
                pn = p[j1 + 1, i1] if j1 < n else p[1, i1]
                ps = p[j1 - 1, i1] if j1 > 1 else p[n, i1]
                r = c["bb"][j1, i1] - c["ae"][j1, i1] * p[j1, i1 + 1] - c["aw"][j1, i1] * p[j1, i1 - 1] \
                    - c["an"][j1, i1] * pn - c["as"][j1, i1] * ps
                p[j1, i1] = r / c["ap"][j1, i1] * 1.7 + p[j1, i1] * (1.0 - 1.7)
    assert not np.array_equal(a[1:-1, 1:-1], p[1:-1, 1:-1])


# ------------------------------------------------------------------------------------------------------------------
# SOR variant 8 (csrc/pf_sor_tb2d.cu): temporally blocked tiles.  A tile is loaded with a ring 2T cells deep -- rows
# wrapped into 1..n, columns clipped at the x-halo columns, which are constants -- runs 2T half-sweeps on its own
# (compute everything from the tile, THEN store: the reference's p_old copy in miniature), never updates its
# outermost ring, and hands back only its owned cells.  Launches ping-pong between two arrays.
def tiled_scheme(c, p, iters, om, T, ow, oh):
    n, m = p.shape[0] - 2, p.shape[1] - 2
    cur = p.copy()
    left = iters
    while left > 0:
        t = min(T, left)
        D = 2 * t
        out = cur.copy()                    # (the kernel writes every owned cell and the x-halo columns next to them)
        for j0 in range(1, n + 1, oh):
            for i0 in range(1, m + 1, ow):
                i1, j1 = min(i0 + ow - 1, m), min(j0 + oh - 1, n)
                xlo, xhi = max(i0 - D, 0), min(i1 + D, m + 1)
                rows = [((j0 - D + r - 1) % n) + 1 for r in range((j1 - j0 + 1) + 2 * D)]     # real row of tile row r
                cols = np.arange(xlo, xhi + 1)
                tile = {k: a[np.ix_(rows, cols)].copy() for k, a in c.items()}
                P = cur[np.ix_(rows, cols)].copy()
                ii, rr = np.meshgrid(cols, np.array(rows), indexing="xy")
                cell = (ii >= 1) & (ii <= m)
                inner = np.zeros_like(cell)
                inner[1:-1, 1:-1] = True    # the outermost ring is never updated
                for hs in range(2 * t):
                    colour = (hs & 1) ^ 1   # (i+j) odd first
                    new = P.copy()
                    I = (slice(1, -1), slice(1, -1))
                    new[I] = ((tile["bb"][I] - tile["ae"][I] * P[1:-1, 2:] - tile["aw"][I] * P[1:-1, :-2]
                               - tile["an"][I] * P[2:, 1:-1] - tile["as"][I] * P[:-2, 1:-1]) / tile["ap"][I] * om
                              + P[I] * (1.0 - om))
                    upd = cell & inner & (((ii + rr) % 2) == colour)
                    P = np.where(upd, new, P)           # two phases: all reads of the half-sweep, then all stores
                own_r = slice(D, D + (j1 - j0 + 1))
                sx0, sx1 = (0 if i0 == 1 else i0), (m + 1 if i1 == m else i1)
                out[j0:j1 + 1, sx0:sx1 + 1] = P[own_r, sx0 - xlo:sx1 - xlo + 1]
        cur = out
        left -= t
    cur[0, 1:m + 1] = cur[n, 1:m + 1]       # the closing halo refresh (:588-605 analogue, sor_refresh in pf_api.cu)
    cur[n + 1, 1:m + 1] = cur[1, 1:m + 1]
    return cur


@pytest.mark.parametrize("n,m,T,ow,oh", [(12, 14, 2, 5, 4), (11, 9, 2, 4, 3), (7, 7, 3, 7, 7), (2, 5, 4, 3, 2), (3, 4, 4, 2, 3),
                                         (16, 20, 4, 8, 8), (9, 33, 4, 16, 4)])
def test_temporally_blocked_tiles_equal_the_reference_iterations(n, m, T, ow, oh):
    """incl. odd n (same-colour neighbours across the periodic seam), tiles that wrap the period several times
    (n = 2, 3 with a ring of 8), iteration counts that are not a multiple of T"""
    rng = np.random.default_rng(100 * n + m)
    c = _coeffs(rng, n, m)
    p0 = rng.standard_normal((n + 2, m + 2))
    for iters in (1, T, 2 * T + 1):
        a = reference_scheme(c, p0.copy(), iters, 1.7)
        b = tiled_scheme(c, p0.copy(), iters, 1.7, T, ow, oh)
        assert np.array_equal(a[1:-1], b[1:-1]), (iters, int((a != b).sum()))
        assert np.array_equal(a[[0, n + 1], 1:m + 1], b[[0, n + 1], 1:m + 1])


def test_a_ring_one_cell_too_thin_is_wrong():
    """sanity of the model: 2T half-sweeps need a ring 2T deep"""
    n, m, T = 16, 20, 2
    rng = np.random.default_rng(3)
    c = _coeffs(rng, n, m)
    p0 = rng.standard_normal((n + 2, m + 2))
    a = reference_scheme(c, p0.copy(), T, 1.7)
    b = tiled_scheme(c, p0.copy(), T, 1.7, T, 8, 8)
    assert np.array_equal(a[1:-1, 1:-1], b[1:-1, 1:-1])
    # ... the same launch with a ring of 2T - 1 cells reaches the owned cells with stale values
    assert np.array_equal(a[1:-1, 1:-1], _tiled_with_ring(c, p0.copy(), T, 1.7, 2 * T, 8, 8)[1:-1, 1:-1])
    assert not np.array_equal(a[1:-1, 1:-1], _tiled_with_ring(c, p0.copy(), T, 1.7, 2 * T - 1, 8, 8)[1:-1, 1:-1])


def _tiled_with_ring(c, p, t, om, D, ow, oh):
    """one launch of t iterations on tiles with a ring of D cells (D < 2t is too thin)"""
    n, m = p.shape[0] - 2, p.shape[1] - 2
    cur, out = p.copy(), p.copy()
    for j0 in range(1, n + 1, oh):
        for i0 in range(1, m + 1, ow):
            i1, j1 = min(i0 + ow - 1, m), min(j0 + oh - 1, n)
            xlo, xhi = max(i0 - D, 0), min(i1 + D, m + 1)
            rows = [((j0 - D + r - 1) % n) + 1 for r in range((j1 - j0 + 1) + 2 * D)]
            cols = np.arange(xlo, xhi + 1)
            tile = {k: a[np.ix_(rows, cols)].copy() for k, a in c.items()}
            P = cur[np.ix_(rows, cols)].copy()
            ii, rr = np.meshgrid(cols, np.array(rows), indexing="xy")
            inner = np.zeros(P.shape, bool)
            inner[1:-1, 1:-1] = True
            for hs in range(2 * t):
                colour = (hs & 1) ^ 1
                new = P.copy()
                I = (slice(1, -1), slice(1, -1))
                new[I] = ((tile["bb"][I] - tile["ae"][I] * P[1:-1, 2:] - tile["aw"][I] * P[1:-1, :-2]
                           - tile["an"][I] * P[2:, 1:-1] - tile["as"][I] * P[:-2, 1:-1]) / tile["ap"][I] * om
                          + P[I] * (1.0 - om))
                P = np.where((ii >= 1) & (ii <= m) & inner & (((ii + rr) % 2) == colour), new, P)
            out[j0:j1 + 1, i0:i1 + 1] = P[D:D + (j1 - j0 + 1), i0 - xlo:i1 - xlo + 1]
    return out
