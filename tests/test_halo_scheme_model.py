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
