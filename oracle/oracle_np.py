"""Independent numpy restatement of the PixelFlow hot path (TEST INFRASTRUCTURE ONLY).

Second, separately written transcription of the reference (src/omp_parallel/*.f90), in
vectorised-slice form.  Its only purpose is to be compared bit for bit with the C oracle
(oracle/pf_oracle.c): two independent restatements agreeing exactly was the first substitute
for running the Fortran, which cannot be compiled in this environment.
Since then the C oracle is also pinned against the machine-translated reference itself
(oracle/f90toc.py, tests/test_ref_translation.py) -- see the header of pf_oracle.c.

Nothing in the product path may import this module (tests/ only).

Arrays are numpy float64 of shape (l+2, n+2, m+2) [2D: (n+2, m+2)], index order [k, j, i],
i.e. the Fortran A(i,j,k) is a[k, j, i].  numpy evaluates elementwise IEEE-754 double
operations without contraction or reassociation, and Python's `*` and `/` associate left to
right like Fortran's, so an expression typed in the reference's order evaluates identically.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SMALL = 1.0e-6   # ibm_3d_uniform_omp_cpu.f90:169
ALPHA = 32.0     # ibm_3d_uniform_omp_cpu.f90:170
PI = math.atan(1.0) * 4.0


@dataclass
class Params:
    m: int
    n: int
    l: int = 1
    dx: float = 1.0
    dy: float = 1.0
    dz: float = 1.0
    dt: float = 1.0
    xnue: float = 0.0
    xlambda: float = 0.0
    density: float = 1.0
    thickness: float = 1.5
    nonslip: bool = True
    iter_max: int = 100
    relux_factor: float = 1.7
    inlet_velocity: float = 1.0
    outlet_pressure: float = 0.0
    AoA: float = 0.0
    # top, bottom, east, west, south, north  (ibm_3d_air_condition_omp_cpu.f90:10-15)
    wall: tuple = (1, 0, 0, 0, 2, 0)


# ------------------------------------------------------------------------------------------
# 3D
# ------------------------------------------------------------------------------------------
def _sh(a, P, di=0, dj=0, dk=0):
    """interior block shifted by (di,dj,dk):  A(i+di, j+dj, k+dk) for i=1..m, j=1..n, k=1..l"""
    return a[1 + dk:P.l + 1 + dk, 1 + dj:P.n + 1 + dj, 1 + di:P.m + 1 + di]


def porosity_halo_3d_uniform(P, e):
    """lib/grid.f90:349-378"""
    m, n, l = P.m, P.n, P.l
    e[1:l + 2, 1:n + 2, 0] = e[1:l + 2, 1:n + 2, 1]
    e[1:l + 2, 1:n + 2, m + 1] = e[1:l + 2, 1:n + 2, m]
    e[:, 0, :] = e[:, n, :]
    e[:, n + 1, :] = e[:, 1, :]
    e[0, :, :] = e[l, :, :]
    e[l + 1, :, :] = e[1, :, :]


def porosity_halo_3d_wall(P, e):
    """lib/grid.f90:215-243"""
    m, n, l = P.m, P.n, P.l
    e[:, :, 0] = e[:, :, 1]
    e[:, :, m + 1] = e[:, :, m]
    e[:, 0, :] = e[:, 1, :]
    e[:, n + 1, :] = e[:, n, :]
    e[0, :, :] = e[1, :, :]
    e[l + 1, :, :] = e[l, :, :]


def divergence_3d(P, air, uo, vo, wo, div):
    """ibm_3d_uniform_omp_cpu.f90:185-222 ; air :198-237"""
    m, n, l = P.m, P.n, P.l
    _sh(div, P)[...] = ((_sh(uo, P, 1) - _sh(uo, P, -1)) / P.dx * 0.5
                        + (_sh(vo, P, 0, 1) - _sh(vo, P, 0, -1)) / P.dy * 0.5
                        + (_sh(wo, P, 0, 0, 1) - _sh(wo, P, 0, 0, -1)) / P.dz * 0.5)
    div[1:l + 1, 1:n + 1, 0] = 0.0
    div[1:l + 1, 1:n + 1, m + 1] = 0.0
    if air:
        div[0, 1:n + 1, 1:m + 1] = 0.0
        div[l + 1, 1:n + 1, 1:m + 1] = 0.0
        div[1:l + 1, 0, 1:m + 1] = 0.0
        div[1:l + 1, n + 1, 1:m + 1] = 0.0
    else:
        div[1:l + 1, 0, 1:m + 1] = div[1:l + 1, n, 1:m + 1]
        div[1:l + 1, n + 1, 1:m + 1] = div[1:l + 1, 1, 1:m + 1]
        div[0, 1:n + 1, 1:m + 1] = div[l, 1:n + 1, 1:m + 1]
        div[l + 1, 1:n + 1, 1:m + 1] = div[1, 1:n + 1, 1:m + 1]


def predictor_3d(P, uo, vo, wo, e, div, u, v, w):
    """ibm_3d_uniform_omp_cpu.f90:228-381"""
    dx, dy, dz, dt, nu, lam = P.dx, P.dy, P.dz, P.dt, P.xnue, P.xlambda
    s = lambda a, di=0, dj=0, dk=0: _sh(a, P, di, dj, dk)
    U, V, W, E, D = s(uo), s(vo), s(wo), s(e), s(div)
    ex = (s(e, 1) - s(e, -1))
    ey = (s(e, 0, 1) - s(e, 0, -1))
    ez = (s(e, 0, 0, 1) - s(e, 0, 0, -1))

    def comp(q, qo, dq, axis):
        """q: output array, qo: old field of this component, dq: spacing of the wall term"""
        Q = s(qo)
        qx = s(qo, 1) - s(qo, -1)
        qy = s(qo, 0, 1) - s(qo, 0, -1)
        qz = s(qo, 0, 0, 1) - s(qo, 0, 0, -1)
        r = Q - dt * U * qx / dx * 0.5
        r = r - dt * V * qy / dy * 0.5
        r = r - dt * W * qz / dz * 0.5
        r = r + dt * nu * (s(qo, 1) - 2. * Q + s(qo, -1)) / dx / dx
        r = r + dt * nu * (s(qo, 0, 1) - 2. * Q + s(qo, 0, -1)) / dy / dy
        r = r + dt * nu * (s(qo, 0, 0, 1) - 2. * Q + s(qo, 0, 0, -1)) / dz / dz
        if axis == 0:
            r = r + dt * (nu + lam) * (s(div, 1) - s(div, -1)) / dx * 0.5
            t1 = (qx / dx * 0.5 + qx / dx * 0.5) * nu * ex / dx * 0.5
            t2 = (qy / dy * 0.5 + (s(vo, 1) - s(vo, -1)) / dx * 0.5) * nu * ey / dy * 0.5
            t3 = (qz / dz * 0.5 + (s(wo, 1) - s(wo, -1)) / dx * 0.5) * nu * ez / dz * 0.5
            t4 = D * ex / dx * 0.5 * lam
        elif axis == 1:
            r = r + dt * (nu + lam) * (s(div, 0, 1) - s(div, 0, -1)) / dy * 0.5
            t1 = (qx / dx * 0.5 + (s(uo, 0, 1) - s(uo, 0, -1)) / dy * 0.5) * nu * ex / dx * 0.5
            t2 = (qy / dy * .5 + qy / dy * 0.5) * nu * ey / dy * 0.5
            t3 = (qz / dz * .5 + (s(wo, 0, 1) - s(wo, 0, -1)) / dy * 0.5) * nu * ez / dz * 0.5
            t4 = D * ey / dy * 0.5 * lam
        else:
            r = r + dt * (nu + lam) * (s(div, 0, 0, 1) - s(div, 0, 0, -1)) / dz * 0.5
            t1 = (qx / dx * 0.5 + (s(uo, 0, 0, 1) - s(uo, 0, 0, -1)) / dz * 0.5) * nu * ex / dx * 0.5
            t2 = (qy / dy * 0.5 + (s(vo, 0, 0, 1) - s(vo, 0, 0, -1)) / dz * 0.5) * nu * ey / dy * 0.5
            t3 = (qz / dz * 0.5 + qz / dz * 0.5) * nu * ez / dz * 0.5
            t4 = D * ez / dz * 0.5 * lam
        r = r + dt * (t1 + t2 + t3 + t4) / E
        if P.nonslip:
            r = r - dt * nu * Q / ((P.thickness * dq) * (P.thickness * dq)) * ALPHA * E * (1. - E) * (1. - E)
        s(q)[...] = r

    comp(u, uo, dx, 0)
    comp(v, vo, dy, 1)
    comp(w, wo, dz, 2)


def matrix_3d(P, u, v, w, e, c):
    """ibm_3d_uniform_omp_cpu.f90:386-414 ; c is a dict of coefficient arrays"""
    dx, dy, dz, dt, rho = P.dx, P.dy, P.dz, P.dt, P.density
    s = lambda a, di=0, dj=0, dk=0: _sh(a, P, di, dj, dk)
    E = s(e)
    s(c['ae'])[...] = dt * np.maximum(SMALL, (s(e, 1) + E) * 0.5) / dx / dx
    s(c['aw'])[...] = dt * np.maximum(SMALL, (E + s(e, -1)) * 0.5) / dx / dx
    s(c['an'])[...] = dt * np.maximum(SMALL, (s(e, 0, 1) + E) * 0.5) / dy / dy
    s(c['as'])[...] = dt * np.maximum(SMALL, (E + s(e, 0, -1)) * 0.5) / dy / dy
    s(c['at'])[...] = dt * np.maximum(SMALL, (s(e, 0, 0, 1) + E) * 0.5) / dz / dz
    s(c['ab'])[...] = dt * np.maximum(SMALL, (E + s(e, 0, 0, -1)) * 0.5) / dz / dz
    s(c['ap'])[...] = -s(c['ae']) - s(c['aw']) - s(c['an']) - s(c['as']) - s(c['at']) - s(c['ab'])
    s(c['bb'])[...] = (((s(e, 1) * s(u) + E * s(u, 1)) * 0.5 - (s(e, -1) * s(u) + E * s(u, -1)) * 0.5) * rho / dx
                       + ((s(e, 0, 1) * s(v) + E * s(v, 0, 1)) * 0.5 - (s(e, 0, -1) * s(v) + E * s(v, 0, -1)) * 0.5) * rho / dy
                       + ((s(e, 0, 0, 1) * s(w) + E * s(w, 0, 0, 1)) * 0.5 - (s(e, 0, 0, -1) * s(w) + E * s(w, 0, 0, -1)) * 0.5) * rho / dz)


def boundary_matrix_3d_uniform(P, p, c):
    """ibm_3d_uniform_omp_cpu.f90:618-663"""
    m, n, l = P.m, P.n, P.l
    J, K = slice(1, n + 1), slice(1, l + 1)
    c['ae'][K, J, 1] = c['ae'][K, J, 1] + c['aw'][K, J, 1]
    c['aw'][K, J, 1] = 0.0
    c['bb'][K, J, m] = c['bb'][K, J, m] + c['ae'][K, J, m] * p[K, J, m + 1]
    for name in ('ae', 'aw', 'an', 'as', 'at', 'ab'):
        c[name][K, J, m] = 0.0


def boundary_matrix_3d_air(P, p, e, c, own_top=True, own_bottom=True, bb1=None):
    """ibm_3d_air_condition_omp_cpu.f90:665-867 (faces in order top,bottom,east,west,north,south).
    Slab models only (tests/test_slab_schedule_gloo.py): own_top / own_bottom = this slab holds that z face; bb1 = the
    raw right-hand side of GLOBAL plane 1 (what `bb(i,j,1)` of :702 is when plane 1 lives on another slab)."""
    m, n, l = P.m, P.n, P.l
    top, bottom, east, west, south, north = P.wall
    six = ('ae', 'aw', 'an', 'as', 'at', 'ab')

    def face(idx, code, grow, shrink, halo_idx, top_quirk=False):
        poro = e[idx]
        if code in (0, 1):
            wallmask = np.ones(poro.shape, dtype=bool)
        else:
            wallmask = poro < 0.9
        g, sname = c[grow][idx], c[shrink][idx]
        outmask = ~wallmask
        # outlet first uses the un-folded coefficient; the two masks are disjoint
        if outmask.any():
            bbv = c['bb'][idx]
            base = (c['bb'][1, :, :] if bb1 is None else bb1) if top_quirk else bbv      # bb(i,j,l)=bb(i,j,1)+... :702
            bbv[outmask] = (base + sname * p[halo_idx])[outmask]
            for nm in six:
                c[nm][idx][outmask] = 0.0
        g[wallmask] = (g + sname)[wallmask]
        sname[wallmask] = 0.0

    A = slice(None)
    if own_top:
        face((l, A, A), top, 'ab', 'at', (l + 1, A, A), top_quirk=True)
    if own_bottom:
        face((1, A, A), bottom, 'at', 'ab', (0, A, A))
    face((A, A, m), east, 'aw', 'ae', (A, A, m + 1))
    face((A, A, 1), west, 'ae', 'aw', (A, A, 0))
    face((A, n, A), north, 'as', 'an', (A, n + 1, A))
    face((A, 1, A), south, 'an', 'as', (A, 0, A))


def _colour_mask_3d(P, parity):
    k, j, i = np.meshgrid(np.arange(1, P.l + 1), np.arange(1, P.n + 1), np.arange(1, P.m + 1), indexing='ij')
    return ((i + j + k) % 2) == parity


def sor_3d(P, periodic, iters, p, c):
    """ibm_3d_uniform_omp_cpu.f90:433-614: colour 1 = (i+j+k) even, colour 2 = odd; error is the
    running max of |p - p_old| after the second half-sweep (p_old = snapshot before it)."""
    m, n, l, om = P.m, P.n, P.l, P.relux_factor
    s = lambda a, di=0, dj=0, dk=0: _sh(a, P, di, dj, dk)
    masks = (_colour_mask_3d(P, 0), _colour_mask_3d(P, 1))
    error = 0.0

    def halo():
        if periodic:
            p[1:l + 1, 0, 1:m + 1] = p[1:l + 1, n, 1:m + 1]
            p[1:l + 1, n + 1, 1:m + 1] = p[1:l + 1, 1, 1:m + 1]
            p[0, 1:n + 1, 1:m + 1] = p[l, 1:n + 1, 1:m + 1]
            p[l + 1, 1:n + 1, 1:m + 1] = p[1, 1:n + 1, 1:m + 1]

    for _ in range(iters):
        for half in (0, 1):
            halo()
            po = p.copy()
            new = ((s(c['bb']) - s(c['ae']) * s(po, 1) - s(c['aw']) * s(po, -1)
                    - s(c['an']) * s(po, 0, 1) - s(c['as']) * s(po, 0, -1)
                    - s(c['at']) * s(po, 0, 0, 1) - s(c['ab']) * s(po, 0, 0, -1))
                   / s(c['ap']) * om + s(po) * (1. - om))
            s(p)[masks[half]] = new[masks[half]]
        error = max(error, float(np.max(np.abs(s(p) - s(po)))))
    halo()
    return error


def project_3d(P, p, u, v, w):
    """ibm_3d_uniform_omp_cpu.f90:110-125"""
    s = lambda a, di=0, dj=0, dk=0: _sh(a, P, di, dj, dk)
    s(u)[...] = s(u) - P.dt / P.density * (s(p, 1) - s(p, -1)) / P.dx * 0.5
    s(v)[...] = s(v) - P.dt / P.density * (s(p, 0, 1) - s(p, 0, -1)) / P.dy * 0.5
    s(w)[...] = s(w) - P.dt / P.density * (s(p, 0, 0, 1) - s(p, 0, 0, -1)) / P.dz * 0.5


def boundary_3d_uniform(P, p, u, v, w):
    """ibm_3d_uniform_omp_cpu.f90:669-752"""
    m, n, l = P.m, P.n, P.l
    J, K = slice(1, n + 1), slice(1, l + 1)
    u[K, J, 1] = P.inlet_velocity * math.cos(P.AoA / 1300. * PI)
    v[K, J, 1] = P.inlet_velocity * math.sin(P.AoA / 1300. * PI)
    w[K, J, 1] = 0.0
    for a in (u, v, w):
        a[K, J, 0] = a[K, J, 1]
    p[K, J, 0] = p[K, J, 2]
    for a in (u, v, w):
        a[K, J, m + 1] = a[K, J, m - 1]
    p[K, J, m + 1] = P.outlet_pressure
    for a in (u, v, w, p):
        a[:, 0, :] = a[:, n, :]
        a[:, n + 1, :] = a[:, 1, :]
    for a in (u, v, w, p):
        a[0, :, :] = a[l, :, :]
        a[l + 1, :, :] = a[1, :, :]


def boundary_3d_air(P, e, p, u, v, w, own_top=True, own_bottom=True, e_top=None):
    """ibm_3d_air_condition_omp_cpu.f90:873-1170 (serial, order top,bottom,west,east,north,south).
    Slab models only: own_top / own_bottom as in boundary_matrix_3d_air; e_top = the porosity of GLOBAL plane l (what
    `porosity(i,j,l)` of :948 is when plane l lives on another slab)."""
    m, n, l = P.m, P.n, P.l
    top, bottom, east, west, south, north = P.wall
    uin, pout = P.inlet_velocity, P.outlet_pressure
    A = slice(None)

    def face(code, fluid, on, ghost, inner, inner_outlet, normal, inlet_vec, wall_mirror):
        """on: index of the boundary layer; ghost: ghost layer; inner: second layer (mirror source)
        inner_outlet: layer the outlet ghost copies; inlet_vec: (u,v,w) inlet values;
        wall_mirror: which component the wall ghost mirrors with a minus sign"""
        fields = (u, v, w)
        if code == 0:
            wallm = np.ones(fluid.shape, dtype=bool)
            inm = outm = ~wallm
        elif code == 1:
            inm, outm, wallm = fluid, np.zeros(fluid.shape, dtype=bool), ~fluid
        else:
            inm, outm, wallm = np.zeros(fluid.shape, dtype=bool), fluid, ~fluid
        # snapshot of the sources (all reads of one face point precede nothing else's writes:
        # every face point touches only its own column)
        src_inner = [f[inner].copy() for f in fields]
        src_out = [f[inner_outlet].copy() for f in fields]
        p_inner = p[inner].copy()
        for f, val in zip(fields, inlet_vec):
            f[on][inm] = val
            f[ghost][inm] = val
        p[ghost][inm] = p_inner[inm]
        for f, so in zip(fields, src_out):
            f[ghost][outm] = so[outm]
        p[ghost][outm] = pout
        for f in fields:
            f[on][wallm] = 0.0
        fields[wall_mirror][ghost][wallm] = -src_inner[wall_mirror][wallm]
        p[ghost][wallm] = p_inner[wallm]

    # top: inlet w=-uin; outlet ghosts copy l-1; wall mirrors w
    if own_top:
        face(top, e[l] >= 0.9, (l, A, A), (l + 1, A, A), (l - 1, A, A), (l - 1, A, A), 2, (0., 0., -uin), 2)
    # bottom: inlet tests porosity(i,j,l) (sic :948); outlet tests porosity(i,j,1) and copies k=1
    fl_b = ((e[l] if e_top is None else e_top) >= 0.9) if bottom == 1 else (e[1] >= 0.9)
    if own_bottom:
        face(bottom, fl_b, (1, A, A), (0, A, A), (2, A, A), (1, A, A), 2, (0., 0., uin), 2)
    # west: inlet u=uin; outlet copies i=1; wall mirrors u
    face(west, e[:, :, 1] >= 0.9, (A, A, 1), (A, A, 0), (A, A, 2), (A, A, 1), 0, (uin, 0., 0.), 0)
    # east: inlet u=-uin; outlet copies i=m; wall mirrors u
    face(east, e[:, :, m] >= 0.9, (A, A, m), (A, A, m + 1), (A, A, m - 1), (A, A, m), 0, (-uin, 0., 0.), 0)
    # north: inlet u=-uin (sic); outlet copies j=n; wall mirrors u (sic)
    face(north, e[:, n, :] >= 0.9, (A, n, A), (A, n + 1, A), (A, n - 1, A), (A, n, A), 1, (-uin, 0., 0.), 0)
    # south: inlet u=uin (sic); outlet copies j=2; wall mirrors v
    face(south, e[:, 1, :] >= 0.9, (A, 1, A), (A, 0, A), (A, 2, A), (A, 2, A), 1, (uin, 0., 0.), 1)


def initial_3d(P, air, p, u, v, w):
    """ibm_3d_uniform_omp_cpu.f90:756-793 ; air :1174-1211"""
    s = lambda a: _sh(a, P)
    s(u)[...] = 0.0 if air else P.inlet_velocity * math.cos(P.AoA / 360 * PI)
    s(v)[...] = 0.0 if air else P.inlet_velocity * math.sin(P.AoA / 360 * PI)
    s(w)[...] = 0.0
    s(p)[...] = P.outlet_pressure


@dataclass
class State3D:
    P: Params
    air: bool
    e: np.ndarray
    p: np.ndarray = None
    u: np.ndarray = None
    v: np.ndarray = None
    w: np.ndarray = None
    c: dict = field(default_factory=dict)

    def __post_init__(self):
        shape = self.e.shape
        for nm in ('p', 'u', 'v', 'w'):
            if getattr(self, nm) is None:
                setattr(self, nm, np.zeros(shape))
        for nm in ('ap', 'ae', 'aw', 'an', 'as', 'at', 'ab', 'bb', 'div'):
            self.c[nm] = np.zeros(shape)

    def step(self):
        P, c = self.P, self.c
        uo, vo, wo = self.u.copy(), self.v.copy(), self.w.copy()
        divergence_3d(P, self.air, uo, vo, wo, c['div'])
        predictor_3d(P, uo, vo, wo, self.e, c['div'], self.u, self.v, self.w)
        matrix_3d(P, self.u, self.v, self.w, self.e, c)
        if self.air:
            boundary_matrix_3d_air(P, self.p, self.e, c)
        else:
            boundary_matrix_3d_uniform(P, self.p, c)
        err = sor_3d(P, not self.air, P.iter_max, self.p, c)
        project_3d(P, self.p, self.u, self.v, self.w)
        if self.air:
            boundary_3d_air(P, self.e, self.p, self.u, self.v, self.w)
        else:
            boundary_3d_uniform(P, self.p, self.u, self.v, self.w)
        return err


# ------------------------------------------------------------------------------------------
# 2D
# ------------------------------------------------------------------------------------------
def _s2(a, P, di=0, dj=0):
    return a[1 + dj:P.n + 1 + dj, 1 + di:P.m + 1 + di]


def porosity_halo_2d(P, e):
    """lib/grid.f90:92-106"""
    m, n = P.m, P.n
    e[1:n + 2, 0] = e[1:n + 2, 1]
    e[1:n + 2, m + 1] = e[1:n + 2, m]
    e[0, :] = e[n, :]
    e[n + 1, :] = e[1, :]


def step_2d(P, backstep, e, p, u, v, c):
    """one time step of ibm_2d_uniform_omp_cpu.f90:80-125 (backstep deltas :533-534)"""
    m, n = P.m, P.n
    dx, dy, dt, nu, lam, rho, om = P.dx, P.dy, P.dt, P.xnue, P.xlambda, P.density, P.relux_factor
    s = lambda a, di=0, dj=0: _s2(a, P, di, dj)
    uo, vo = u.copy(), v.copy()
    div = c['div']
    # :172-194  (dx twice, sic)
    s(div)[...] = (s(uo, 1) - s(uo, -1)) / dx * .5 + (s(vo, 0, 1) - s(vo, 0, -1)) / dx * .5
    div[1:n + 1, 0] = 0.0
    div[1:n + 1, m + 1] = 0.0
    div[0, 1:m + 1] = div[n, 1:m + 1]
    div[n + 1, 1:m + 1] = div[1, 1:m + 1]
    U, V, E, D = s(uo), s(vo), s(e), s(div)
    ex, ey = s(e, 1) - s(e, -1), s(e, 0, 1) - s(e, 0, -1)
    wall = ((P.thickness * dx) * (P.thickness * dx))
    # u :200-228
    ux, uy = s(uo, 1) - s(uo, -1), s(uo, 0, 1) - s(uo, 0, -1)
    vx, vy = s(vo, 1) - s(vo, -1), s(vo, 0, 1) - s(vo, 0, -1)
    r = U - dt * (U * ux / dx / 2.)
    r = r - dt * (V * uy / dy / 2.)
    r = r + dt * nu * (s(uo, 1) - 2. * U + s(uo, -1)) / dx / dx
    r = r + dt * nu * (s(uo, 0, 1) - 2. * U + s(uo, 0, -1)) / dy / dy
    r = r + dt * (nu + lam) * (s(div, 1) - s(div, -1)) / dx * .5
    r = r + dt * ((ux / dx * .5 + ux / dx * .5) * nu * ex / dx * .5
                  + (uy / dy * .5 + vx / dx * .5) * nu * ey / dy * .5
                  + D * ex / dx * 0.5 * lam) / E
    if P.nonslip:
        r = r - dt * nu * U / wall * ALPHA * E * (1. - E) * (1. - E)
    s(u)[...] = r
    # v :232-258 (wall term with dx, sic)
    r = V - dt * (U * vx / dx / 2.)
    r = r - dt * (V * vy / dy / 2.)
    r = r + dt * nu * (s(vo, 1) - 2. * V + s(vo, -1)) / dx / dx
    r = r + dt * nu * (s(vo, 0, 1) - 2. * V + s(vo, 0, -1)) / dy / dy
    r = r + dt * (nu + lam) * (s(div, 0, 1) - s(div, 0, -1)) / dy * .5
    r = r + dt * ((vx / dx * .5 + uy / dy * .5) * nu * ex / dx * .5
                  + (vy / dy * .5 + vy / dy * .5) * nu * ey / dy * .5
                  + D * ey / dy * 0.5 * lam) / E
    if P.nonslip:
        r = r - dt * nu * V / wall * ALPHA * E * (1. - E) * (1. - E)
    s(v)[...] = r
    # matrix :262-278
    s(c['ae'])[...] = dt * np.maximum(SMALL, (s(e, 1) + E) * 0.5) / dx / dx
    s(c['aw'])[...] = dt * np.maximum(SMALL, (E + s(e, -1)) * 0.5) / dx / dx
    s(c['an'])[...] = dt * np.maximum(SMALL, (s(e, 0, 1) + E) * 0.5) / dy / dy
    s(c['as'])[...] = dt * np.maximum(SMALL, (E + s(e, 0, -1)) * 0.5) / dy / dy
    s(c['ap'])[...] = -s(c['ae']) - s(c['aw']) - s(c['an']) - s(c['as'])
    s(c['bb'])[...] = (((s(e, 1) * s(u) + E * s(u, 1)) * 0.5 - (s(e, -1) * s(u) + E * s(u, -1)) * 0.5) * rho / dx
                       + ((s(e, 0, 1) * s(v) + E * s(v, 0, 1)) * 0.5 - (s(e, 0, -1) * s(v) + E * s(v, 0, -1)) * 0.5) * rho / dy)
    # boundrary_matrix :410-454
    J = slice(1, n + 1)
    c['ae'][J, 1] = c['ae'][J, 1] + c['aw'][J, 1]
    c['aw'][J, 1] = 0.0
    c['bb'][J, m] = c['bb'][J, m] + c['ae'][J, m] * p[J, m + 1]
    for nm in ('ae', 'aw', 'an', 'as'):
        c[nm][J, m] = 0.0
    # SOR :293-406, first colour (i+j) odd, error in both half-sweeps
    jj, ii = np.meshgrid(np.arange(1, n + 1), np.arange(1, m + 1), indexing='ij')
    masks = (((ii + jj) % 2) == 1, ((ii + jj) % 2) == 0)
    error = 0.0
    for _ in range(P.iter_max):
        for half in (0, 1):
            p[0, 1:m + 1] = p[n, 1:m + 1]
            p[n + 1, 1:m + 1] = p[1, 1:m + 1]
            po = p.copy()
            new = ((s(c['bb']) - s(c['ae']) * s(po, 1) - s(c['aw']) * s(po, -1)
                    - s(c['an']) * s(po, 0, 1) - s(c['as']) * s(po, 0, -1)) / s(c['ap']) * om
                   + s(po) * (1. - om))
            mk = masks[half]
            s(p)[mk] = new[mk]
            if mk.any():
                error = max(error, float(np.max(np.abs(new[mk] - s(po)[mk]))))
    p[0, 1:m + 1] = p[n, 1:m + 1]
    p[n + 1, 1:m + 1] = p[1, 1:m + 1]
    # projection :103-115
    s(u)[...] = s(u) - dt / rho * (s(p, 1) - s(p, -1)) / dx * 0.5
    s(v)[...] = s(v) - dt / rho * (s(p, 0, 1) - s(p, 0, -1)) / dy * 0.5
    # boundary :460-538
    uin = P.inlet_velocity * math.cos(P.AoA / 180. * PI)
    vin = P.inlet_velocity * math.sin(P.AoA / 180. * PI)
    if backstep:
        u[J, 1] = uin * e[J, 1]
        v[J, 1] = vin * e[J, 1]
    else:
        u[J, 1] = uin
        v[J, 1] = vin
    u[J, 0] = u[J, 1]
    v[J, 0] = v[J, 1]
    p[J, 0] = p[J, 2]
    u[J, m + 1] = u[J, m - 1]
    v[J, m + 1] = v[J, m - 1]
    p[J, m + 1] = P.outlet_pressure
    for a in (u, v, p):
        a[0, :] = a[n, :]
        a[n + 1, :] = a[1, :]
    return error


# ---------------------------------------------------------------------------------------------------
# ASCII VTK snapshot bodies: lib/output.f90:968-1088 (3D) and :421-537 (2D), format "(3(f16.4,1x))".
# gfortran formats F editing through snprintf, i.e. correctly rounded (half-to-even on the exact binary
# value) like Python's % operator; a record's trailing 1x is dropped; a value wider than 16 columns prints
# as asterisks; NaN / Infinity are right-justified words.  (Pinned against libgfortran itself: tests/test_gfortran_io.py.)
def f16_4(x: float) -> str:
    if x != x:
        return "NaN".rjust(16)
    if x in (float("inf"), float("-inf")):
        return ("-Infinity" if x < 0 else "Infinity").rjust(16)
    s = "%.4f" % x
    return s.rjust(16) if len(s) <= 16 else "*" * 16


def _records(cols) -> bytes:
    """cols: 1 or 3 flat float64 arrays of equal length -> the records of one section"""
    fm = [np.char.mod("%16.4f", c) for c in cols]
    for f, c in zip(fm, cols):
        bad = ~np.isfinite(c) | (np.char.str_len(f) > 16)
        for q in np.nonzero(bad)[0]:
            f[q] = f16_4(float(c[q]))
    line = fm[0]
    for f in fm[1:]:
        line = np.char.add(np.char.add(line, " "), f)
    return ("\n".join(line.tolist()) + "\n").encode() if len(line) else b""


def vtk_section(section: str, dim: int, u, v, w, p, e, xp, yp, zp=None, inlet_velocity=1.0) -> bytes:
    """one section body for all interior points; arrays [k][j][i] (3D) or [j][i] (2D) with halos"""
    if dim == 3:
        I = (slice(1, -1), slice(1, -1), slice(1, -1))
        sh = lambda a, di=0, dj=0, dk=0: a[1 + dk:a.shape[0] - 1 + dk, 1 + dj:a.shape[1] - 1 + dj, 1 + di:a.shape[2] - 1 + di]
        K, J, X = np.meshgrid(np.arange(1, u.shape[0] - 1), np.arange(1, u.shape[1] - 1), np.arange(1, u.shape[2] - 1), indexing="ij")
    else:
        I = (slice(1, -1), slice(1, -1))
        sh = lambda a, di=0, dj=0, dk=0: a[1 + dj:a.shape[0] - 1 + dj, 1 + di:a.shape[1] - 1 + di]
        J, X = np.meshgrid(np.arange(1, u.shape[0] - 1), np.arange(1, u.shape[1] - 1), indexing="ij")
        K = None
    zero = np.zeros(u[I].size)
    flat = lambda a: np.ascontiguousarray(a).reshape(-1)
    if section == "points":
        return _records([flat(xp[X]), flat(yp[J]), flat(zp[K]) if dim == 3 else zero])
    if section == "velocity":
        return _records([flat(u[I]), flat(v[I]), flat(w[I]) if dim == 3 else zero])
    if section == "velocityInFluid":
        return _records([flat(u[I] * e[I]), flat(v[I] * e[I]), flat(w[I] * e[I]) if dim == 3 else zero])
    if section == "dimless_v":
        return _records([flat(u[I] * e[I] / inlet_velocity), flat(v[I] * e[I] / inlet_velocity), zero])
    if section == "porosity":
        return _records([flat(e[I])])
    if section == "pressure":
        return _records([flat(p[I])])
    if section == "VelocityDivergent":
        d = (sh(u, 1) - sh(u, -1)) / (xp[X + 1] - xp[X - 1]) + (sh(v, 0, 1) - sh(v, 0, -1)) / (yp[J + 1] - yp[J - 1])
        if dim == 3:
            d = d + (sh(w, 0, 0, 1) - sh(w, 0, 0, -1)) / (zp[K + 1] - zp[K - 1])
        return _records([flat(d)])
    if section == "abs_dimless_v":
        a0, a1 = u[I] * e[I] / inlet_velocity, v[I] * e[I] / inlet_velocity
        return _records([flat(np.sqrt(a0 * a0 + a1 * a1))])
    raise KeyError(section)
