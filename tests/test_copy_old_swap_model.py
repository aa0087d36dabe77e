"""CPU model of `u_old = u` done by exchanging two buffers plus a halo-shell copy (do_copy_old_by_swap in
csrc/pf_api.cu, shell_copy_kernel) against the reference's full copy (ibm_3d_uniform_omp_cpu.f90:85-100), through
whole time steps of the numpy restatement: the predictor rewrites every interior cell, the halos must carry the
previous step's values."""
import numpy as np
import pytest

from oracle import oracle_np as onp


def _shell_copy(dst, src):
    dst[0], dst[-1] = src[0], src[-1]
    dst[:, 0], dst[:, -1] = src[:, 0], src[:, -1]
    dst[:, :, 0], dst[:, :, -1] = src[:, :, 0], src[:, :, -1]


class SwapState(onp.State3D):
    """State3D whose step() never copies a whole velocity array"""

    def step(self):
        P, c = self.P, self.c
        if not hasattr(self, "spare"):
            # garbage on purpose: whatever the spare buffers hold inside must not matter
            self.spare = [np.full(self.u.shape, np.nan) for _ in range(3)]
        uo, vo, wo = self.u, self.v, self.w
        self.u, self.v, self.w = self.spare
        self.spare = [uo, vo, wo]
        for new, old in ((self.u, uo), (self.v, vo), (self.w, wo)):
            _shell_copy(new, old)
        onp.divergence_3d(P, self.air, uo, vo, wo, c["div"])
        onp.predictor_3d(P, uo, vo, wo, self.e, c["div"], self.u, self.v, self.w)
        onp.matrix_3d(P, self.u, self.v, self.w, self.e, c)
        if self.air:
            onp.boundary_matrix_3d_air(P, self.p, self.e, c)
        else:
            onp.boundary_matrix_3d_uniform(P, self.p, c)
        err = onp.sor_3d(P, not self.air, P.iter_max, self.p, c)
        onp.project_3d(P, self.p, self.u, self.v, self.w)
        if self.air:
            onp.boundary_3d_air(P, self.e, self.p, self.u, self.v, self.w)
        else:
            onp.boundary_3d_uniform(P, self.p, self.u, self.v, self.w)
        return err


@pytest.mark.parametrize("air", [False, True])
def test_swap_plus_shell_copy_equals_full_copy(air):
    m, n, l = 7, 6, 5
    rng = np.random.default_rng(17)
    P = onp.Params(m=m, n=n, l=l, dx=0.01, dy=0.011, dz=0.009, dt=2e-4, xnue=1e-3, xlambda=0.1, iter_max=4,
                   relux_factor=1.7, inlet_velocity=1.0, outlet_pressure=0.0, AoA=4.0)
    shape = (l + 2, n + 2, m + 2)
    e = np.zeros(shape)
    e[1:-1, 1:-1, 1:-1] = np.clip(rng.random((l, n, m)), 1e-6, 1.0)
    (onp.porosity_halo_3d_wall if air else onp.porosity_halo_3d_uniform)(P, e)
    f = {nm: 0.1 * rng.standard_normal(shape) for nm in ("p", "u", "v", "w")}
    a = onp.State3D(P, air, e.copy(), **{k: v.copy() for k, v in f.items()})
    b = SwapState(P, air, e.copy(), **{k: v.copy() for k, v in f.items()})
    for _ in range(4):          # an even and an odd number of exchanges
        ea, eb = a.step(), b.step()
        assert ea == eb
        for nm in ("u", "v", "w", "p"):
            assert np.array_equal(getattr(a, nm), getattr(b, nm)), nm
