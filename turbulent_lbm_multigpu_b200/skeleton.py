"""Physical -> lattice parametrisation of the LBM solver (host logic, no device code).

Mirror of the reference's ``CLbmSkeleton<T>::init / updateValues``
(reference src/CLbmSkeleton.hpp:81-116,132-164): every operation is carried out in the
simulation type ``T`` (numpy float32 / float64 scalars), in the reference's operation
order, so that ``inv_tau``, the lattice gravitation and the lid velocity passed to the
kernels are bit-identical to what the reference's host code computes (known values:
64^3, L = 0.1 -> inv_tau = 1.42278874, u_lid = 0.0126204686, g_y = -9.99999975e-05,
SURVEY.md appendix A.2).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FLAG_OBSTACLE = 1 << 0
FLAG_FLUID = 1 << 1
FLAG_VELOCITY_INJECTION = 1 << 2
FLAG_GHOST_LAYER = 1 << 3

#: D3Q19 lattice vectors in slot order (reference src/main.cpp:38-66, lbm_header.h:15-27)
LBM_UNITS = (
    (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0),
    (1, 1, 0), (-1, -1, 0), (1, -1, 0), (-1, 1, 0),
    (1, 0, 1), (-1, 0, -1), (1, 0, -1), (-1, 0, 1),
    (0, 1, 1), (0, -1, -1), (0, 1, -1), (0, -1, 1),
    (0, 0, 1), (0, 0, -1), (0, 0, 0),
)
SIZE_DD_HOST = 19


class SkeletonError(ValueError):
    """tau left the stable interval [0.51, 2.5] (reference CLbmSkeleton.hpp:108-112)."""


@dataclass
class LbmParameters:
    dtype: type
    domain_cells: tuple
    d_cell_length: float
    d_timestep: float
    tau: float
    inv_tau: float
    inv_trt_tau: float
    gravitation: tuple
    drivenCavityVelocity: tuple
    d_reynolds: float
    error: str = field(default="")

    @property
    def u_lid(self):
        """kernel argument: CLbmSkeleton::drivenCavityVelocity[0] (CLbmSolver.hpp:572,597,608)"""
        return self.drivenCavityVelocity[0]


def compute_parameters(domain_size, domain_length, gravitation=(0.0, -9.81, 0.0),
                       viscosity=0.001308, cavity_velocity=(100.0, 0.0, 0.0, 1.0),
                       dtype=np.float32, mass_exchange_factor=1.0,
                       max_sim_gravitation_length=0.0001, tau=0.953575, strict=True):
    """CLbmSkeleton::init followed by updateValues(true), all arithmetic in ``dtype``."""
    T = np.dtype(dtype).type
    with np.errstate(all="ignore"):
        d_grav = [T(g) for g in gravitation]
        d_visc = T(viscosity)
        mef = T(mass_exchange_factor)
        max_g = T(max_sim_gravitation_length)
        tau = T(tau)
        d_cav = [T(v) for v in cavity_velocity]
        d_domain_x_length = T(domain_length[0])
        # CLbmSkeleton.hpp:157
        d_cell_length = T(d_domain_x_length / T(int(domain_size[0])))

        def vlen(v):  # CVector<3,T>::length, libmath/CVector3.hpp:185-188
            return T(np.sqrt(T(T(T(v[0] * v[0]) + T(v[1] * v[1])) + T(v[2] * v[2]))))

        # updateValues, CLbmSkeleton.hpp:83-84
        d_timestep = T(T(T(d_cell_length * d_cell_length) * T(T(T(2.0) * tau) - T(1.0)))
                       / T(T(T(6.0) * d_visc) * T(np.sqrt(mef))))
        s = T(T(d_timestep * d_timestep) / d_cell_length)
        grav = [T(g * s) for g in d_grav]
        if vlen(grav) >= max_g:  # :96-106 gravitation limiting
            d_timestep = T(np.sqrt(T(T(max_g * d_cell_length) / vlen(d_grav))))
            s = T(T(d_timestep * d_timestep) / d_cell_length)
            grav = [T(g * s) for g in d_grav]
            tau = T(T(T(T(0.5) * T(T(T(d_timestep * d_visc) * T(np.sqrt(mef))) * T(6.0)))
                      / T(d_cell_length * d_cell_length)) + T(0.5))
        error = ""
        if float(tau) < 0.51 or float(tau) > 2.5:  # :108-112 (comparison promotes to double)
            error = ("tau has to be within the boundary [0.51; 2.5]\n"
                     "otherwise the simulation becomes unstable! current value: %r\n" % float(tau))
            if strict:
                raise SkeletonError(error)
        inv_tau = T(T(1.0) / tau)
        inv_trt_tau = T(T(1.0) / T(T(0.5) + T(T(3.0) / T(T(T(16.0) * tau) - T(8.0)))))
        cav = tuple(T(v * d_timestep) for v in d_cav)  # :163
        reynolds = T(T(d_domain_x_length * d_cav[0]) / d_visc)  # :164
    return LbmParameters(dtype=T, domain_cells=tuple(int(v) for v in domain_size),
                         d_cell_length=d_cell_length, d_timestep=d_timestep, tau=tau,
                         inv_tau=inv_tau, inv_trt_tau=inv_trt_tau, gravitation=tuple(grav),
                         drivenCavityVelocity=cav, d_reynolds=reynolds, error=error)
