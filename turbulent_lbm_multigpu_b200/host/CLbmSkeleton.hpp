/*
 * CLbmSkeleton.hpp -- physical -> lattice parametrisation of the solver.
 *
 * Same public members and the same arithmetic, operation by operation and in the simulation
 * type T, as the reference's src/CLbmSkeleton.hpp:81-116 (updateValues) and :132-164 (init), so
 * that tau, inv_tau, the lattice gravitation and the lid velocity handed to the kernels are
 * bit-identical (known values: SURVEY.md §8 a9; tests/test_host_cpp.py).  Compile the host code
 * without FMA contraction (-ffp-contract=off), as host/build.py does.
 */
#ifndef LBM_B200_HOST_CLBMSKELETON_HPP
#define LBM_B200_HOST_CLBMSKELETON_HPP

#include <cmath>
#include <iostream>

#include "CDomain.hpp"
#include "CError.hpp"
#include "CVector.hpp"

template <typename T>
class CLbmSkeleton {
public:
	CError error;
	bool debug;

	CDomain<T> domain;
	CVector<3, int> domain_cells;          /* cells per axis, ghost layers included */
	int domain_cells_count;
	CVector<4, T> d_drivenCavityVelocity;  /* with dimension */

	T d_domain_x_length;
	CVector<3, T> d_gravitation;
	T d_viscosity;
	T mass_exchange_factor;

	T d_cell_length;
	T d_timestep;
	T d_reynolds;

	T tau, inv_tau, inv_trt_tau;
	CVector<4, T> drivenCavityVelocity;    /* lattice units: d_drivenCavityVelocity * d_timestep */
	CVector<3, T> gravitation;             /* lattice units */
	T max_sim_gravitation_length;

	CLbmSkeleton(CDomain<T> _domain, CVector<4, T> _drivenCavityVelocity)
		: debug(false), domain(_domain), domain_cells_count(0), d_drivenCavityVelocity(_drivenCavityVelocity) {}

	/* src/CLbmSkeleton.hpp:81-116 */
	void updateValues(bool info_output = false)
	{
		const T cell2 = d_cell_length * d_cell_length;
		const T sqrt_mef = (T)std::sqrt(mass_exchange_factor);
		d_timestep = cell2 * ((T)2.0 * tau - (T)1.0) / ((T)6.0 * d_viscosity * sqrt_mef);
		gravitation = d_gravitation * ((d_timestep * d_timestep) / d_cell_length);

		/* keep the lattice force small: larger values make the scheme unstable */
		if (gravitation.length() >= max_sim_gravitation_length) {
			if (info_output)
				std::cout << "limiting timestep (gravitation: " << gravitation << ")" << std::endl;
			d_timestep = (T)std::sqrt((max_sim_gravitation_length * d_cell_length) / d_gravitation.length());
			gravitation = d_gravitation * ((d_timestep * d_timestep) / d_cell_length);
			tau = (T)0.5 * (d_timestep * d_viscosity * sqrt_mef * (T)6.0) / cell2 + (T)0.5;
		}
		if (tau < 0.51 || tau > 2.5) {
			error << "tau has to be within the boundary [0.51; 2.5]" << std::endl;
			error << "otherwise the simulation becomes unstable! current value: " << tau << std::endl;
		}
		inv_tau = (T)1.0 / tau;
		inv_trt_tau = (T)1.0 / ((T)0.5 + (T)3.0 / ((T)16.0 * tau - (T)8.0));
	}

	void setGravitation(CVector<3, T> p_d_gravitation)
	{
		d_gravitation = p_d_gravitation;
		updateValues();
	}

	/* src/CLbmSkeleton.hpp:132-164 */
	void init(CVector<3, T> &p_d_gravitation, T p_d_viscosity, T p_mass_exchange_factor,
			T p_max_sim_gravitation_length = (T)0.0001, T p_tau = (T)0.953575)
	{
		domain_cells = domain.getSize();
		d_domain_x_length = domain.getLength()[0];
		d_gravitation = p_d_gravitation;
		domain_cells_count = domain_cells.elements();
		d_viscosity = p_d_viscosity;
		mass_exchange_factor = p_mass_exchange_factor;
		d_cell_length = d_domain_x_length / (T)domain_cells[0];
		max_sim_gravitation_length = p_max_sim_gravitation_length;
		tau = p_tau;

		updateValues(true);
		drivenCavityVelocity = d_drivenCavityVelocity * d_timestep;
		d_reynolds = d_domain_x_length * d_drivenCavityVelocity[0] / d_viscosity;
		if (debug) {
			std::cout << "dim cell length: " << d_cell_length << "  dim timestep: " << d_timestep << std::endl;
			std::cout << "tau: " << tau << "  inv_tau: " << inv_tau << "  gravitation: " << gravitation << std::endl;
		}
		std::cout << "dim reynolds number: " << d_reynolds << std::endl;
	}
};

#endif
