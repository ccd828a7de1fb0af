/*
 * CLbmSolver.hpp -- the solver facade of the reference (src/CLbmSolver.hpp) over the C ABI.
 *
 * Same public surface: constructor argument list, public data (drivenCavityVelocity,
 * SIZE_DD_HOST, simulation_step_counter, error + everything inherited from CLbmSkeleton) and
 * methods (reload, reset, simulationStep[Alpha|Beta], wait, addDrivenCavityValue, the store.../set... family
 * in full and rect form, getVelocityChecksum, debug_print, debugDD).  Every body is one call into
 * liblbm_b200.so (include/lbm_b200.h); a non-zero status becomes `error << message`, the
 * reference's convention (checked by the caller at src/CController.hpp:218-221).  There is no
 * CPU fallback: without a CUDA device reload() records an error.
 */
#ifndef LBM_B200_HOST_CLBMSOLVER_HPP
#define LBM_B200_HOST_CLBMSOLVER_HPP

#include <cstdio>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <list>
#include <vector>

#include "../../include/lbm_b200.h"
#include "CCL.hpp"
#include "CLbmDebug.hpp"
#include "CLbmSkeleton.hpp"
#include "common.h"

template <typename T>
class CLbmSolver : public CLbmSkeleton<T> {
public:
	CVector<4, T> drivenCavityVelocity;
	static const size_t SIZE_DD_HOST = LBM_SIZE_DD_HOST;
	size_t simulation_step_counter;
	using CLbmSkeleton<T>::error;

	/* extensions (no reference counterpart): set before reload() */
	T smagorinsky_constant;
	int beta_order;            /* LBM_BETA_ORDER_SHIPPED (default) | LBM_BETA_ORDER_LINEAR */

private:
	int _UID;
	int _BC[3][2];
	CCL::CCommandQueue &cCommandQueue;
	CCL::CContext &cContext;
	CCL::CDevice &cDevice;
	size_t computation_kernel_count;
	bool store_velocity, store_density;
	lbm_t _handle;

	bool ok(int status)
	{
		if (status == LBM_OK) return true;
		error << "liblbm_b200 status " << status << ": " << lbmGetLastErrorString(_handle) << std::endl;
		return false;
	}

public:
	CLbmSolver(int UID, CCL::CCommandQueue &p_cCommandQueue, CCL::CContext &p_cContext, CCL::CDevice &p_cDevice,
			int BC[3][2], CDomain<T> &domain, CVector<3, T> &p_d_gravitation, T p_d_viscosity,
			size_t p_computation_kernel_count, bool p_store_velocity, bool p_store_density, T /*p_d_timestep*/,
			CVector<4, T> &_drivenCavityVelocity, std::list<int> & /*p_lbm_opencl_number_of_work_items_list*/,
			std::list<int> & /*p_lbm_opencl_number_of_registers_list*/, T p_smagorinsky_constant = (T)0,
			int p_beta_order = LBM_BETA_ORDER_SHIPPED)
		: CLbmSkeleton<T>(CDomain<T>(domain), _drivenCavityVelocity), drivenCavityVelocity(_drivenCavityVelocity),
		  simulation_step_counter(0), smagorinsky_constant(p_smagorinsky_constant), beta_order(p_beta_order),
		  _UID(UID), cCommandQueue(p_cCommandQueue), cContext(p_cContext), cDevice(p_cDevice),
		  computation_kernel_count(p_computation_kernel_count), store_velocity(p_store_velocity),
		  store_density(p_store_density), _handle(NULL)
	{
		for (int a = 0; a < 3; a++) for (int s = 0; s < 2; s++) _BC[a][s] = BC[a][s];
		CLbmSkeleton<T>::init(p_d_gravitation, p_d_viscosity, (T)1.0);
		/* tau out of range: recorded, and like the reference (src/CLbmSolver.hpp:238-260) the
		 * buffers are still created; the caller checks error() (src/CController.hpp:218-221) */
		const bool unstable = CLbmSkeleton<T>::error();
		reload();
		(void)unstable;
	}

	~CLbmSolver() { if (_handle) lbmDestroy(_handle); }

	lbm_t handle() const { return _handle; }
	int getUid() const { return _UID; }

	/* src/CLbmSolver.hpp:262-270 */
	void addDrivenCavityValue(T value)
	{
		drivenCavityVelocity[0] += value;
		ok(lbmSetDrivenCavityVelocity(_handle, (double)(drivenCavityVelocity[0] * this->d_timestep)));
	}

	/* src/CLbmSolver.hpp:272-617: buffers + kernel specialisation; ends with reset() */
	void reload()
	{
		if (_handle) { lbmDestroy(_handle); _handle = NULL; }
		lbm_desc d;
		std::memset(&d, 0, sizeof(d));
		d.struct_size = sizeof(d);
		d.device = cDevice.ordinal;
		d.dtype = sizeof(T) == 4 ? LBM_F32 : LBM_F64;
		for (int a = 0; a < 3; a++) { d.size[a] = this->domain_cells[a]; d.gravitation[a] = (double)this->gravitation[a]; }
		for (int a = 0; a < 3; a++) for (int s = 0; s < 2; s++) d.bc[2 * a + s] = _BC[a][s];
		d.inv_tau = (double)this->inv_tau;
		d.tau = (double)this->tau;
		d.u_lid = (double)CLbmSkeleton<T>::drivenCavityVelocity[0];   /* kernel arg 8, src/CLbmSolver.hpp:597,608 */
		d.store_velocity = store_velocity;
		d.store_density = store_density;
		d.smagorinsky_cs = (double)smagorinsky_constant;
		d.beta_order = beta_order;
		d.work_group_size = (int)computation_kernel_count;
		d.compute_stream = cCommandQueue.compute_stream;
		d.comm_stream = cCommandQueue.comm_stream;
		if (!ok(lbmCreate(&_handle, &d))) { _handle = NULL; return; }
		simulation_step_counter = 0;
	}

	void reset() { if (ok(lbmReset(_handle))) simulation_step_counter = 0; }

	void simulationStepAlpha() { ok(lbmStepAlpha(_handle)); }
	void simulationStepBeta() { ok(lbmStepBeta(_handle)); }
	/* src/CLbmSolver.hpp:664-676: odd counter -> alpha, even -> beta */
	void simulationStep()
	{
		if (ok(lbmStep(_handle))) simulation_step_counter++;
	}
	/* CController::computeNextStep with the device-resident halo exchange fused in */
	void simulationStepWithHalo()
	{
		if (ok(lbmCommStep(_handle))) simulation_step_counter++;
	}
	void wait() { ok(lbmWait(_handle)); }

	/* ---- populations: src/CLbmSolver.hpp:688-757 */
	void storeDensityDistribution(T *dst) { ok(lbmStoreDD(_handle, dst, NULL, NULL)); }
	void storeDensityDistribution(T *dst, CVector<3, int> &origin, CVector<3, int> &size)
	{
		ok(lbmStoreDD(_handle, dst, origin.data, size.data));
	}
	void setDensityDistribution(T *src, CVector<3, int> &origin, CVector<3, int> &size)
	{
		ok(lbmSetDD(_handle, src, origin.data, size.data, NULL));
	}
	void setDensityDistribution(T *src, CVector<3, int> &origin, CVector<3, int> &size, CVector<3, int> norm)
	{
		ok(lbmSetDD(_handle, src, origin.data, size.data, norm.data));
	}

	/* ---- velocity / density / flags: src/CLbmSolver.hpp:764-978 */
	void storeVelocity(T *dst) { ok(lbmStoreVelocity(_handle, dst, NULL, NULL)); }
	void storeVelocity(T *dst, CVector<3, int> &origin, CVector<3, int> &size) { ok(lbmStoreVelocity(_handle, dst, origin.data, size.data)); }
	void setVelocity(T *src, CVector<3, int> &origin, CVector<3, int> &size) { ok(lbmSetVelocity(_handle, src, origin.data, size.data)); }
	void storeDensity(T *dst) { ok(lbmStoreDensity(_handle, dst, NULL, NULL)); }
	void storeDensity(T *dst, CVector<3, int> &origin, CVector<3, int> &size) { ok(lbmStoreDensity(_handle, dst, origin.data, size.data)); }
	void setDensity(T *src, CVector<3, int> &origin, CVector<3, int> &size) { ok(lbmSetDensity(_handle, src, origin.data, size.data)); }
	void storeFlags(int *dst) { ok(lbmStoreFlags(_handle, dst, NULL, NULL)); }
	void storeFlags(int *dst, CVector<3, int> &origin, CVector<3, int> &size) { ok(lbmStoreFlags(_handle, dst, origin.data, size.data)); }
	void setFlags(int *src, CVector<3, int> &origin, CVector<3, int> &size) { ok(lbmSetFlags(_handle, src, origin.data, size.data)); }

	/* src/CLbmSolver.hpp:1103-1123 (float accumulator, index order: bit-compatible) */
	float getVelocityChecksum()
	{
		double v = 0.0;
		ok(lbmChecksumVelocity(_handle, &v, 1));
		return (float)v;
	}
	/* the same sum reduced on the device (warp shuffles), double accumulator */
	double getVelocityChecksumDevice()
	{
		double v = 0.0;
		ok(lbmChecksumVelocity(_handle, &v, 0));
		return v;
	}

	/* src/CLbmSolver.hpp:1032-1057: dump of all four device arrays, for tiny domains only
	 * (the caller's guard is <= 512 cells, src/CController.hpp:439-443) */
	void debug_print()
	{
		const size_t n = (size_t)this->domain_cells.elements();
		std::vector<T> dd(n * SIZE_DD_HOST), vel(n * 3), rho(n);
		std::vector<int> fl(n);
		storeDensityDistribution(dd.data());
		storeVelocity(vel.data());
		storeDensity(rho.data());
		storeFlags(fl.data());
		lbm_debug::debugPrint(std::cout, dd.data(), vel.data(), rho.data(), fl.data(), n);
	}

	/* src/CLbmSolver.hpp:1062-1101: one slot of the populations */
	void debugDD(size_t dd_id = 0, size_t wrap_size = 16, size_t empty_line = 16)
	{
		const size_t n = (size_t)this->domain_cells.elements();
		std::vector<T> dd(n * SIZE_DD_HOST);
		storeDensityDistribution(dd.data());
		lbm_debug::debugDD(std::cout, dd.data(), n, dd_id, wrap_size, empty_line);
	}
};

#endif
