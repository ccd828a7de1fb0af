/*
 * CCL.hpp -- what is left of the reference's OpenCL wrapper layer (src/libcl/CCL.hpp, 2600
 * lines: platforms, contexts, queues, buffers, JIT programs, kernels, events) once the device
 * side is liblbm_b200.so: three plain handle types so that code written against
 * CLbmSolver's constructor (src/CLbmSolver.hpp:221-230) keeps compiling.  The library owns
 * streams, buffers and kernels; these types only carry the CUDA device ordinal and,
 * optionally, caller-owned streams.
 */
#ifndef LBM_B200_HOST_CCL_HPP
#define LBM_B200_HOST_CCL_HPP

#include "../../include/lbm_b200.h"

namespace CCL {

struct CDevice {
	int ordinal;
	explicit CDevice(int device_ordinal = 0) : ordinal(device_ordinal) {}
};

struct CContext {
	/* number of CUDA devices visible to the process (CContext::load + CDevices, src/CController.hpp:133-139) */
	static int deviceCount()
	{
		int n = 0;
		if (lbmGetDeviceCount(&n) != LBM_OK) return 0;
		return n;
	}
};

struct CCommandQueue {
	void *compute_stream;   /* cudaStream_t or NULL: the library creates its own */
	void *comm_stream;
	CCommandQueue() : compute_stream(0), comm_stream(0) {}
};

} /* namespace CCL */

#endif
