/*
 * ref_shim.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A ~60 line OpenCL-C -> C++ shim that lets g++ compile the reference's kernel
 * sources *where they lie* (/root/reference/src/cl_programs/*.cl, pulled in with
 * -I/root/reference, nothing is copied into this repository) so that the
 * reference's own arithmetic can be executed on the CPU:
 *
 *   g++ -x c++ -ffp-contract=off -include oracle/ref_shim.h -I/root/reference \
 *       -DREF_T=float -DDOMAIN_CELLS_X=16 ... -Dlbm_kernel_beta=<unique name> \
 *       -c /root/reference/src/cl_programs/lbm_beta.cl
 *
 * The #defines mirror what CLbmSolver::reload generates at run time
 * (reference src/CLbmSolver.hpp:319-373 plus LOCAL_WORK_GROUP_SIZE :419-422).
 *
 * Work-item identity is thread-local state set by the harness (ref_harness.cpp);
 * barrier() yields the current work-item fiber back to the work-group scheduler
 * (real OpenCL work-group semantics for the shipped shared-memory beta path),
 * and is a no-op when a kernel is called outside a fiber.
 */
#ifndef LBM_ORACLE_REF_SHIM_H
#define LBM_ORACLE_REF_SHIM_H

#include <stddef.h>
#include <math.h>

#ifndef REF_T
#define REF_T float
#endif
typedef REF_T T;
struct T4 { T x, y, z, w; };

struct ref_work_item {
	size_t gid[2];   /* get_global_id(0), get_global_id(1) */
	size_t lid;      /* get_local_id(0) */
};
extern thread_local ref_work_item ref_wi;
extern "C" void ref_barrier(void);

static inline size_t get_global_id(int d) { return ref_wi.gid[d]; }
static inline size_t get_local_id(int)    { return ref_wi.lid; }
#define barrier(x) ref_barrier()
#define CLK_LOCAL_MEM_FENCE 0

/* address-space / function qualifiers */
#define __kernel extern "C"
#define __global
/* __const is already a g++ keyword alias for const */

/*
 * __local appears twice in lbm_beta.cl: on the work-group array (:167) and on a
 * private pointer into it (:256).  The array must be shared by the work-items of
 * a group (which all run as fibers on one OS thread), the pointer must stay
 * private: select by order of appearance with __COUNTER__.
 */
#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)
#define REF_LOCAL_0 static thread_local
#define REF_LOCAL_1
#define __local REF_CAT(REF_LOCAL_, __COUNTER__)

/* each translation unit has its own DOMAIN_CELLS: keep the helper functions internal */
#define inline static inline

/* what CLbmSolver::reload prepends (flag values: reference src/common.h:19-22) */
#define GLOBAL_WORK_GROUP_SIZE (DOMAIN_CELLS_X*DOMAIN_CELLS_Y*DOMAIN_CELLS_Z)
#define FLAG_OBSTACLE (1)
#define FLAG_FLUID (2)
#define FLAG_VELOCITY_INJECTION (4)
#define FLAG_GHOST_LAYER (8)
#define SIZE_DD_HOST_BYTES (19*sizeof(T))
#define STORE_VELOCITY 1
#define STORE_DENSITY 1
#ifndef LOCAL_WORK_GROUP_SIZE
#define LOCAL_WORK_GROUP_SIZE (128)
#endif

#endif
