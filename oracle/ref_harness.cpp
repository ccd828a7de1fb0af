/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Executes the reference's OpenCL kernels -- compiled as C++ from the sources under
 * /root/reference by oracle/build_ref.py, see ref_shim.h -- with OpenCL NDRange
 * semantics on the CPU:
 *   - init_kernel / lbm_kernel_alpha / copy_buffer_rect have no barriers: one call per
 *     work-item (reference launch shapes: src/CLbmSolver.hpp:172-209, 619-658);
 *   - lbm_kernel_beta as shipped (USE_SHARED_MEMORY 1, src/cl_programs/lbm_beta.cl:6)
 *     needs real work-group semantics (7 barriers, __local staging): every work-group
 *     of LOCAL_WORK_GROUP_SIZE items runs as that many fibers on one OS thread, a
 *     fiber yields at barrier() and the group scheduler resumes the items round-robin.
 * The table of compiled (type, size) instances is generated into ref_registry.cpp.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <omp.h>

struct ref_work_item { size_t gid[2]; size_t lid; };
thread_local ref_work_item ref_wi;

typedef void (*step_f32)(float*, const int*, float*, float*, float, float, float, float, float);
typedef void (*step_f64)(double*, const int*, double*, double*, double, double, double, double, double);
typedef void (*init_f32)(float*, int*, float*, float*, int*, float);
typedef void (*init_f64)(double*, int*, double*, double*, int*, double);

struct ref_instance {
	int dtype_bytes, sx, sy, sz, wg;
	void *init, *alpha, *beta_shm, *beta_noshm;
};
extern const ref_instance ref_instances[];
extern const int ref_instance_count;

/* ------------------------------------------------------------------ fibers */
#if !defined(__x86_64__)
#error "ref_harness.cpp: the fiber switch is written for x86-64"
#endif
extern "C" void ref_switch(void **save_sp, void *load_sp);
asm(".text\n.globl ref_switch\n.type ref_switch,@function\nref_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size ref_switch,.-ref_switch\n");

struct step_args {
	const ref_instance *inst; void *fn;
	void *dd; const int *flags; void *vel; void *rho;
	double inv_tau, gx, gy, gz, ulid;
};

static void call_step(const step_args &a)
{
	if (a.inst->dtype_bytes == 4)
		((step_f32)a.fn)((float*)a.dd, a.flags, (float*)a.vel, (float*)a.rho,
				(float)a.inv_tau, (float)a.gx, (float)a.gy, (float)a.gz, (float)a.ulid);
	else
		((step_f64)a.fn)((double*)a.dd, a.flags, (double*)a.vel, (double*)a.rho,
				a.inv_tau, a.gx, a.gy, a.gz, a.ulid);
}

struct group_ctx {
	int wg;
	std::vector<void*> sp;          /* saved stack pointer per fiber */
	std::vector<char>  done;
	std::vector<char*> stacks;
	void *sched_sp;
	int current;
	const step_args *args;
	size_t group_base;
};
static thread_local group_ctx *tl_group = NULL;
static const size_t FIBER_STACK = 64 * 1024;

extern "C" void ref_barrier(void)
{
	group_ctx *g = tl_group;
	if (!g) return;                 /* kernel called outside a work-group fiber */
	ref_switch(&g->sp[g->current], g->sched_sp);
}

static void fiber_entry(void)
{
	group_ctx *g = tl_group;
	call_step(*g->args);
	g->done[g->current] = 1;
	ref_switch(&g->sp[g->current], g->sched_sp);
	abort();                        /* a finished fiber is never resumed */
}

static void run_group(group_ctx &g, const step_args &a, size_t group_base, size_t n_items)
{
	g.args = &a; g.group_base = group_base;
	for (int i = 0; i < g.wg; i++) {
		uintptr_t top = ((uintptr_t)(g.stacks[i] + FIBER_STACK)) & ~(uintptr_t)15;
		void **s = (void**)top;
		*--s = NULL;                       /* fake return address of fiber_entry */
		*--s = (void*)fiber_entry;         /* popped by ret in ref_switch */
		for (int r = 0; r < 6; r++) *--s = NULL;
		g.sp[i] = (void*)s;
		g.done[i] = (group_base + i >= n_items) ? 2 : 0;  /* items past the NDRange never exist */
	}
	tl_group = &g;
	for (;;) {
		int alive = 0;
		for (int i = 0; i < g.wg; i++) {
			if (g.done[i]) continue;
			alive++;
			g.current = i;
			ref_wi.gid[0] = group_base + i; ref_wi.gid[1] = 0; ref_wi.lid = i;
			ref_switch(&g.sched_sp, g.sp[i]);
		}
		if (!alive) break;
	}
	tl_group = NULL;
}

/* ------------------------------------------------------------------ C API */
extern "C" {

int ref_count(void) { return ref_instance_count; }

void ref_describe(int idx, int out[5])
{
	const ref_instance &r = ref_instances[idx];
	out[0] = r.dtype_bytes; out[1] = r.sx; out[2] = r.sy; out[3] = r.sz; out[4] = r.wg;
}

int ref_lookup(int dtype_bytes, int sx, int sy, int sz, int wg)
{
	for (int i = 0; i < ref_instance_count; i++) {
		const ref_instance &r = ref_instances[i];
		if (r.dtype_bytes == dtype_bytes && r.sx == sx && r.sy == sy && r.sz == sz && r.wg == wg)
			return i;
	}
	return -1;
}

void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int  ref_get_max_threads(void) { return omp_get_max_threads(); }

void ref_init(int idx, void *dd, int *flags, void *vel, void *rho, int *bc, double ulid)
{
	const ref_instance &r = ref_instances[idx];
	const long n = (long)r.sx * r.sy * r.sz;
	#pragma omp parallel for schedule(static)
	for (long gid = 0; gid < n; gid++) {
		ref_wi.gid[0] = gid; ref_wi.gid[1] = 0; ref_wi.lid = gid % r.wg;
		if (r.dtype_bytes == 4)
			((init_f32)r.init)((float*)dd, flags, (float*)vel, (float*)rho, bc, (float)ulid);
		else
			((init_f64)r.init)((double*)dd, flags, (double*)vel, (double*)rho, bc, ulid);
	}
}

void ref_alpha(int idx, void *dd, const int *flags, void *vel, void *rho,
		double inv_tau, double gx, double gy, double gz, double ulid)
{
	const ref_instance &r = ref_instances[idx];
	const long n = (long)r.sx * r.sy * r.sz;
	step_args a = { &r, r.alpha, dd, flags, vel, rho, inv_tau, gx, gy, gz, ulid };
	#pragma omp parallel for schedule(static)
	for (long gid = 0; gid < n; gid++) {
		ref_wi.gid[0] = gid; ref_wi.gid[1] = 0; ref_wi.lid = gid % r.wg;
		call_step(a);
	}
}

/* variant 0: the kernel as shipped (shared-memory path, fibers);
 * variant 1: the reference's own USE_SHARED_MEMORY 0 path (no barriers needed). */
void ref_beta(int idx, int variant, void *dd, const int *flags, void *vel, void *rho,
		double inv_tau, double gx, double gy, double gz, double ulid)
{
	const ref_instance &r = ref_instances[idx];
	const long n = (long)r.sx * r.sy * r.sz;
	if (variant == 1) {
		step_args a = { &r, r.beta_noshm, dd, flags, vel, rho, inv_tau, gx, gy, gz, ulid };
		#pragma omp parallel for schedule(static)
		for (long gid = 0; gid < n; gid++) {
			ref_wi.gid[0] = gid; ref_wi.gid[1] = 0; ref_wi.lid = gid % r.wg;
			call_step(a);
		}
		return;
	}
	step_args a = { &r, r.beta_shm, dd, flags, vel, rho, inv_tau, gx, gy, gz, ulid };
	const long groups = (n + r.wg - 1) / r.wg;
	#pragma omp parallel
	{
		group_ctx g;
		g.wg = r.wg;
		g.sp.resize(r.wg); g.done.resize(r.wg); g.stacks.resize(r.wg);
		for (int i = 0; i < r.wg; i++) g.stacks[i] = (char*)malloc(FIBER_STACK);
		#pragma omp for schedule(static)
		for (long grp = 0; grp < groups; grp++)
			run_group(g, a, (size_t)grp * r.wg, (size_t)n);
		for (int i = 0; i < r.wg; i++) free(g.stacks[i]);
	}
}

/* copy_buffer_rect.cl instances (T = float, double); launch shape of
 * CLbmSolver::enqueueCopyRectKernel: global = (block_y, block_z), one x-row per item. */
void ref_copy_rect_f32(float*, int, int, int, int, int, int, int, float*, int, int, int, int, int, int, int, int);
void ref_copy_rect_f64(double*, int, int, int, int, int, int, int, double*, int, int, int, int, int, int, int, int);

void ref_copy_rect(int dtype_bytes, void *src, int src_off, const int so[3], const int ss[3],
		void *dst, int dst_off, const int dorg[3], const int ds[3], const int block[3])
{
	for (int k = 0; k < block[2]; k++)
		for (int j = 0; j < block[1]; j++) {
			ref_wi.gid[0] = j; ref_wi.gid[1] = k; ref_wi.lid = 0;
			if (dtype_bytes == 4)
				ref_copy_rect_f32((float*)src, src_off, so[0], so[1], so[2], ss[0], ss[1], ss[2],
						(float*)dst, dst_off, dorg[0], dorg[1], dorg[2], ds[0], ds[1], ds[2], block[0]);
			else
				ref_copy_rect_f64((double*)src, src_off, so[0], so[1], so[2], ss[0], ss[1], ss[2],
						(double*)dst, dst_off, dorg[0], dorg[1], dorg[2], ds[0], ds[1], ds[2], block[0]);
		}
}

} /* extern "C" */
