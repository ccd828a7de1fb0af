/*
 * lbm_capi.cu -- host side of liblbm_b200.so: the C ABI of include/lbm_b200.h over the CUDA
 * runtime.  Replaces the reference's OpenCL plumbing (src/libcl/CCL.hpp) and the bodies of
 * CLbmSolver<T> (src/CLbmSolver.hpp); see the header for the per-entry-point mapping.
 * There is deliberately no CPU fallback: without a CUDA device lbmCreate fails.
 */
#include "lbm_kernels.cuh"
#include "../../include/lbm_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   /* header-only NVTX v3: no link dependency */

using namespace lbm;

namespace {

thread_local std::string g_last_error;

/* iteration box; zb/nzb: optional second z range [zb, zb+nzb) above the first, same x/y extent,
 * updated by the SAME launch (the two z shells of a slab) */
struct Box { int x0, nx, y0, ny, z0, nz, zb, nzb; };

} // namespace

struct lbm_face {
	int dst_rank;
	int send_origin[3], recv_origin[3], size[3], dir[3];
	int axis;
	uint32_t send_mask[2], recv_mask[2], write_mask[2];   /* per sync kind (alpha, beta) */
	size_t stage_off[2], stage_elems[2];                  /* layout of the RECEIVER block I own */
	size_t peer_stage_off[2];                             /* layout of the block I write into */
	char *local_block; size_t local_bytes;                /* [flags 256 B][alpha staging][beta staging] */
	char *peer_block; bool peer_is_ipc; bool connected;
	/* device words of this face: [0..1] syncs pushed per kind, [2..3] syncs pulled per kind,
	 * [4] block counter of the running push launch.  The sequence numbers live on the device so
	 * that a captured CUDA graph of the step keeps counting when it is replayed. */
	unsigned int *counters;
};

struct lbm_solver {
	lbm_desc desc;
	std::vector<lbm_face> faces;
	int axis_order;              /* LBM_AXIS_ORDER_*: phase order of a sync */
	int x_in_kernel;             /* 1 + sync kind whose x faces the last step kernels pushed themselves (XFUSE), 0 = none */
	int x_pending;               /* 1 + sync kind whose x faces were received but not scattered into dd: the next
	                                XFUSE step reads them out of the receive block; anything else flushes first */
	int xfuse;                   /* fused x push: 0 off, 1 rows that a block size divides, 2 any row (LBM_B200_XFUSE) */
	unsigned int *d_error;       /* device word: a halo wait gave up (neighbour never arrived) */
	unsigned long long wait_timeout_ns;
	int device;
	int dtype;
	int sx, sy, sz;
	long long n;
	long long stride;            /* slot stride of dd in cells (n + padding) */
	size_t elem;                 /* sizeof(T) */
	void *dd, *velocity, *density;
	int *flags;
	void *staging; size_t staging_bytes;
	double *d_checksum;
	cudaStream_t compute, comm;
	int xshell;                  /* width (cells) of the shell next to an x ghost face */
	cudaStream_t step_aux;       /* non-NULL while the comm stream is forked off the compute stream */
	bool own_compute, own_comm;
	cudaEvent_t ev_compute, ev_comm, ev_t0, ev_t1, ev_h2d;
	uint64_t counter;
	uint64_t launches;
	int vec;
	int block;
	int wg_quirk;                /* >0 only when the work-group x-shift is observable */
	bool smag;
	double u_lid;
	std::string error;
	/* per-kernel device timeline (replaces CL_QUEUE_PROFILING_ENABLE + CProfilerEvent,
	 * src/libcl/CCL.hpp:1752-1778, src/libtools/CProfilerEvent.hpp:29-38) */
	int profile_mode;            /* LBM_PROFILE_EVENTS | LBM_PROFILE_NVTX */
	cudaEvent_t prof_base;       /* time zero: recorded on the compute stream by lbmProfileEnable */
	struct prof_rec { const char *name; cudaEvent_t e0, e1; };
	std::vector<prof_rec> prof;
	std::vector<cudaEvent_t> prof_pool;   /* recycled by lbmProfileClear */
	uint64_t prof_dropped;
};

namespace {

int fail(lbm_t h, int code, const std::string &msg)
{
	if (h) h->error = msg;
	g_last_error = msg;
	return code;
}

#define CUDA_TRY(h, expr)                                                                     \
	do {                                                                                      \
		cudaError_t e__ = (expr);                                                             \
		if (e__ != cudaSuccess)                                                               \
			return fail((h), LBM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
	} while (0)

#define CHECK_HANDLE(h) \
	do { if (!(h)) return fail(NULL, LBM_ERR_INVALID, "null handle"); } while (0)

int use_device(lbm_t h) { CUDA_TRY(h, cudaSetDevice(h->device)); return LBM_OK; }

const size_t kProfMaxEvents = 1u << 18;

cudaEvent_t prof_event(lbm_t h)
{
	cudaEvent_t e = NULL;
	if (!h->prof_pool.empty()) { e = h->prof_pool.back(); h->prof_pool.pop_back(); return e; }
	if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return NULL; }
	return e;
}

/* Brackets ONE kernel launch: counts it and, when profiling is on, gives it an NVTX range and a
 * pair of CUDA events on the launching stream.  Unlike the reference (clWaitForEvents after every
 * enqueue, CCL.hpp:1765) nothing blocks: the events are resolved when they are read. */
struct LaunchScope {
	lbm_t h; cudaStream_t s; cudaEvent_t e1; bool nvtx;
	LaunchScope(lbm_t h_, const char *name, cudaStream_t s_) : h(h_), s(s_), e1(NULL), nvtx(false)
	{
		if (!h->profile_mode) return;
		if (h->profile_mode & LBM_PROFILE_NVTX) { nvtxRangePushA(name); nvtx = true; }
		if (h->profile_mode & LBM_PROFILE_EVENTS) {
			if (h->prof.size() >= kProfMaxEvents) { h->prof_dropped++; return; }
			cudaEvent_t e0 = prof_event(h); e1 = prof_event(h);
			if (!e0 || !e1) { if (e0) h->prof_pool.push_back(e0); if (e1) h->prof_pool.push_back(e1); e1 = NULL; h->prof_dropped++; return; }
			cudaEventRecord(e0, s);
			h->prof.push_back(lbm_solver::prof_rec{ name, e0, e1 });
		}
	}
	~LaunchScope()
	{
		h->launches++;
		if (e1) cudaEventRecord(e1, s);
		if (nvtx) nvtxRangePop();
	}
};

uint32_t popcount19(uint32_t m);
int flush_x_pending(lbm_t h, cudaStream_t s);

/* threads per block of the XFUSE launches (whole rows, 3-D grid): the configured size when a row is a whole
 * number of such blocks, else the largest multiple of 32 (>= 64) below it that divides a row -- measured good:
 * 384-cell rows in 96-thread blocks cost what 256- and 512-cell rows cost in 128-thread blocks.
 * *exact = false: no such size.  Rows are then fused only on request (LBM_B200_XFUSE=2: the size that leaves the
 * fewest idle threads, last block of a row partly empty): 288..320-cell rows in 160-thread blocks cost 16-25 % of
 * a step against 12-13 % with the separate x push/pull kernels (profiles/r2/xfuse_row_lengths.md). */
int x_block(lbm_t h, bool *exact = NULL)
{
	const int groups = h->sx / h->vec;
	if (exact) *exact = true;
	if (groups % h->block == 0) return h->block;
#if LBM_XFUSE_GRID3D
	for (int b = h->block - 32; b >= 64; b -= 32) if (groups % b == 0) return b;
	if (exact) *exact = false;
	int best = h->block;
	long long best_idle = -1;
	for (int b = 256; b >= 64; b -= 32) {
		const long long idle = (long long)((groups + b - 1) / b) * b - groups;
		if (best_idle < 0 || idle < best_idle || (idle == best_idle && std::abs(b - h->block) < std::abs(best - h->block))) {
			best = b; best_idle = idle;
		}
	}
	return best;
#else
	if (exact) *exact = false;
	return h->block;
#endif
}

/* fused x push: only with the z,y,x phase order, the 5-slot payload and every x face connected */
bool x_fusable(lbm_t h)
{
	if (!h->xfuse || h->axis_order != LBM_AXIS_ORDER_ZYX) return false;
	/* one lane per thread (x = 1 and x = sx-2 in different groups), 32-bit face offsets */
	if (h->sx % h->vec != 0 || h->sx < 2 * h->vec + 2 || 5LL * h->sy * h->sz >= 0x7fffffffLL) return false;
	bool exact;
	const int xb = x_block(h, &exact);
	if (!exact && h->xfuse < 2) return false;                /* rows that no block size divides: on request only */
	const long long bpr = (h->sx / h->vec + xb - 1) / xb;
	/* lane key = (block in row << 10) | thread */
	if (xb > 1024 || bpr > (1 << 20)) return false;
#if LBM_XFUSE_GRID3D
	if (h->sy > 65535 || h->sz > 65535) return false;          /* grid (blocks of a row, rows, z rows) */
#else
	/* rows must be whole blocks; row of a block = blockIdx.x / blocks per row by a multiply that is exact
	 * while blockIdx.x * blocks per row < 2^32 */
	if ((h->sx / h->vec) % xb != 0 || bpr * bpr * h->sy >= 0x100000000LL) return false;
#endif
	bool any = false;
	for (size_t i = 0; i < h->faces.size(); i++) {
		const lbm_face &f = h->faces[i];
		if (f.axis != 0) continue;
		if (!f.connected || popcount19(f.send_mask[0]) != 5 || popcount19(f.send_mask[1]) != 5) return false;
		any = true;
	}
	return any;
}

template <typename T>
StepParams<T> make_params(lbm_t h, const Box &b, bool alpha, bool xfuse)
{
	StepParams<T> P;
	P.xstage[0] = P.xstage[1] = NULL;
	P.xpull[0] = P.xpull[1] = NULL;
	P.xface_n = (long long)h->sy * h->sz;
	{
		const int fn = (int)P.xface_n;                          /* 5 * fn < 2^31: x_fusable */
		const int dy[5] = { 0, -1, 1, 0, 0 }, dz[5] = { 0, 0, 0, -1, 1 };   /* (e_y, e_z) of the low lane's slots 1,5,7,9,11 */
		for (int k = 0; k < 5; k++) {
			P.xoff[0][k] = k * fn + dy[k] + dz[k] * h->sy;
			P.xoff[1][k] = k * fn - dy[k] - dz[k] * h->sy;
			P.xoff[2][k] = k * fn;
		}
	}
	{
		const int xb = x_block(h), groups = h->sx / h->vec;
		const int g_hi = (h->sx - 2) / h->vec;                  /* group of the cell next to the high face */
		P.xbpr = (unsigned)((groups + xb - 1) / xb);
		if (P.xbpr < 1) P.xbpr = 1;
		P.xmagic = P.xbpr == 1 ? 0u : (unsigned)((0x100000000ULL + P.xbpr - 1) / P.xbpr);
		P.xkey_hi = ((unsigned)(g_hi / xb) << 10) | (unsigned)(g_hi % xb);
	}
	if (xfuse) {
		const int kind = alpha ? LBM_SYNC_ALPHA : LBM_SYNC_BETA;          /* the sync this step feeds */
		const int consumed = alpha ? LBM_SYNC_BETA : LBM_SYNC_ALPHA;      /* the sync this step consumes */
		for (size_t i = 0; i < h->faces.size(); i++) {
			const lbm_face &f = h->faces[i];
			if (f.axis != 0) continue;
			const int side = f.dir[0] > 0 ? 0 : 1;
			P.xstage[side] = (T *)(f.peer_block + f.peer_stage_off[kind]);
			if (h->x_pending == 1 + consumed) P.xpull[side] = (const T *)(f.local_block + f.stage_off[consumed]);
		}
		/* timing experiments only (results are wrong without the exchange): LBM_B200_XFUSE_DEBUG=nopush|nopull|none */
		static const char *dbg = getenv("LBM_B200_XFUSE_DEBUG");
		if (dbg && (!strcmp(dbg, "nopush") || !strcmp(dbg, "none"))) P.xstage[0] = P.xstage[1] = NULL;
		if (dbg && (!strcmp(dbg, "nopull") || !strcmp(dbg, "none"))) P.xpull[0] = P.xpull[1] = NULL;
	}
	P.dd = (T *)h->dd; P.flags = h->flags; P.velocity = (T *)h->velocity; P.density = (T *)h->density;
	P.n = h->n; P.ns = h->stride; P.sx = h->sx; P.sy = h->sy; P.sz = h->sz; P.sxy = (long long)h->sx * h->sy;
	P.inv_tau = (T)h->desc.inv_tau; P.tau = (T)h->desc.tau;
	/* smag_k = 18*sqrt(2)*C_s^2 evaluated in double, rounded once to T (oracle/port.py) */
	P.smag_k = (T)(18.0 * std::sqrt(2.0) * h->desc.smagorinsky_cs * h->desc.smagorinsky_cs);
	P.gx = (T)h->desc.gravitation[0]; P.gy = (T)h->desc.gravitation[1]; P.gz = (T)h->desc.gravitation[2];
	P.u_lid = (T)h->u_lid;
	P.x0 = b.x0; P.nx = b.nx; P.y0 = b.y0; P.ny = b.ny; P.z0 = b.z0; P.nz = b.nz + b.nzb;
	if (b.nzb > 0) { P.zsplit = b.nz; P.zjump = b.zb - (b.z0 + b.nz); }
	else { P.zsplit = 0x7fffffff; P.zjump = 0; }
	P.wg = h->wg_quirk;
	P.store_v = h->desc.store_velocity; P.store_r = h->desc.store_density;
	static const int e[18][3] = { { 1, 0, 0 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, -1, 0 }, { 1, 1, 0 }, { -1, -1, 0 },
		{ 1, -1, 0 }, { -1, 1, 0 }, { 1, 0, 1 }, { -1, 0, -1 }, { 1, 0, -1 }, { -1, 0, 1 }, { 0, 1, 1 }, { 0, -1, -1 },
		{ 0, 1, -1 }, { 0, -1, 1 }, { 0, 0, 1 }, { 0, 0, -1 } };      /* src/main.cpp:38-66 */
	for (int i = 0; i < 18; i++) P.boff[i] = (long long)i * P.ns + e[i][0] + (long long)e[i][1] * P.sx + (long long)e[i][2] * P.sxy;
	return P;
}

template <typename T, int VEC, bool SMAG, bool STORE, bool XPUSH>
void launch_alpha(lbm_t h, const StepParams<T> &P, dim3 grid, dim3 block, cudaStream_t s, cudaStream_t)
{
	LaunchScope ls(h, "lbm_kernel_alpha", s);
	lbm_alpha_kernel<T, VEC, SMAG, STORE, XPUSH><<<grid, block, 0, s>>>(P);
}

template <typename T, int VEC, bool SMAG, bool STORE, bool XPUSH>
void launch_beta(lbm_t h, const StepParams<T> &P, dim3 grid, dim3 block, cudaStream_t s, cudaStream_t sg)
{
	const bool shipped = h->desc.beta_order == LBM_BETA_ORDER_SHIPPED;
	/* general kernel: only z planes whose blocks can reach across the array ends
	 * (|delta| <= sxy + sx + 1 -> planes 0,1 and sz-2,sz-1 when sy >= 2; one more when sy == 1),
	 * or everything with the quirk */
	const int reach_planes = (int)((P.sxy + P.sx + 1 + P.sxy - 1) / P.sxy);
	const int zlo_end = P.wg > 0 ? P.sz : reach_planes, zhi_begin = P.wg > 0 ? P.sz : P.sz - reach_planes;
	/* the box is one z range [a0,a1) or two ([a0,a1) below [b0,b1)): the low general planes can
	 * only lie in the first piece, the high ones only in the last */
	const bool two = P.zsplit < P.nz;
	const int a0 = P.z0, a1 = P.z0 + (two ? P.zsplit : P.nz);
	const int b0 = two ? a1 + P.zjump : a0, b1 = two ? P.z0 + P.nz + P.zjump : a1;
	int ranges[2][2] = { { a0, (a1 < zlo_end ? a1 : zlo_end) },
	                     { (b0 > zhi_begin ? b0 : zhi_begin), b1 } };
	if (P.wg > 0 && two) { ranges[0][1] = a1; ranges[1][0] = b0; }     /* quirk: every plane is general */
	if (!two && ranges[1][0] < ranges[0][1]) ranges[1][0] = ranges[0][1];
	/* both z ranges in ONE launch: grid rows [0, nlo) -> low planes, [nlo, nlo+nhi) -> high planes */
	const int nlo = ranges[0][1] > ranges[0][0] ? ranges[0][1] - ranges[0][0] : 0;
	const int nhi = ranges[1][1] > ranges[1][0] ? ranges[1][1] - ranges[1][0] : 0;
	if (nlo + nhi > 0) {
		StepParams<T> Q = P;
		if (nlo > 0) { Q.z0 = ranges[0][0]; Q.zsplit = nlo; Q.zjump = ranges[1][0] - (ranges[0][0] + nlo); }
		else { Q.z0 = ranges[1][0]; Q.zsplit = 0x7fffffff; Q.zjump = 0; }
		Q.nz = nlo + nhi;
		const dim3 g2 = (XPUSH && LBM_XFUSE_GRID3D) ? dim3(grid.x, grid.y, (unsigned)Q.nz) : dim3(grid.x, (unsigned)Q.nz);
		/* with the quirk live this launch IS the whole beta step */
		LaunchScope ls(h, P.wg > 0 ? "lbm_kernel_beta" : "lbm_kernel_beta.wrap", sg);
		if (shipped) lbm_beta_general_kernel<T, VEC, SMAG, STORE, 0, XPUSH><<<g2, block, 0, sg>>>(Q);
		else lbm_beta_general_kernel<T, VEC, SMAG, STORE, 1, XPUSH><<<g2, block, 0, sg>>>(Q);
	}
	/* the vectorised kernel second: when sg is another stream the small wrapping kernel is
	 * already resident and both run side by side */
	if (P.wg == 0) {   /* with the work-group quirk live every block takes the general path */
		LaunchScope ls(h, "lbm_kernel_beta", s);
		if (shipped) lbm_beta_kernel<T, VEC, SMAG, STORE, 0, XPUSH><<<grid, block, 0, s>>>(P);
		else lbm_beta_kernel<T, VEC, SMAG, STORE, 1, XPUSH><<<grid, block, 0, s>>>(P);
	}
}

template <typename T, int VEC>
int launch_step_tv(lbm_t h, bool alpha, const Box &b, cudaStream_t s, cudaStream_t sg, bool xpush)
{
	if (b.nx <= 0 || b.ny <= 0 || b.nz + b.nzb <= 0) return LBM_OK;
	/* the fused x exchange works on whole rows (the boxes of a z,y,x step are) */
	if (xpush && !(b.nx == h->sx && b.x0 == 0)) xpush = false;
	const StepParams<T> P = make_params<T>(h, b, alpha, xpush);
	const long long groups = ((long long)b.nx * b.ny) / VEC;
	const int threads = xpush ? x_block(h) : h->block;
	dim3 block(threads);
	/* (blocks of a plane of the box, z rows); XFUSE launches (LBM_XFUSE_GRID3D): (blocks of a row, rows, z rows) */
	const dim3 grid = (xpush && LBM_XFUSE_GRID3D) ? dim3((unsigned)((h->sx / VEC + threads - 1) / threads), (unsigned)b.ny, (unsigned)(b.nz + b.nzb))
	                        : dim3((unsigned)((groups + threads - 1) / threads), (unsigned)(b.nz + b.nzb));
	const bool store = h->desc.store_velocity || h->desc.store_density;
#define LBM_DISPATCH(FN, XP)                                          \
	do {                                                              \
		if (h->smag) { if (store) FN<T, VEC, true, true, XP>(h, P, grid, block, s, sg);   \
		               else       FN<T, VEC, true, false, XP>(h, P, grid, block, s, sg); }\
		else         { if (store) FN<T, VEC, false, true, XP>(h, P, grid, block, s, sg);  \
		               else       FN<T, VEC, false, false, XP>(h, P, grid, block, s, sg); } \
	} while (0)
	if (xpush) { if (alpha) LBM_DISPATCH(launch_alpha, true); else LBM_DISPATCH(launch_beta, true); }
	else       { if (alpha) LBM_DISPATCH(launch_alpha, false); else LBM_DISPATCH(launch_beta, false); }
#undef LBM_DISPATCH
	CUDA_TRY(h, cudaGetLastError());
	return LBM_OK;
}

/* s: stream of the step kernel; sg: stream of beta's wrapping ("general") kernel -- the two
 * touch disjoint (slot, location) pairs, so sg may be a stream that runs concurrently with s */
int launch_step(lbm_t h, bool alpha, const Box &b, cudaStream_t s, cudaStream_t sg = NULL, bool xpush = false)
{
	if (!sg) sg = s;
	/* a lazily pulled x face is consumed by the XFUSE step of the matching parity; anything else needs it in dd */
	if (h->x_pending && (!xpush || h->x_pending != 1 + (alpha ? LBM_SYNC_BETA : LBM_SYNC_ALPHA)))
		if (int rc = flush_x_pending(h, s)) return rc;
	if (h->dtype == LBM_F32) {
		switch (h->vec) {
		case 4: return launch_step_tv<float, 4>(h, alpha, b, s, sg, xpush);
		case 2: return launch_step_tv<float, 2>(h, alpha, b, s, sg, xpush);
		default: return launch_step_tv<float, 1>(h, alpha, b, s, sg, xpush);
		}
	}
	switch (h->vec) {
	case 2: return launch_step_tv<double, 2>(h, alpha, b, s, sg, xpush);
	default: return launch_step_tv<double, 1>(h, alpha, b, s, sg, xpush);
	}
}

Box full_box(lbm_t h) { Box b = { 0, h->sx, 0, h->sy, 0, h->sz, 0, 0 }; return b; }

/*
 * Partition of the sub-domain for communication/computation overlap: the shell is every
 * cell within the two outermost layers of a face that has a neighbour (x faces: XS cells so
 * that whole vectors/warps stay coalesced), cut into disjoint boxes; the interior is the rest.
 * ghost_faces bit (axis*2 + side).
 */
void partition(lbm_t h, int ghost_faces, std::vector<Box> &shell, Box &interior)
{
	const int S[3] = { h->sx, h->sy, h->sz };
	int lo[3], hi[3];
	for (int a = 0; a < 3; a++) {
		int t = 2;
		if (a == 0) { t = h->xshell; while (t > 2 && (t > S[0] / 4 || (S[0] % t) != 0)) t >>= 1; if (t < 2) t = 2; t = (t + h->vec - 1) / h->vec * h->vec; }   /* whole vectors: box x0/nx stay multiples of VEC */
		lo[a] = (ghost_faces >> (2 * a)) & 1 ? t : 0;
		hi[a] = (ghost_faces >> (2 * a + 1)) & 1 ? S[a] - t : S[a];
		if (lo[a] > hi[a]) { lo[a] = 0; hi[a] = 0; }   /* everything is shell */
	}
	shell.clear();
	/* z shells: full x,y */
	if (lo[2] > 0 && hi[2] < S[2] && hi[2] > lo[2])      /* both z shells: one launch */
		shell.push_back(Box{ 0, S[0], 0, S[1], 0, lo[2], hi[2], S[2] - hi[2] });
	else {
		if (lo[2] > 0) shell.push_back(Box{ 0, S[0], 0, S[1], 0, lo[2], 0, 0 });
		if (hi[2] < S[2]) shell.push_back(Box{ 0, S[0], 0, S[1], hi[2], S[2] - hi[2], 0, 0 });
	}
	const int z0 = lo[2], nz = hi[2] - lo[2];
	/* y shells: full x, interior z */
	if (lo[1] > 0) shell.push_back(Box{ 0, S[0], 0, lo[1], z0, nz, 0, 0 });
	if (hi[1] < S[1]) shell.push_back(Box{ 0, S[0], hi[1], S[1] - hi[1], z0, nz, 0, 0 });
	const int y0 = lo[1], ny = hi[1] - lo[1];
	/* x shells: interior y,z */
	if (lo[0] > 0) shell.push_back(Box{ 0, lo[0], y0, ny, z0, nz, 0, 0 });
	if (hi[0] < S[0]) shell.push_back(Box{ hi[0], S[0] - hi[0], y0, ny, z0, nz, 0, 0 });
	interior = Box{ lo[0], hi[0] - lo[0], y0, ny, z0, nz, 0, 0 };
}

int ensure_staging(lbm_t h, size_t bytes)
{
	if (bytes <= h->staging_bytes) return LBM_OK;
	if (h->staging) { CUDA_TRY(h, cudaFree(h->staging)); h->staging = NULL; h->staging_bytes = 0; }
	CUDA_TRY(h, cudaMalloc(&h->staging, bytes));
	h->staging_bytes = bytes;
	return LBM_OK;
}

int ensure_field(lbm_t h, void **buf, int comps)
{
	if (*buf) return LBM_OK;
	const size_t bytes = (size_t)comps * h->n * h->elem;
	CUDA_TRY(h, cudaMalloc(buf, bytes));
	CUDA_TRY(h, cudaMemsetAsync(*buf, 0, bytes, h->compute));
	return LBM_OK;
}

int check_rect(lbm_t h, const int origin[3], const int size[3])
{
	const int S[3] = { h->sx, h->sy, h->sz };
	for (int a = 0; a < 3; a++)
		if (origin[a] < 0 || size[a] <= 0 || origin[a] + size[a] > S[a])
			return fail(h, LBM_ERR_INVALID, "rect outside the sub-domain");
	return LBM_OK;
}

template <typename T>
void launch_rect(lbm_t h, const T *src, T *dst, const RectCopy &R, cudaStream_t s)
{
	const long long total = (long long)R.block[0] * R.block[1] * R.block[2] * R.ncomp;
	const int block = 256;
	long long grid = (total + block - 1) / block;
	if (grid > 148LL * 16) grid = 148LL * 16;
	if (grid < 1) grid = 1;
	LaunchScope ls(h, "copy_buffer_rect", s);
	rect_copy_kernel<T><<<(unsigned)grid, block, 0, s>>>(src, dst, R);
}

void launch_rect_bytes(lbm_t h, size_t elem, const void *src, void *dst, const RectCopy &R, cudaStream_t s)
{
	if (elem == 4) launch_rect<float>(h, (const float *)src, (float *)dst, R, s);
	else launch_rect<double>(h, (const double *)src, (double *)dst, R, s);
}

/* field (ncomp_total components of n cells) rect  <->  packed buffer [comp][z][y][x] */
RectCopy rect_desc(lbm_t h, const int origin[3], const int size[3], bool to_packed,
		const int *field_comps, const int *packed_comps, int ncomp, long long field_stride = 0)
{
	if (field_stride == 0) field_stride = h->n;
	RectCopy R;
	const long long cells = (long long)size[0] * size[1] * size[2];
	const int zero[3] = { 0, 0, 0 };
	const int S[3] = { h->sx, h->sy, h->sz };
	for (int a = 0; a < 3; a++) {
		R.block[a] = size[a];
		R.so[a] = to_packed ? origin[a] : zero[a];
		R.ss[a] = to_packed ? S[a] : size[a];
		R.dorg[a] = to_packed ? zero[a] : origin[a];
		R.ds[a] = to_packed ? size[a] : S[a];
	}
	R.src_comp_stride = to_packed ? field_stride : cells;
	R.dst_comp_stride = to_packed ? cells : field_stride;
	R.ncomp = ncomp;
	for (int c = 0; c < ncomp; c++) {
		R.src_comp[c] = to_packed ? field_comps[c] : packed_comps[c];
		R.dst_comp[c] = to_packed ? packed_comps[c] : field_comps[c];
	}
	return R;
}

int store_field(lbm_t h, const void *field, size_t elem, int comps, void *host_dst,
		const int origin[3], const int size[3], long long field_stride = 0)
{
	if (field_stride == 0) field_stride = h->n;
	if (int rc = use_device(h)) return rc;
	if (!host_dst) return fail(h, LBM_ERR_INVALID, "null host pointer");
	if (!origin || !size) {
		/* no cudaMemcpy2D: its pitch is capped at 2 GiB, a sub-domain slot is not */
		if (field_stride == h->n)
			CUDA_TRY(h, cudaMemcpyAsync(host_dst, field, (size_t)comps * h->n * elem, cudaMemcpyDeviceToHost, h->compute));
		else for (int c = 0; c < comps; c++)
			CUDA_TRY(h, cudaMemcpyAsync((char *)host_dst + (size_t)c * h->n * elem, (const char *)field + (size_t)c * field_stride * elem,
					(size_t)h->n * elem, cudaMemcpyDeviceToHost, h->compute));
		CUDA_TRY(h, cudaStreamSynchronize(h->compute));
		return LBM_OK;
	}
	if (int rc = check_rect(h, origin, size)) return rc;
	const size_t bytes = (size_t)comps * size[0] * size[1] * size[2] * elem;
	if (int rc = ensure_staging(h, bytes)) return rc;
	int ids[19];
	for (int c = 0; c < comps; c++) ids[c] = c;
	const RectCopy R = rect_desc(h, origin, size, true, ids, ids, comps, field_stride);
	launch_rect_bytes(h, elem, field, h->staging, R, h->compute);
	CUDA_TRY(h, cudaGetLastError());
	CUDA_TRY(h, cudaMemcpyAsync(host_dst, h->staging, bytes, cudaMemcpyDeviceToHost, h->compute));
	CUDA_TRY(h, cudaStreamSynchronize(h->compute));
	return LBM_OK;
}

int set_field(lbm_t h, void *field, size_t elem, int comps, const void *host_src,
		const int origin[3], const int size[3], const int *keep /* per component or NULL */,
		long long field_stride = 0)
{
	if (field_stride == 0) field_stride = h->n;
	if (int rc = use_device(h)) return rc;
	if (!host_src) return fail(h, LBM_ERR_INVALID, "null host pointer");
	if (!origin || !size) {
		if (field_stride == h->n)
			CUDA_TRY(h, cudaMemcpyAsync(field, host_src, (size_t)comps * h->n * elem, cudaMemcpyHostToDevice, h->compute));
		else for (int c = 0; c < comps; c++)
			CUDA_TRY(h, cudaMemcpyAsync((char *)field + (size_t)c * field_stride * elem, (const char *)host_src + (size_t)c * h->n * elem,
					(size_t)h->n * elem, cudaMemcpyHostToDevice, h->compute));
		CUDA_TRY(h, cudaStreamSynchronize(h->compute));
		return LBM_OK;
	}
	if (int rc = check_rect(h, origin, size)) return rc;
	const size_t bytes = (size_t)comps * size[0] * size[1] * size[2] * elem;
	if (int rc = ensure_staging(h, bytes)) return rc;
	CUDA_TRY(h, cudaMemcpyAsync(h->staging, host_src, bytes, cudaMemcpyHostToDevice, h->compute));
	CUDA_TRY(h, cudaEventRecord(h->ev_h2d, h->compute));
	int ids[19], nsel = 0;
	for (int c = 0; c < comps; c++) if (!keep || keep[c]) ids[nsel++] = c;
	if (nsel > 0) {
		const RectCopy R = rect_desc(h, origin, size, false, ids, ids, nsel, field_stride);
		launch_rect_bytes(h, elem, h->staging, field, R, h->compute);
		CUDA_TRY(h, cudaGetLastError());
	}
	/* the host buffer is consumed before return (CL_MEM_COPY_HOST_PTR semantics): wait for the
	 * upload only -- the scatter kernel stays queued ahead of whatever the caller launches next */
	CUDA_TRY(h, cudaEventSynchronize(h->ev_h2d));
	return LBM_OK;
}

const int kUnits[19][3] = {
	{ 1, 0, 0 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, -1, 0 },
	{ 1, 1, 0 }, { -1, -1, 0 }, { 1, -1, 0 }, { -1, 1, 0 },
	{ 1, 0, 1 }, { -1, 0, -1 }, { 1, 0, -1 }, { -1, 0, 1 },
	{ 0, 1, 1 }, { 0, -1, -1 }, { 0, 1, -1 }, { 0, -1, 1 },
	{ 0, 0, 1 }, { 0, 0, -1 }, { 0, 0, 0 } };

uint32_t popcount19(uint32_t m) { uint32_t c = 0; for (int f = 0; f < 19; f++) c += (m >> f) & 1; return c; }

} // namespace

/* ====================================================================== C ABI */
extern "C" {

int lbmGetVersion(void) { return 100; }

const char *lbmGetLastErrorString(lbm_t h) { return h ? h->error.c_str() : g_last_error.c_str(); }

int lbmGetDeviceCount(int *count)
{
	if (!count) return fail(NULL, LBM_ERR_INVALID, "null count");
	int c = 0;
	cudaError_t e = cudaGetDeviceCount(&c);
	if (e != cudaSuccess) { *count = 0; return fail(NULL, LBM_ERR_NO_DEVICE, cudaGetErrorString(e)); }
	*count = c;
	return LBM_OK;
}

int lbmCreate(lbm_t *out, const lbm_desc *d)
{
	if (!out || !d) return fail(NULL, LBM_ERR_INVALID, "null argument");
	*out = NULL;
	if (d->struct_size != sizeof(lbm_desc)) return fail(NULL, LBM_ERR_INVALID, "lbm_desc.struct_size mismatch");
	if (d->dtype != LBM_F32 && d->dtype != LBM_F64) return fail(NULL, LBM_ERR_INVALID, "unsupported class type T");
	for (int a = 0; a < 3; a++) if (d->size[a] < 1) return fail(NULL, LBM_ERR_INVALID, "domain size must be positive");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
		return fail(NULL, LBM_ERR_NO_DEVICE, "no CUDA device available (liblbm_b200 has no CPU fallback)");
	if (d->device < 0 || d->device >= ndev) return fail(NULL, LBM_ERR_INVALID, "invalid device number");

	lbm_solver *h = new lbm_solver();
	h->desc = *d;
	h->device = d->device; h->dtype = d->dtype;
	h->sx = d->size[0]; h->sy = d->size[1]; h->sz = d->size[2];
	h->n = (long long)h->sx * h->sy * h->sz;
	{
		/* slot stride: dense by default.  Padding the stride (to keep the 19 slot streams of a
		 * warp off one L2 slice / HBM channel group) was measured on B200 at 256^3 and 512^3
		 * fp32 and 384^3 fp64: 0 B is best, 4 KiB..1 MiB of padding cost 1-2 %
		 * (profiles/r1_slot_pad_sweep.md) -- the address hash already spreads power-of-two
		 * strides.  The knob stays as a tuning hook. */
		long long pad_bytes = 0;
		if (const char *e = getenv("LBM_B200_SLOT_PAD_BYTES")) pad_bytes = atoll(e);
		pad_bytes = (pad_bytes + 255) / 256 * 256;
		h->stride = h->n + pad_bytes / (d->dtype == LBM_F32 ? 4 : 8);
	}
	h->elem = d->dtype == LBM_F32 ? 4 : 8;
	h->dd = h->velocity = h->density = NULL; h->flags = NULL;
	h->staging = NULL; h->staging_bytes = 0; h->d_checksum = NULL;
	h->counter = 0; h->launches = 0; h->step_aux = NULL;
	h->profile_mode = 0; h->prof_base = NULL; h->prof_dropped = 0;
	h->xshell = 32;
	if (const char *e = getenv("LBM_B200_XSHELL")) h->xshell = atoi(e) > 1 ? atoi(e) : 2;
	h->x_in_kernel = 0; h->x_pending = 0; h->d_error = NULL;
	h->xfuse = 1;
	if (const char *e = getenv("LBM_B200_XFUSE")) h->xfuse = atoi(e);      /* 0 off, 1 default, 2 also rows no block size divides */
	h->wait_timeout_ns = 30ull * 1000000000ull;
	if (const char *e = getenv("LBM_B200_WAIT_TIMEOUT_MS")) if (atoll(e) > 0) h->wait_timeout_ns = (unsigned long long)atoll(e) * 1000000ull;
	h->axis_order = LBM_AXIS_ORDER_XYZ;
	if (const char *e = getenv("LBM_B200_AXIS_ORDER")) if (!strcmp(e, "zyx") || !strcmp(e, "ZYX")) h->axis_order = LBM_AXIS_ORDER_ZYX;
	h->smag = d->smagorinsky_cs != 0.0;
	h->u_lid = d->u_lid;
	const int maxvec = d->dtype == LBM_F32 ? 4 : 2;
	/* default: 2 cells per thread (measured best on B200 for fp32: 80-96 registers -> 20-24
	 * resident warps per SM, profiles/r1_sweep_launch_config_f32.jsonl, r1_summary.md); 4 is
	 * available on request */
	int vec = d->vector_width > 0 ? d->vector_width : 2;
	if (vec > maxvec) vec = maxvec;
	while (vec > 1 && (h->sx % vec) != 0) vec >>= 1;
	if (vec != 1 && vec != 2 && vec != 4) vec = 1;
	h->vec = vec;
	h->block = d->block_size > 0 ? d->block_size : 128;
	if (h->block % 32 != 0 || h->block > 1024) { delete h; return fail(NULL, LBM_ERR_INVALID, "block_size must be a multiple of 32 and <= 1024"); }
	const int wg = d->work_group_size;
	h->wg_quirk = (wg > 0 && d->beta_order == LBM_BETA_ORDER_SHIPPED && (wg % h->sx) == 0 && (h->n % wg) == 0) ? wg : 0;

#define CREATE_TRY(expr)                                                                     \
	do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) {                                  \
		std::string m = std::string(#expr) + ": " + cudaGetErrorString(e__);                  \
		lbmDestroy(h); return fail(NULL, LBM_ERR_CUDA, m); } } while (0)
	CREATE_TRY(cudaSetDevice(h->device));
	h->own_compute = d->compute_stream == NULL; h->own_comm = d->comm_stream == NULL;
	h->compute = (cudaStream_t)d->compute_stream; h->comm = (cudaStream_t)d->comm_stream;
	if (h->own_compute) CREATE_TRY(cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking));
	if (h->own_comm) {
		int prio_lo = 0, prio_hi = 0;
		CREATE_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		CREATE_TRY(cudaStreamCreateWithPriority(&h->comm, cudaStreamNonBlocking, prio_hi));
	}
	CREATE_TRY(cudaEventCreateWithFlags(&h->ev_compute, cudaEventDisableTiming));
	CREATE_TRY(cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming));
	CREATE_TRY(cudaEventCreateWithFlags(&h->ev_h2d, cudaEventDisableTiming));
	CREATE_TRY(cudaEventCreate(&h->ev_t0));
	CREATE_TRY(cudaEventCreate(&h->ev_t1));
	CREATE_TRY(cudaMalloc(&h->dd, (size_t)19 * h->stride * h->elem));
	CREATE_TRY(cudaMemsetAsync(h->dd, 0, (size_t)19 * h->stride * h->elem, h->compute));
	CREATE_TRY(cudaMalloc((void **)&h->flags, (size_t)h->n * sizeof(int)));
	CREATE_TRY(cudaMalloc((void **)&h->d_checksum, sizeof(double)));
	CREATE_TRY(cudaMalloc((void **)&h->d_error, sizeof(unsigned int)));
	CREATE_TRY(cudaMemsetAsync(h->d_error, 0, sizeof(unsigned int), h->compute));
	if (d->store_velocity) { CREATE_TRY(cudaMalloc(&h->velocity, (size_t)3 * h->n * h->elem)); }
	if (d->store_density) { CREATE_TRY(cudaMalloc(&h->density, (size_t)h->n * h->elem)); }
#undef CREATE_TRY
	*out = h;
	/* LBM_B200_PROFILE=1|2|3: the reference's compile-time PROFILE switch (SConstruct) as a run-time
	 * one, on from the first launch so that init_kernel is on the timeline too */
	if (const char *e = getenv("LBM_B200_PROFILE")) if (atoi(e) > 0) lbmProfileEnable(h, atoi(e) & 3);
	int rc = lbmReset(h);
	if (rc != LBM_OK) { std::string m = h->error; lbmDestroy(h); *out = NULL; return fail(NULL, rc, m); }
	return LBM_OK;
}

int lbmDestroy(lbm_t h)
{
	if (!h) return LBM_OK;
	cudaSetDevice(h->device);
	if (h->compute) cudaStreamSynchronize(h->compute);
	if (h->comm) cudaStreamSynchronize(h->comm);
	for (size_t i = 0; i < h->faces.size(); i++) {
		lbm_face &f = h->faces[i];
		if (f.peer_is_ipc && f.peer_block) cudaIpcCloseMemHandle(f.peer_block);
		cudaFree(f.local_block); cudaFree(f.counters);
	}
	cudaFree(h->d_error);
	cudaFree(h->dd); cudaFree(h->flags); cudaFree(h->velocity); cudaFree(h->density);
	cudaFree(h->staging); cudaFree(h->d_checksum);
	if (h->ev_compute) cudaEventDestroy(h->ev_compute);
	if (h->ev_comm) cudaEventDestroy(h->ev_comm);
	if (h->ev_h2d) cudaEventDestroy(h->ev_h2d);
	if (h->ev_t0) cudaEventDestroy(h->ev_t0);
	if (h->ev_t1) cudaEventDestroy(h->ev_t1);
	for (size_t i = 0; i < h->prof.size(); i++) { cudaEventDestroy(h->prof[i].e0); cudaEventDestroy(h->prof[i].e1); }
	for (size_t i = 0; i < h->prof_pool.size(); i++) cudaEventDestroy(h->prof_pool[i]);
	if (h->prof_base) cudaEventDestroy(h->prof_base);
	if (h->own_compute && h->compute) cudaStreamDestroy(h->compute);
	if (h->own_comm && h->comm) cudaStreamDestroy(h->comm);
	delete h;
	return LBM_OK;
}

int lbmReset(lbm_t h)
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	h->counter = 0;
	h->x_in_kernel = 0; h->x_pending = 0;
	const int block = 256;
	const unsigned grid = (unsigned)((h->n + block - 1) / block);
	const int *bc = h->desc.bc;
	{
	LaunchScope ls(h, "init_kernel", h->compute);
	if (h->dtype == LBM_F32)
		lbm_init_kernel<float><<<grid, block, 0, h->compute>>>((float *)h->dd, h->flags, (float *)h->velocity,
				(float *)h->density, h->n, h->stride, h->sx, h->sy, h->sz, bc[0], bc[1], bc[2], bc[3], bc[4], bc[5],
				h->velocity != NULL, h->density != NULL);
	else
		lbm_init_kernel<double><<<grid, block, 0, h->compute>>>((double *)h->dd, h->flags, (double *)h->velocity,
				(double *)h->density, h->n, h->stride, h->sx, h->sy, h->sz, bc[0], bc[1], bc[2], bc[3], bc[4], bc[5],
				h->velocity != NULL, h->density != NULL);
	}
	CUDA_TRY(h, cudaGetLastError());
	return LBM_OK;
}

int lbmStepAlpha(lbm_t h)
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	return launch_step(h, true, full_box(h), h->compute);
}

int lbmStepBeta(lbm_t h)
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	if (h->wg_quirk > 0 || h->sz < 8)
		return launch_step(h, false, full_box(h), h->compute);
	/* the wrapping kernel (planes at the array ends) runs on the comm stream NEXT TO the
	 * vectorised kernel instead of behind it: fork, two launches, join */
	if (int rc = lbmStreamWaitStream(h, 1)) return rc;
	if (int rc = launch_step(h, false, full_box(h), h->compute, h->comm)) return rc;
	return lbmStreamWaitStream(h, 0);
}

int lbmStep(lbm_t h)
{
	CHECK_HANDLE(h);
	/* CLbmSolver::simulationStep, src/CLbmSolver.hpp:664-676 */
	int rc = (h->counter & 1) ? lbmStepAlpha(h) : lbmStepBeta(h);
	h->x_in_kernel = 0;
	if (rc == LBM_OK) h->counter++;
	return rc;
}

int lbmSteps(lbm_t h, int nsteps)
{
	CHECK_HANDLE(h);
	for (int i = 0; i < nsteps; i++)
		if (int rc = lbmStep(h)) return rc;
	return LBM_OK;
}

static int step_shell_on(lbm_t h, int ghost_faces, cudaStream_t s)
{
	if (int rc = use_device(h)) return rc;
	std::vector<Box> shell; Box interior;
	partition(h, ghost_faces, shell, interior);
	const bool alpha = (h->counter & 1) != 0;
	/* x faces that are not split off (z,y,x order) leave the step kernels themselves */
	const bool xpush = (ghost_faces & 3) == 0 && x_fusable(h);
	for (size_t i = 0; i < shell.size(); i++)
		if (int rc = launch_step(h, alpha, shell[i], s, NULL, xpush)) return rc;
	return LBM_OK;
}

int lbmStepShell(lbm_t h, int ghost_faces)
{
	CHECK_HANDLE(h);
	return step_shell_on(h, ghost_faces, h->compute);
}

int lbmStepShellComm(lbm_t h, int ghost_faces)
{
	CHECK_HANDLE(h);
	return step_shell_on(h, ghost_faces, h->comm);
}

int lbmStepInterior(lbm_t h, int ghost_faces)
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	std::vector<Box> shell; Box interior;
	partition(h, ghost_faces, shell, interior);
	const bool alpha = (h->counter & 1) != 0;
	const bool xpush = (ghost_faces & 3) == 0 && x_fusable(h);
	if (int rc = launch_step(h, alpha, interior, h->compute, h->step_aux, xpush)) return rc;
	h->x_in_kernel = xpush ? 1 + (alpha ? LBM_SYNC_ALPHA : LBM_SYNC_BETA) : 0;
	h->x_pending = 0;            /* consumed by the shell + interior kernels just launched (or flushed) */
	h->counter++;
	return LBM_OK;
}

int lbmStreamWaitStream(lbm_t h, int waiter_is_comm)
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	if (waiter_is_comm) {
		CUDA_TRY(h, cudaEventRecord(h->ev_compute, h->compute));
		CUDA_TRY(h, cudaStreamWaitEvent(h->comm, h->ev_compute, 0));
	} else {
		CUDA_TRY(h, cudaEventRecord(h->ev_comm, h->comm));
		CUDA_TRY(h, cudaStreamWaitEvent(h->compute, h->ev_comm, 0));
	}
	return LBM_OK;
}

int lbmGetStreams(lbm_t h, void **compute_stream, void **comm_stream)
{
	CHECK_HANDLE(h);
	if (compute_stream) *compute_stream = (void *)h->compute;
	if (comm_stream) *comm_stream = (void *)h->comm;
	return LBM_OK;
}

int lbmWait(lbm_t h)
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	CUDA_TRY(h, cudaStreamSynchronize(h->compute));
	CUDA_TRY(h, cudaStreamSynchronize(h->comm));
	if (!h->faces.empty()) {
		unsigned int err = 0;
		CUDA_TRY(h, cudaMemcpy(&err, h->d_error, sizeof(err), cudaMemcpyDeviceToHost));
		if (err) {
			CUDA_TRY(h, cudaMemset(h->d_error, 0, sizeof(err)));
			return fail(h, LBM_ERR_TIMEOUT, "halo wait timed out: a neighbour never pushed its face (LBM_B200_WAIT_TIMEOUT_MS)");
		}
	}
	return LBM_OK;
}

int lbmGetStepCounter(lbm_t h, uint64_t *counter)
{
	CHECK_HANDLE(h);
	if (!counter) return fail(h, LBM_ERR_INVALID, "null counter");
	*counter = h->counter;
	return LBM_OK;
}

int lbmSetStepCounter(lbm_t h, uint64_t counter) { CHECK_HANDLE(h); h->counter = counter; return LBM_OK; }

int lbmSetDrivenCavityVelocity(lbm_t h, double u_lid) { CHECK_HANDLE(h); h->u_lid = u_lid; return LBM_OK; }

int lbmStoreDD(lbm_t h, void *host_dst, const int origin[3], const int size[3])
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	if (int rc = flush_x_pending(h, h->compute)) return rc;
	return store_field(h, h->dd, h->elem, 19, host_dst, origin, size, h->stride);
}

int lbmSetDD(lbm_t h, const void *host_src, const int origin[3], const int size[3], const int norm[3])
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	if (int rc = flush_x_pending(h, h->compute)) return rc;
	int keep[19];
	for (int f = 0; f < 19; f++)   /* src/CLbmSolver.hpp:748: norm.dotProd(lbm_units[f]) > 0 */
		keep[f] = !norm || (norm[0] * kUnits[f][0] + norm[1] * kUnits[f][1] + norm[2] * kUnits[f][2] > 0);
	return set_field(h, h->dd, h->elem, 19, host_src, origin, size, keep, h->stride);
}

int lbmStoreVelocity(lbm_t h, void *host_dst, const int origin[3], const int size[3])
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	if (int rc = ensure_field(h, &h->velocity, 3)) return rc;
	return store_field(h, h->velocity, h->elem, 3, host_dst, origin, size);
}

int lbmSetVelocity(lbm_t h, const void *host_src, const int origin[3], const int size[3])
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	if (int rc = ensure_field(h, &h->velocity, 3)) return rc;
	return set_field(h, h->velocity, h->elem, 3, host_src, origin, size, NULL);
}

int lbmStoreDensity(lbm_t h, void *host_dst, const int origin[3], const int size[3])
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	if (int rc = ensure_field(h, &h->density, 1)) return rc;
	return store_field(h, h->density, h->elem, 1, host_dst, origin, size);
}

int lbmSetDensity(lbm_t h, const void *host_src, const int origin[3], const int size[3])
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	if (int rc = ensure_field(h, &h->density, 1)) return rc;
	return set_field(h, h->density, h->elem, 1, host_src, origin, size, NULL);
}

int lbmStoreFlags(lbm_t h, int *host_dst, const int origin[3], const int size[3])
{
	CHECK_HANDLE(h);
	return store_field(h, h->flags, sizeof(int), 1, host_dst, origin, size);
}

int lbmSetFlags(lbm_t h, const int *host_src, const int origin[3], const int size[3])
{
	CHECK_HANDLE(h);
	return set_field(h, h->flags, sizeof(int), 1, host_src, origin, size, NULL);
}

int lbmChecksumVelocity(lbm_t h, double *out, int host_order)
{
	CHECK_HANDLE(h);
	if (!out) return fail(h, LBM_ERR_INVALID, "null out");
	if (int rc = use_device(h)) return rc;
	if (int rc = ensure_field(h, &h->velocity, 3)) return rc;
	if (host_order) {
		/* src/CLbmSolver.hpp:1103-1123 verbatim semantics: float accumulator, index order */
		std::vector<char> vel((size_t)3 * h->n * h->elem);
		std::vector<int> fl((size_t)h->n);
		if (int rc = store_field(h, h->velocity, h->elem, 3, vel.data(), NULL, NULL)) return rc;
		if (int rc = store_field(h, h->flags, sizeof(int), 1, fl.data(), NULL, NULL)) return rc;
		float checksum = 0;
		if (h->dtype == LBM_F32) {
			const float *v = (const float *)vel.data();
			for (long long a = 0; a < h->n; a++)
				if (fl[a] == LBM_FLAG_FLUID) checksum += v[a] + v[h->n + a] + v[2 * h->n + a];
		} else {
			const double *v = (const double *)vel.data();
			for (long long a = 0; a < h->n; a++)
				if (fl[a] == LBM_FLAG_FLUID) checksum += v[a] + v[h->n + a] + v[2 * h->n + a];
		}
		*out = checksum;
		return LBM_OK;
	}
	CUDA_TRY(h, cudaMemsetAsync(h->d_checksum, 0, sizeof(double), h->compute));
	const int block = 256, grid = 148 * 8;
	{
	LaunchScope ls(h, "checksum_kernel", h->compute);
	if (h->dtype == LBM_F32)
		checksum_kernel<float><<<grid, block, 0, h->compute>>>((const float *)h->velocity, h->flags, h->n, h->d_checksum);
	else
		checksum_kernel<double><<<grid, block, 0, h->compute>>>((const double *)h->velocity, h->flags, h->n, h->d_checksum);
	}
	CUDA_TRY(h, cudaGetLastError());
	CUDA_TRY(h, cudaMemcpyAsync(out, h->d_checksum, sizeof(double), cudaMemcpyDeviceToHost, h->compute));
	CUDA_TRY(h, cudaStreamSynchronize(h->compute));
	return LBM_OK;
}

/* ---------------------------------------------------------------- halo path */
int lbmHaloSlotMask(int sync_kind, const int recv_dir[3], int slots, uint32_t *mask)
{
	if (!mask || !recv_dir) return fail(NULL, LBM_ERR_INVALID, "null argument");
	uint32_t m = 0;
	for (int f = 0; f < 19; f++) {
		const int dot = recv_dir[0] * kUnits[f][0] + recv_dir[1] * kUnits[f][1] + recv_dir[2] * kUnits[f][2];
		bool keep;
		if (sync_kind == LBM_SYNC_BETA) keep = dot > 0;            /* src/CLbmSolver.hpp:748 */
		else keep = (slots == LBM_HALO_SLOTS_MINIMAL) ? (dot < 0) : true;
		if (keep) m |= 1u << f;
	}
	*mask = m;
	return LBM_OK;
}

int lbmHaloBytes(lbm_t h, const int size[3], uint32_t slot_mask, size_t *bytes)
{
	CHECK_HANDLE(h);
	if (!size || !bytes) return fail(h, LBM_ERR_INVALID, "null argument");
	*bytes = (size_t)popcount19(slot_mask) * size[0] * size[1] * size[2] * h->elem;
	return LBM_OK;
}

int lbmHaloPack(lbm_t h, const int origin[3], const int size[3], uint32_t slot_mask, void *dev_buf, void *stream)
{
	CHECK_HANDLE(h);
	if (!origin || !size || !dev_buf) return fail(h, LBM_ERR_INVALID, "null argument");
	if (int rc = use_device(h)) return rc;
	if (int rc = check_rect(h, origin, size)) return rc;
	if (int rc = flush_x_pending(h, stream ? (cudaStream_t)stream : h->comm)) return rc;
	int field[19], packed[19], n = 0;
	for (int f = 0; f < 19; f++) if ((slot_mask >> f) & 1) { field[n] = f; packed[n] = n; n++; }
	if (n == 0) return LBM_OK;
	const RectCopy R = rect_desc(h, origin, size, true, field, packed, n, h->stride);
	launch_rect_bytes(h, h->elem, h->dd, dev_buf, R, stream ? (cudaStream_t)stream : h->comm);
	CUDA_TRY(h, cudaGetLastError());
	return LBM_OK;
}

int lbmHaloUnpack(lbm_t h, const int origin[3], const int size[3], uint32_t buf_slot_mask, uint32_t write_mask,
		const void *dev_buf, void *stream)
{
	CHECK_HANDLE(h);
	if (!origin || !size || !dev_buf) return fail(h, LBM_ERR_INVALID, "null argument");
	if (write_mask & ~buf_slot_mask) return fail(h, LBM_ERR_INVALID, "write_mask selects slots the buffer does not hold");
	if (int rc = use_device(h)) return rc;
	if (int rc = check_rect(h, origin, size)) return rc;
	if (int rc = flush_x_pending(h, stream ? (cudaStream_t)stream : h->comm)) return rc;
	int field[19], packed[19], n = 0, pos = 0;
	for (int f = 0; f < 19; f++) {
		if (!((buf_slot_mask >> f) & 1)) continue;
		if ((write_mask >> f) & 1) { field[n] = f; packed[n] = pos; n++; }
		pos++;
	}
	if (n == 0) return LBM_OK;
	const RectCopy R = rect_desc(h, origin, size, false, field, packed, n, h->stride);
	launch_rect_bytes(h, h->elem, dev_buf, h->dd, R, stream ? (cudaStream_t)stream : h->comm);
	CUDA_TRY(h, cudaGetLastError());
	return LBM_OK;
}

int lbmHaloCopyPeer(lbm_t src, const int src_origin[3], lbm_t dst, const int dst_origin[3],
		const int size[3], uint32_t slot_mask, void *stream)
{
	CHECK_HANDLE(src); CHECK_HANDLE(dst);
	if (!src_origin || !dst_origin || !size) return fail(src, LBM_ERR_INVALID, "null argument");
	if (src->dtype != dst->dtype) return fail(src, LBM_ERR_INVALID, "peer dtype mismatch");
	if (int rc = use_device(src)) return rc;
	if (int rc = check_rect(src, src_origin, size)) return rc;
	if (int rc = check_rect(dst, dst_origin, size)) return fail(src, rc, dst->error);
	if (int rc = flush_x_pending(src, stream ? (cudaStream_t)stream : src->comm)) return rc;
	if (dst->x_pending) { if (int rc = use_device(dst)) return rc; if (int rc = flush_x_pending(dst, dst->compute)) return rc; if (int rc = use_device(src)) return rc; }
	if (src->device != dst->device) {
		int can = 0;
		CUDA_TRY(src, cudaDeviceCanAccessPeer(&can, src->device, dst->device));
		if (!can) return fail(src, LBM_ERR_CUDA, "devices are not NVLink/PCIe peers");
		cudaError_t e = cudaDeviceEnablePeerAccess(dst->device, 0);
		if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
			return fail(src, LBM_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
		cudaGetLastError();
	}
	RectCopy R;
	const int S[3] = { src->sx, src->sy, src->sz }, D[3] = { dst->sx, dst->sy, dst->sz };
	for (int a = 0; a < 3; a++) {
		R.block[a] = size[a]; R.so[a] = src_origin[a]; R.ss[a] = S[a]; R.dorg[a] = dst_origin[a]; R.ds[a] = D[a];
	}
	R.src_comp_stride = src->stride; R.dst_comp_stride = dst->stride;
	int n = 0;
	for (int f = 0; f < 19; f++) if ((slot_mask >> f) & 1) { R.src_comp[n] = f; R.dst_comp[n] = f; n++; }
	R.ncomp = n;
	if (n == 0) return LBM_OK;
	launch_rect_bytes(src, src->elem, src->dd, dst->dd, R, stream ? (cudaStream_t)stream : src->comm);
	CUDA_TRY(src, cudaGetLastError());
	return LBM_OK;
}

/* ---------------------------------------------------------------- peer-memory halo exchange */
} /* extern "C" */
namespace {

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

uint32_t slot_mask(int kind, const int dir[3], int slots)
{
	uint32_t m = 0;
	lbmHaloSlotMask(kind, dir, slots, &m);
	return m;
}

int face_check(lbm_t h, int face_id)
{
	if (face_id < 0 || face_id >= (int)h->faces.size()) return fail(h, LBM_ERR_INVALID, "invalid face id");
	return LBM_OK;
}

/* widest vector (in elements, <= 16 B) that every row of the face can be moved with */
int face_vec(lbm_t h, const int origin[3], const int size[3])
{
	int v = (int)(16 / h->elem);
	while (v > 1 && ((size[0] % v) || (origin[0] % v) || (h->sx % v) || (h->stride % v))) v >>= 1;
	return v;
}

void fill_face(lbm_t h, HaloFace &F, const int origin[3], const int size[3])
{
	F.dd = h->dd; F.dd_stride = h->stride;
	for (int a = 0; a < 3; a++) { F.origin[a] = origin[a]; F.size[a] = size[a]; }
	F.ss[0] = h->sx; F.ss[1] = h->sy;
	F.vec = face_vec(h, origin, size);
}

/* hidden = the kernel shares the SMs with the step kernel (few blocks: every push block ends with a
 * system fence); exposed = it runs alone behind the step kernel (x faces): use the whole machine */
unsigned face_blocks(const HaloFace &F, bool exposed)
{
	const long long work = ((long long)F.size[0] * F.size[1] * F.size[2] * F.ncomp) / F.vec;
	long long g = (work + 255) / 256;
	const long long cap = exposed ? 148 * 8 : 148;
	if (g > cap) g = cap;
	return (unsigned)(g < 1 ? 1 : g);
}

/* descriptors of the faces of one axis, push side (my dd rect -> the neighbour's receive block) */
int build_push(lbm_t h, int kind, int axis, HaloAxis &A, unsigned &nf, unsigned &blocks, bool exposed)
{
	memset(&A, 0, sizeof(A));
	nf = 0; blocks = 1;
	for (size_t i = 0; i < h->faces.size(); i++) {
		lbm_face &f = h->faces[i];
		if (f.axis != axis) continue;
		if (!f.connected) return fail(h, LBM_ERR_INVALID, "halo face is not connected to its peer");
		if (nf == 2) return fail(h, LBM_ERR_INVALID, "more than two halo faces on one axis");
		HaloFace &F = A.f[nf];
		fill_face(h, F, kind == LBM_SYNC_BETA ? f.recv_origin : f.send_origin, f.size);
		int n = 0;
		for (int k = 0; k < 19; k++) if ((f.send_mask[kind] >> k) & 1) { F.dd_comp[n] = k; F.st_comp[n] = n; n++; }
		F.ncomp = n;
		F.staging = f.peer_block + f.peer_stage_off[kind];
		F.flag = (volatile unsigned int *)(f.peer_block + 64 * kind);
		F.sync_count = f.counters + kind;
		F.block_counter = f.counters + 4;
		const unsigned b = face_blocks(F, exposed);
		if (b > blocks) blocks = b;
		nf++;
	}
	return LBM_OK;
}

/* ... pull side (my receive block -> my dd rect) */
int build_pull(lbm_t h, int kind, int axis, HaloAxis &A, unsigned &nf, unsigned &blocks, bool exposed)
{
	memset(&A, 0, sizeof(A));
	nf = 0; blocks = 1;
	for (size_t i = 0; i < h->faces.size(); i++) {
		lbm_face &f = h->faces[i];
		if (f.axis != axis) continue;
		if (nf == 2) return fail(h, LBM_ERR_INVALID, "more than two halo faces on one axis");
		HaloFace &F = A.f[nf];
		fill_face(h, F, kind == LBM_SYNC_BETA ? f.send_origin : f.recv_origin, f.size);
		int n = 0, pos = 0;
		for (int k = 0; k < 19; k++) {
			if (!((f.recv_mask[kind] >> k) & 1)) continue;
			if ((f.write_mask[kind] >> k) & 1) { F.dd_comp[n] = k; F.st_comp[n] = pos; n++; }
			pos++;
		}
		F.ncomp = n;
		F.staging = f.local_block + f.stage_off[kind];
		F.flag = (volatile unsigned int *)(f.local_block + 64 * kind);
		F.sync_count = f.counters + 2 + kind;
		F.block_counter = NULL;
		const unsigned b = face_blocks(F, exposed);
		if (b > blocks) blocks = b;
		nf++;
	}
	return LBM_OK;
}

int axis_unpack(lbm_t h, int kind, int axis, cudaStream_t s, bool exposed)
{
	HaloAxis A; unsigned nf, blocks;
	if (int rc = build_pull(h, kind, axis, A, nf, blocks, exposed)) return rc;
	if (nf == 0) return LBM_OK;
	dim3 grid(blocks, nf);
	LaunchScope ls(h, "halo_pull", s);
	if (h->dtype == LBM_F32) halo_unpack_kernel<float><<<grid, 256, 0, s>>>(A);
	else halo_unpack_kernel<double><<<grid, 256, 0, s>>>(A);
	CUDA_TRY(h, cudaGetLastError());
	return LBM_OK;
}

/* an x face that was received but left in its receive block (lazy pull) is scattered into dd now */
int flush_x_pending(lbm_t h, cudaStream_t s)
{
	if (!h->x_pending) return LBM_OK;
	const int kind = h->x_pending - 1;
	h->x_pending = 0;
	return axis_unpack(h, kind, 0, s, true);
}

/* every face of one axis in ONE launch.  rim_only: the bulk of the (x) faces already left the step
 * kernels (XFUSE); send the rim lines and raise the flags -- and, with_wait, wait for the neighbour's
 * flags in the same launch. */
int axis_push(lbm_t h, int kind, int axis, cudaStream_t s, bool exposed, bool rim_only, bool with_wait = false)
{
	HaloAxis A; unsigned nf, blocks;
	if (int rc = build_push(h, kind, axis, A, nf, blocks, exposed)) return rc;
	if (nf == 0) return LBM_OK;
	if (!rim_only) {
		if (int rc = flush_x_pending(h, s)) return rc;      /* the face is read out of dd */
		dim3 grid(blocks, nf);
		LaunchScope ls(h, "halo_push", s);
		if (h->dtype == LBM_F32) halo_push_kernel<float><<<grid, 256, 0, s>>>(A);
		else halo_push_kernel<double><<<grid, 256, 0, s>>>(A);
	} else {
		HaloAxis W; unsigned nw = 0, wb;
		if (with_wait) { if (int rc = build_pull(h, kind, axis, W, nw, wb, exposed)) return rc; }
		else memset(&W, 0, sizeof(W));
		/* rim lines only exist where a y or z face has a neighbour (forwarded edges, ghost rims) */
		int do_rim = 0;
		for (size_t i = 0; i < h->faces.size(); i++) if (h->faces[i].axis != axis) do_rim = 1;
		/* one rim element per thread: 4 * (Sy + Sz) cells x 5 slots per face */
		int rim_blocks = 1;
		if (do_rim) {
			const long long elems = 4LL * (h->sy + h->sz) * 5;
			rim_blocks = (int)((elems + 255) / 256);
			if (rim_blocks > 128) rim_blocks = 128;
		}
		dim3 grid((unsigned)(rim_blocks + (nw > 0 ? 1 : 0)), nf);
		LaunchScope ls(h, "halo_xrim", s);
		if (h->dtype == LBM_F32) halo_xrim_flag_kernel<float><<<grid, 256, 0, s>>>(A, W, (int)nw, rim_blocks, do_rim, h->wait_timeout_ns, h->d_error);
		else halo_xrim_flag_kernel<double><<<grid, 256, 0, s>>>(A, W, (int)nw, rim_blocks, do_rim, h->wait_timeout_ns, h->d_error);
	}
	CUDA_TRY(h, cudaGetLastError());
	return LBM_OK;
}

/* wait (one thread per face), then unpack -- or, lazy (x faces of an XFUSE run), leave the data in the
 * receive block for the next step kernel to read */
int axis_pull(lbm_t h, int kind, int axis, cudaStream_t s, bool exposed, bool lazy = false)
{
	HaloAxis A; unsigned nf, blocks;
	if (int rc = build_pull(h, kind, axis, A, nf, blocks, exposed)) return rc;
	if (nf == 0) return LBM_OK;
	if (axis == 0) { if (int rc = flush_x_pending(h, s)) return rc; }
	{
	LaunchScope ls(h, "halo_wait", s);
	halo_wait_kernel<<<1, 32, 0, s>>>(A, (int)nf, h->wait_timeout_ns, h->d_error);
	}
	CUDA_TRY(h, cudaGetLastError());
	if (lazy) { h->x_pending = 1 + kind; return LBM_OK; }
	dim3 grid(blocks, nf);
	{
	LaunchScope ls(h, "halo_pull", s);
	if (h->dtype == LBM_F32) halo_unpack_kernel<float><<<grid, 256, 0, s>>>(A);
	else halo_unpack_kernel<double><<<grid, 256, 0, s>>>(A);
	}
	CUDA_TRY(h, cudaGetLastError());
	return LBM_OK;
}

int ghost_mask_of_faces(lbm_t h)
{
	int m = 0;
	for (size_t i = 0; i < h->faces.size(); i++) {
		const lbm_face &f = h->faces[i];
		m |= 1 << (2 * f.axis + (f.dir[f.axis] > 0 ? 0 : 1));
	}
	return m;
}

} // namespace
extern "C" {

int lbmCommAddFace(lbm_t h, int dst_rank, const int send_origin[3], const int recv_origin[3], const int size[3],
		const int dir[3], int slots, int *face_id)
{
	CHECK_HANDLE(h);
	if (!send_origin || !recv_origin || !size || !dir) return fail(h, LBM_ERR_INVALID, "null argument");
	if (int rc = use_device(h)) return rc;
	if (int rc = check_rect(h, send_origin, size)) return rc;
	if (int rc = check_rect(h, recv_origin, size)) return rc;
	lbm_face f;
	memset(&f, 0, sizeof(f));
	f.dst_rank = dst_rank;
	int nz = 0;
	for (int a = 0; a < 3; a++) {
		f.send_origin[a] = send_origin[a]; f.recv_origin[a] = recv_origin[a]; f.size[a] = size[a]; f.dir[a] = dir[a];
		if (dir[a] != 0) { f.axis = a; nz++; }
	}
	if (nz != 1) return fail(h, LBM_ERR_INVALID, "comm direction must be a signed unit vector");
	const int opp[3] = { -dir[0], -dir[1], -dir[2] };
	const size_t cells = (size_t)size[0] * size[1] * size[2];
	size_t off = 256, poff = 256;
	for (int k = 0; k < 2; k++) {
		f.recv_mask[k] = slot_mask(k, dir, slots);
		f.send_mask[k] = slot_mask(k, opp, slots);
		f.write_mask[k] = k == LBM_SYNC_BETA ? slot_mask(k, dir, LBM_HALO_SLOTS_MINIMAL) : f.recv_mask[k];
		f.stage_elems[k] = (size_t)popcount19(f.recv_mask[k]) * cells;
		f.stage_off[k] = off;
		off += align256(f.stage_elems[k] * h->elem);
		f.peer_stage_off[k] = poff;
		poff += align256((size_t)popcount19(f.send_mask[k]) * cells * h->elem);
	}
	f.local_bytes = off;
	CUDA_TRY(h, cudaMalloc((void **)&f.local_block, f.local_bytes));
	CUDA_TRY(h, cudaMemset(f.local_block, 0, 256));
	CUDA_TRY(h, cudaMalloc((void **)&f.counters, 8 * sizeof(unsigned int)));
	CUDA_TRY(h, cudaMemset(f.counters, 0, 8 * sizeof(unsigned int)));
	h->faces.push_back(f);
	if (face_id) *face_id = (int)h->faces.size() - 1;
	return LBM_OK;
}

int lbmCommFaceCount(lbm_t h, int *count)
{
	CHECK_HANDLE(h);
	if (!count) return fail(h, LBM_ERR_INVALID, "null count");
	*count = (int)h->faces.size();
	return LBM_OK;
}

int lbmCommGetIpcHandle(lbm_t h, int face_id, void *handle64)
{
	CHECK_HANDLE(h);
	if (int rc = face_check(h, face_id)) return rc;
	if (!handle64) return fail(h, LBM_ERR_INVALID, "null handle buffer");
	if (int rc = use_device(h)) return rc;
	cudaIpcMemHandle_t hd;
	CUDA_TRY(h, cudaIpcGetMemHandle(&hd, h->faces[face_id].local_block));
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
	memcpy(handle64, &hd, 64);
	return LBM_OK;
}

int lbmCommConnectIpc(lbm_t h, int face_id, const void *peer_handle64)
{
	CHECK_HANDLE(h);
	if (int rc = face_check(h, face_id)) return rc;
	if (!peer_handle64) return fail(h, LBM_ERR_INVALID, "null handle");
	if (int rc = use_device(h)) return rc;
	cudaIpcMemHandle_t hd;
	memcpy(&hd, peer_handle64, 64);
	void *p = NULL;
	CUDA_TRY(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
	lbm_face &f = h->faces[face_id];
	f.peer_block = (char *)p; f.peer_is_ipc = true; f.connected = true;
	return LBM_OK;
}

int lbmCommConnectLocal(lbm_t h, int face_id, lbm_t peer, int peer_face_id)
{
	CHECK_HANDLE(h); CHECK_HANDLE(peer);
	if (int rc = face_check(h, face_id)) return rc;
	if (peer_face_id < 0 || peer_face_id >= (int)peer->faces.size()) return fail(h, LBM_ERR_INVALID, "invalid peer face id");
	if (int rc = use_device(h)) return rc;
	if (h->device != peer->device) {
		int can = 0;
		CUDA_TRY(h, cudaDeviceCanAccessPeer(&can, h->device, peer->device));
		if (!can) return fail(h, LBM_ERR_CUDA, "devices are not NVLink/PCIe peers");
		cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
		if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
			return fail(h, LBM_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
		cudaGetLastError();
	}
	lbm_face &f = h->faces[face_id];
	const lbm_face &g = peer->faces[peer_face_id];
	for (int k = 0; k < 2; k++)
		if (f.send_mask[k] != g.recv_mask[k] || f.size[0] != g.size[0] || f.size[1] != g.size[1] || f.size[2] != g.size[2])
			return fail(h, LBM_ERR_INVALID, "peer face does not mirror this face");
	f.peer_block = g.local_block; f.peer_is_ipc = false; f.connected = true;
	return LBM_OK;
}

int lbmCommBeginSync(lbm_t h, int sync_kind)
{
	CHECK_HANDLE(h);
	if (sync_kind != LBM_SYNC_ALPHA && sync_kind != LBM_SYNC_BETA) return fail(h, LBM_ERR_INVALID, "bad sync kind");
	/* nothing to do: the sequence numbers are counted by the push / wait kernels in device memory */
	return LBM_OK;
}

int lbmCommPush(lbm_t h, int sync_kind, int axis)
{
	CHECK_HANDLE(h);
	if (sync_kind != LBM_SYNC_ALPHA && sync_kind != LBM_SYNC_BETA) return fail(h, LBM_ERR_INVALID, "bad sync kind");
	if (int rc = use_device(h)) return rc;
	const bool rim_only = axis == 0 && h->x_in_kernel == 1 + sync_kind;
	if (axis == 0) h->x_in_kernel = 0;
	return axis_push(h, sync_kind, axis, h->comm, false, rim_only);
}

int lbmCommPull(lbm_t h, int sync_kind, int axis)
{
	CHECK_HANDLE(h);
	if (sync_kind != LBM_SYNC_ALPHA && sync_kind != LBM_SYNC_BETA) return fail(h, LBM_ERR_INVALID, "bad sync kind");
	if (int rc = use_device(h)) return rc;
	return axis_pull(h, sync_kind, axis, h->comm, false, axis == 0 && x_fusable(h));
}

int lbmCommSetAxisOrder(lbm_t h, int order)
{
	CHECK_HANDLE(h);
	if (order != LBM_AXIS_ORDER_XYZ && order != LBM_AXIS_ORDER_ZYX) return fail(h, LBM_ERR_INVALID, "unknown axis order");
	h->axis_order = order;
	return LBM_OK;
}

int lbmCommGetAxisOrder(lbm_t h, int *order)
{
	CHECK_HANDLE(h);
	if (!order) return fail(h, LBM_ERR_INVALID, "null order");
	*order = h->axis_order;
	return LBM_OK;
}

int lbmCommSync(lbm_t h, int sync_kind)
{
	CHECK_HANDLE(h);
	if (int rc = lbmCommBeginSync(h, sync_kind)) return rc;
	/* one axis after the other: later axes carry the rims the earlier ones delivered
	 * (x, y, z = the reference's sequential CComm walk, src/CManager.hpp:122-199) */
	for (int i = 0; i < 3; i++) {
		const int axis = h->axis_order == LBM_AXIS_ORDER_ZYX ? 2 - i : i;
		if (int rc = lbmCommPush(h, sync_kind, axis)) return rc;
		if (int rc = lbmCommPull(h, sync_kind, axis)) return rc;
	}
	return LBM_OK;
}

static int comm_step(lbm_t h, cudaEvent_t *marks /* NULL or 5 timing events */)
{
	if (int rc = use_device(h)) return rc;
	const int faces = ghost_mask_of_faces(h);
	/* z,y,x order: x faces are not split off; they are exchanged after the step kernel */
	const bool x_last = h->axis_order == LBM_AXIS_ORDER_ZYX && (faces & 3) != 0;
	const int split = x_last ? (faces & ~3) : faces;
	const int kind = (h->counter & 1) ? LBM_SYNC_ALPHA : LBM_SYNC_BETA;   /* the sync that follows this step */
	/* fork: the (high-priority) comm stream runs shell -> push -> wait -> unpack while the
	 * compute stream runs the interior kernel; shell and interior touch disjoint
	 * (slot, location) pairs (A-A invariant), so they may run concurrently.  The shell is
	 * enqueued first so that its few blocks are scheduled ahead of the interior's. */
	if (marks) CUDA_TRY(h, cudaEventRecord(marks[0], h->compute));
	if (int rc = lbmStreamWaitStream(h, 1)) return rc;
	if (int rc = lbmStepShellComm(h, split)) return rc;
	if (marks) CUDA_TRY(h, cudaEventRecord(marks[1], h->comm));
	h->step_aux = h->comm;                                  /* the comm stream is forked: wrapping kernel there */
	int rc_int = lbmStepInterior(h, split);
	h->step_aux = NULL;
	if (rc_int) return rc_int;
	if (marks) CUDA_TRY(h, cudaEventRecord(marks[2], h->compute));
	if (!x_last) {
		if (int rc = lbmCommSync(h, kind)) return rc;
		if (marks) CUDA_TRY(h, cudaEventRecord(marks[3], h->comm));
		if (int rc = lbmStreamWaitStream(h, 0)) return rc;     /* join: the next step needs the halo */
	} else {
		for (int axis = 2; axis >= 1; axis--) {                 /* hidden under the interior kernel */
			if (int rc = axis_push(h, kind, axis, h->comm, false, false)) return rc;
			if (int rc = axis_pull(h, kind, axis, h->comm, false)) return rc;
		}
		/* join first: the x faces follow the step kernel AND the y/z unpack (whose rims they forward),
		 * on the compute stream itself -- no further stream hop between them and the next step.
		 * Their bulk is already in the neighbour's block (XPUSH step kernels): rim lines + flag, then
		 * wait + unpack with the whole machine. */
		if (int rc = lbmStreamWaitStream(h, 0)) return rc;
		const bool rim_only = h->x_in_kernel == 1 + kind;
		const bool lazy = x_fusable(h);
		h->x_in_kernel = 0;
		if (rim_only && lazy) {
			/* the whole exposed x tail is ONE small kernel: rim lines, flag, wait for the neighbour's flag;
			 * the received faces stay in their blocks for the next step kernels */
			if (int rc = axis_push(h, kind, 0, h->compute, true, true, true)) return rc;
			h->x_pending = 1 + kind;
		} else {
			if (int rc = axis_push(h, kind, 0, h->compute, true, rim_only)) return rc;
			if (int rc = axis_pull(h, kind, 0, h->compute, true, lazy)) return rc;
		}
		if (marks) CUDA_TRY(h, cudaEventRecord(marks[3], h->compute));
	}
	if (marks) CUDA_TRY(h, cudaEventRecord(marks[4], h->compute));
	return LBM_OK;
}

int lbmCommStep(lbm_t h)
{
	CHECK_HANDLE(h);
	if (h->faces.empty()) return lbmStep(h);
	return comm_step(h, NULL);
}

int lbmCommStepTimed(lbm_t h, float ms[4])
{
	CHECK_HANDLE(h);
	if (!ms) return fail(h, LBM_ERR_INVALID, "null ms");
	if (h->faces.empty()) return fail(h, LBM_ERR_INVALID, "no halo faces registered");
	if (int rc = use_device(h)) return rc;
	cudaEvent_t ev[5];
	for (int i = 0; i < 5; i++) CUDA_TRY(h, cudaEventCreate(&ev[i]));
	int rc = comm_step(h, ev);
	if (rc == LBM_OK) {
		cudaError_t e = cudaEventSynchronize(ev[4]);
		for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventElapsedTime(&ms[i], ev[0], ev[i + 1]);
		if (e != cudaSuccess) rc = fail(h, LBM_ERR_CUDA, cudaGetErrorString(e));
	}
	for (int i = 0; i < 5; i++) cudaEventDestroy(ev[i]);
	return rc;
}

int lbmGetDevicePointer(lbm_t h, int which, void **ptr, size_t *bytes)
{
	CHECK_HANDLE(h);
	if (!ptr) return fail(h, LBM_ERR_INVALID, "null ptr");
	size_t b = 0;
	switch (which) {
	case LBM_BUF_DD:
		if (int rc = use_device(h)) return rc;
		if (int rc = flush_x_pending(h, h->compute)) return rc;
		*ptr = h->dd; b = (size_t)19 * h->stride * h->elem; break;
	case LBM_BUF_FLAGS: *ptr = h->flags; b = (size_t)h->n * sizeof(int); break;
	case LBM_BUF_VELOCITY: *ptr = h->velocity; b = h->velocity ? (size_t)3 * h->n * h->elem : 0; break;
	case LBM_BUF_DENSITY: *ptr = h->density; b = h->density ? (size_t)h->n * h->elem : 0; break;
	default: return fail(h, LBM_ERR_INVALID, "unknown buffer id");
	}
	if (bytes) *bytes = b;
	return LBM_OK;
}

int lbmGetSlotStride(lbm_t h, size_t *cells)
{
	CHECK_HANDLE(h);
	if (!cells) return fail(h, LBM_ERR_INVALID, "null cells");
	*cells = (size_t)h->stride;
	return LBM_OK;
}

int lbmTimerStart(lbm_t h)
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	CUDA_TRY(h, cudaEventRecord(h->ev_t0, h->compute));
	return LBM_OK;
}

int lbmTimerStop(lbm_t h, float *milliseconds)
{
	CHECK_HANDLE(h);
	if (!milliseconds) return fail(h, LBM_ERR_INVALID, "null milliseconds");
	if (int rc = use_device(h)) return rc;
	CUDA_TRY(h, cudaEventRecord(h->ev_t1, h->compute));
	CUDA_TRY(h, cudaEventSynchronize(h->ev_t1));
	CUDA_TRY(h, cudaEventElapsedTime(milliseconds, h->ev_t0, h->ev_t1));
	return LBM_OK;
}

int lbmGetLaunchCount(lbm_t h, uint64_t *launches)
{
	CHECK_HANDLE(h);
	if (!launches) return fail(h, LBM_ERR_INVALID, "null launches");
	*launches = h->launches;
	return LBM_OK;
}

int lbmProfileEnable(lbm_t h, int mode)
{
	CHECK_HANDLE(h);
	if (mode < 0 || mode > (LBM_PROFILE_EVENTS | LBM_PROFILE_NVTX)) return fail(h, LBM_ERR_INVALID, "invalid profile mode");
	if (int rc = use_device(h)) return rc;
	if ((mode & LBM_PROFILE_EVENTS) && !h->prof_base) {
		CUDA_TRY(h, cudaEventCreate(&h->prof_base));
		CUDA_TRY(h, cudaEventRecord(h->prof_base, h->compute));
	}
	h->profile_mode = mode;
	return LBM_OK;
}

int lbmProfileClear(lbm_t h)
{
	CHECK_HANDLE(h);
	if (int rc = use_device(h)) return rc;
	CUDA_TRY(h, cudaStreamSynchronize(h->compute));
	CUDA_TRY(h, cudaStreamSynchronize(h->comm));
	for (size_t i = 0; i < h->prof.size(); i++) { h->prof_pool.push_back(h->prof[i].e0); h->prof_pool.push_back(h->prof[i].e1); }
	h->prof.clear();
	h->prof_dropped = 0;
	if (h->prof_base) CUDA_TRY(h, cudaEventRecord(h->prof_base, h->compute));
	return LBM_OK;
}

int lbmProfileEventCount(lbm_t h, uint64_t *count, uint64_t *dropped)
{
	CHECK_HANDLE(h);
	if (!count) return fail(h, LBM_ERR_INVALID, "null count");
	*count = h->prof.size();
	if (dropped) *dropped = h->prof_dropped;
	return LBM_OK;
}

int lbmProfileGetEvent(lbm_t h, uint64_t index, char *name, size_t name_bytes, uint64_t *start_ns, uint64_t *end_ns)
{
	CHECK_HANDLE(h);
	if (index >= h->prof.size()) return fail(h, LBM_ERR_INVALID, "profile event index out of range");
	if (int rc = use_device(h)) return rc;
	const lbm_solver::prof_rec &r = h->prof[index];
	CUDA_TRY(h, cudaEventSynchronize(r.e1));
	float t0 = 0, t1 = 0;
	CUDA_TRY(h, cudaEventElapsedTime(&t0, h->prof_base, r.e0));
	CUDA_TRY(h, cudaEventElapsedTime(&t1, h->prof_base, r.e1));
	if (t0 < 0) t0 = 0;
	if (t1 < t0) t1 = t0;
	if (name && name_bytes > 0) { strncpy(name, r.name, name_bytes - 1); name[name_bytes - 1] = 0; }
	/* CL_PROFILING_COMMAND_START/END are nanoseconds (CProfilerEvent.hpp:31-35) */
	if (start_ns) *start_ns = (uint64_t)((double)t0 * 1e6 + 0.5);
	if (end_ns) *end_ns = (uint64_t)((double)t1 * 1e6 + 0.5);
	return LBM_OK;
}

int lbmGetConfig(lbm_t h, int *vector_width, int *block_size, int *wg_quirk)
{
	CHECK_HANDLE(h);
	if (vector_width) *vector_width = h->vec;
	if (block_size) *block_size = h->block;
	if (wg_quirk) *wg_quirk = h->wg_quirk;
	return LBM_OK;
}

} /* extern "C" */
