/*
 * lbm_kernels.cuh -- sm_100a device code of the D3Q19 alpha/beta time step.
 *
 * What the kernels compute is defined by the reference's OpenCL kernels
 * (src/cl_programs/lbm_alpha.cl, lbm_beta.cl, lbm_init.cl, copy_buffer_rect.cl); HOW they
 * compute it is B200-first:
 *   - one thread owns VEC consecutive x cells (shipped: 2 -> 64-bit LDG/STG; 4 x fp32 / 2 x fp64 = 16 B is
 *     available and measured slower: registers): every slot whose lattice vector has e_x = 0 (all
 *     19 in alpha, 9 of 19 in beta) moves as one vector LDG/STG; the +-1 shifted slots of beta move as
 *     32-bit (VEC = 2) or 32/64/32-bit (VEC = 4) pieces;
 *   - no shared memory and no barriers: the reference's __local x-shift staging
 *     (lbm_beta.cl:167-234, 7 barriers) exists to align loads on 2010 hardware; here the
 *     L1/L2 absorb the one-element overlap between neighbouring threads and HBM traffic
 *     stays at the algorithmic 2 x 19 values per cell;
 *   - the periodic linear wrap and the work-group quirk only concern the outermost cells:
 *     a block decides uniformly whether it needs the general (scalar, wrapping) path;
 *   - floating-point expressions are written in the reference's operation order and the
 *     translation unit is compiled with --fmad=false, so fp32/fp64 results are
 *     bit-identical to the reference kernels executed in strict IEEE arithmetic.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

/* launch bounds are a tuning knob (registers vs resident warps); see DESIGN.md */
#if defined(LBM_LB_MAXT) && defined(LBM_LB_MINB)
#define LBM_LB_ALPHA
#define LBM_LB_BETA __launch_bounds__(LBM_LB_MAXT, LBM_LB_MINB)
#else
#define LBM_LB_ALPHA
#define LBM_LB_BETA
#endif

namespace lbm {

enum : int { FLAG_OBSTACLE = 1, FLAG_FLUID = 2, FLAG_LID = 4, FLAG_GHOST = 8 };

template <typename T>
struct StepParams {
	T *dd;
	const int *flags;
	T *velocity;
	T *density;
	long long n;             /* cells */
	long long ns;            /* slot stride of dd in cells (>= n; padded so that the 19 slot
	                            streams of a warp do not alias onto one L2 slice / HBM channel) */
	int sx, sy, sz;
	long long sxy;
	T inv_tau, tau, smag_k;
	T gx, gy, gz, u_lid;
	/* iteration box (cells): x0/nx multiples of VEC */
	int x0, nx, y0, ny, z0, nz;
	/* grid rows >= zsplit map to z0 + row + zjump: two z ranges in ONE launch */
	int zsplit, zjump;
	int wg;                  /* >0: emulate lbm_beta.cl:221-234 work-group x-shift */
	int store_v, store_r;    /* write velocity / density arrays (only read when STORE) */
	/* beta: element offset of location (slot i, cell + e_i) from the cell's slot-0 address,
	 * i*ns + e_i . (1, sx, sxy).  Launch-uniform, so an address is `base + boff[i]` with the
	 * offset read from the constant bank: cheap to rematerialise for the push, which keeps 18
	 * 64-bit pointers out of the register file between pull and push (see beta_uses_offset_table) */
	long long boff[18];
	/* fused x-face halo exchange (XFUSE instantiations; lbmCommStep with the z,y,x phase order).
	 * An x ghost column is one element per row: gathering / scattering it with separate kernels touches
	 * one DRAM page per row and slot and costs 15 % of a 256^3 step.  Instead the thread that owns the
	 * cell next to an x ghost face (x = 1 / x = sx-2)
	 *   - PUSH: stores the 5 populations the x neighbour consumes, which it holds in registers, straight
	 *     into the neighbour's receive block [position][z][y] (peer memory over NVLink);
	 *   - PULL: takes the 5 populations it consumes from the neighbour out of MY receive block instead
	 *     of the ghost column (beta) / its own slots (alpha), so the block is never scattered into dd
	 *     on the step path (the host-side entry points materialise it on demand).
	 * xstage[side] / xpull[side] == NULL: nothing to push / nothing pending on that side. */
	T *xstage[2];
	const T *xpull[2];
	long long xface_n;       /* sy * sz: cells of an x face = stride between staging positions */
	/* element offsets of the five positions of a receive / staging block, relative to the lane's face index:
	 * [0] beta, low lane: k * xface_n + (0, -1, +1, -sy, +sy) -- position k holds the slot with
	 * (e_y, e_z) = (0,0) (-1,0) (1,0) (0,-1) (0,1), and a beta lane reads / writes location c + e;
	 * [1] beta, high lane (e_y, e_z mirrored); [2] alpha: k * xface_n */
	int xoff[3][5];
	/* XFUSE launches cover whole rows with a 3-D grid (blocks of a row, rows, z rows; the last block of a row
	 * may be partly empty): a block's place in its row and the row's index come straight from blockIdx.
	 * xkey_hi: x_key() of the thread that holds the cell next to the high x face (x = sx-2),
	 * (block << 10) | thread of group (sx-2) / VEC.
	 * LBM_XFUSE_GRID3D=0 (2-D grid, rows must be whole blocks): row of a block = blockIdx.x / xbpr as a
	 * multiply, xmagic = ceil(2^32 / xbpr), exact while blockIdx.x * xbpr < 2^32; 0 when xbpr == 1. */
	unsigned int xbpr, xmagic, xkey_hi;
};

/* slots an x face ships, ascending = their position in the staging block.
 *   toward +x (e_x = +1): what the LOW face sends after an alpha step (my x=1 column -> the neighbour's
 *                         high ghost column) and what the HIGH face sends after a beta step (my high ghost
 *                         column -> the neighbour's x=1 column);
 *   toward -x (e_x = -1): the other two cases.  (lbmHaloSlotMask, 5-slot payload) */
__device__ __constant__ const int kXPlus[5] = { 0, 4, 6, 8, 10 };
__device__ __constant__ const int kXMinus[5] = { 1, 5, 7, 9, 11 };

/* ---------------------------------------------------------------- vector access
 * Cache-operator tuning knobs for the ALIGNED slot streams (each value is touched once per step,
 * so nothing is lost by not keeping it): LBM_HINT_LD 0 = ld.global (default), 1 = .cs (evict
 * first), 2 = .cg (L2 only); LBM_HINT_ST 0 = st.global (default), 1 = .cs, 2 = .cg.  The +-1
 * shifted pieces of beta always use the default operators: they live on L1 reuse. */
#ifndef LBM_HINT_LD
#define LBM_HINT_LD 0
#endif
#ifndef LBM_HINT_ST
#define LBM_HINT_ST 0
#endif
template <typename V> __device__ __forceinline__ V hinted_load(const V *p)
{
#if LBM_HINT_LD == 1
	return __ldcs(p);
#elif LBM_HINT_LD == 2
	return __ldcg(p);
#else
	return *p;
#endif
}
template <typename V> __device__ __forceinline__ void hinted_store(V *p, V v)
{
#if LBM_HINT_ST == 1
	__stcs(p, v);
#elif LBM_HINT_ST == 2
	__stcg(p, v);
#else
	*p = v;
#endif
}

template <typename T, int VEC> struct VecIO;

template <> struct VecIO<float, 4> {
	static __device__ __forceinline__ void load(const float *p, float (&v)[4]) {
		float4 t = hinted_load(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
	}
	static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
		hinted_store(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
	}
	/* p is 4 B past / before a 16 B boundary: 32 + 64 + 32 bit pieces */
	static __device__ __forceinline__ void load_shifted(const float *p, float (&v)[4]) {
		v[0] = p[0];
		float2 t = *reinterpret_cast<const float2 *>(p + 1); v[1] = t.x; v[2] = t.y;
		v[3] = p[3];
	}
	static __device__ __forceinline__ void store_shifted(float *p, const float (&v)[4]) {
		p[0] = v[0];
		*reinterpret_cast<float2 *>(p + 1) = make_float2(v[1], v[2]);
		p[3] = v[3];
	}
};
template <> struct VecIO<float, 2> {
	static __device__ __forceinline__ void load(const float *p, float (&v)[2]) {
		float2 t = hinted_load(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y;
	}
	static __device__ __forceinline__ void store(float *p, const float (&v)[2]) {
		hinted_store(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
	}
	static __device__ __forceinline__ void load_shifted(const float *p, float (&v)[2]) { v[0] = p[0]; v[1] = p[1]; }
	static __device__ __forceinline__ void store_shifted(float *p, const float (&v)[2]) { p[0] = v[0]; p[1] = v[1]; }
};
template <> struct VecIO<double, 2> {
	static __device__ __forceinline__ void load(const double *p, double (&v)[2]) {
		double2 t = hinted_load(reinterpret_cast<const double2 *>(p)); v[0] = t.x; v[1] = t.y;
	}
	static __device__ __forceinline__ void store(double *p, const double (&v)[2]) {
		hinted_store(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1]));
	}
	static __device__ __forceinline__ void load_shifted(const double *p, double (&v)[2]) { v[0] = p[0]; v[1] = p[1]; }
	static __device__ __forceinline__ void store_shifted(double *p, const double (&v)[2]) { p[0] = v[0]; p[1] = v[1]; }
};
template <typename T> struct VecIO<T, 1> {
	static __device__ __forceinline__ void load(const T *p, T (&v)[1]) { v[0] = p[0]; }
	static __device__ __forceinline__ void store(T *p, const T (&v)[1]) { p[0] = v[0]; }
	static __device__ __forceinline__ void load_shifted(const T *p, T (&v)[1]) { v[0] = p[0]; }
	static __device__ __forceinline__ void store_shifted(T *p, const T (&v)[1]) { p[0] = v[0]; }
};

template <int VEC> struct FlagIO;
template <> struct FlagIO<4> {
	static __device__ __forceinline__ void load(const int *p, int (&v)[4]) {
		int4 t = *reinterpret_cast<const int4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
	}
};
template <> struct FlagIO<2> {
	static __device__ __forceinline__ void load(const int *p, int (&v)[2]) {
		int2 t = *reinterpret_cast<const int2 *>(p); v[0] = t.x; v[1] = t.y;
	}
};
template <> struct FlagIO<1> {
	static __device__ __forceinline__ void load(const int *p, int (&v)[1]) { v[0] = p[0]; }
};

/* ---------------------------------------------------------------- equilibrium
 * lbm_header.h:68-94; constants are float quotients cast to T, as in the reference. */
template <typename T> struct Eq {
	static __device__ __forceinline__ T w18() { return (T)(1.0f / 18.0f); }
	static __device__ __forceinline__ T w36() { return (T)(1.0f / 36.0f); }
	static __device__ __forceinline__ T w3()  { return (T)(1.0f / 3.0f); }
	static __device__ __forceinline__ T a0(T v, T v2, T p) { return w18() * ((p + (T)(3.0f) * v) + (T)(9.0f / 2.0f) * v2); }
	static __device__ __forceinline__ T a1(T v, T v2, T p) { return w18() * ((p + (T)(-3.0f) * v) + (T)(9.0f / 2.0f) * v2); }
	static __device__ __forceinline__ T d4(T v, T v2, T p) { return w36() * ((p + (T)(3.0f) * v) + (T)(9.0f / 2.0f) * v2); }
	static __device__ __forceinline__ T d5(T v, T v2, T p) { return w36() * ((p + (T)(-3.0f) * v) + (T)(9.0f / 2.0f) * v2); }
	static __device__ __forceinline__ T c18(T p) { return w3() * p; }
};

__device__ __forceinline__ float  lbm_sqrt(float x)  { return sqrtf(x); }
__device__ __forceinline__ double lbm_sqrt(double x) { return sqrt(x); }

/* all 19 equilibria for (p = rho - 1.5 u^2, u) */
template <typename T>
__device__ __forceinline__ void equilibria(T (&eq)[19], T vx, T vy, T vz, T p)
{
	T v2, vv;
	v2 = vx * vx; eq[0] = Eq<T>::a0(vx, v2, p); eq[1] = Eq<T>::a1(vx, v2, p);
	v2 = vy * vy; eq[2] = Eq<T>::a0(vy, v2, p); eq[3] = Eq<T>::a1(vy, v2, p);
	vv = vx + vy; v2 = vv * vv; eq[4] = Eq<T>::d4(vv, v2, p); eq[5] = Eq<T>::d5(vv, v2, p);
	vv = vx - vy; v2 = vv * vv; eq[6] = Eq<T>::d4(vv, v2, p); eq[7] = Eq<T>::d5(vv, v2, p);
	vv = vx + vz; v2 = vv * vv; eq[8] = Eq<T>::d4(vv, v2, p); eq[9] = Eq<T>::d5(vv, v2, p);
	vv = vx - vz; v2 = vv * vv; eq[10] = Eq<T>::d4(vv, v2, p); eq[11] = Eq<T>::d5(vv, v2, p);
	vv = vy + vz; v2 = vv * vv; eq[12] = Eq<T>::d4(vv, v2, p); eq[13] = Eq<T>::d5(vv, v2, p);
	vv = vy - vz; v2 = vv * vv; eq[14] = Eq<T>::d4(vv, v2, p); eq[15] = Eq<T>::d5(vv, v2, p);
	v2 = vz * vz; eq[16] = Eq<T>::a0(vz, v2, p); eq[17] = Eq<T>::a1(vz, v2, p);
	eq[18] = Eq<T>::c18(p);
}

/* Smagorinsky closure: specification = oracle/lbm_oracle_impl.h smag_inv_tau (no reference
 * counterpart).  Returns 1/tau_eff. */
template <typename T>
__device__ __forceinline__ T smagorinsky_inv_tau(const T (&d)[19], const T (&eq)[19], T rho, T tau, T smag_k)
{
	T q[19];
#pragma unroll
	for (int i = 0; i < 19; i++) q[i] = d[i] - eq[i];
	T pxx = q[0]; pxx += q[1]; pxx += q[4]; pxx += q[5]; pxx += q[6]; pxx += q[7];
	pxx += q[8]; pxx += q[9]; pxx += q[10]; pxx += q[11];
	T pyy = q[2]; pyy += q[3]; pyy += q[4]; pyy += q[5]; pyy += q[6]; pyy += q[7];
	pyy += q[12]; pyy += q[13]; pyy += q[14]; pyy += q[15];
	T pzz = q[8]; pzz += q[9]; pzz += q[10]; pzz += q[11]; pzz += q[12]; pzz += q[13];
	pzz += q[14]; pzz += q[15]; pzz += q[16]; pzz += q[17];
	T pxy = q[4]; pxy += q[5]; pxy -= q[6]; pxy -= q[7];
	T pxz = q[8]; pxz += q[9]; pxz -= q[10]; pxz -= q[11];
	T pyz = q[12]; pyz += q[13]; pyz -= q[14]; pyz -= q[15];
	T diag = pxx * pxx; diag += pyy * pyy; diag += pzz * pzz;
	T off = pxy * pxy; off += pxz * pxz; off += pyz * pyz;
	T pi_norm = lbm_sqrt(diag + (T)2.0f * off);
	T t = tau * tau + (smag_k * pi_norm) / rho;
	T tau_eff = (T)0.5f * (tau + lbm_sqrt(t));
	return (T)1.0f / tau_eff;
}

/* moments in slot order: lbm_alpha.cl:61-156 and lbm_beta.cl:53-164 */
template <typename T>
__device__ __forceinline__ void moments_linear(const T (&d)[19], T &rho, T &vx, T &vy, T &vz)
{
	rho = d[0]; vx = d[0];
	rho += d[1]; vx -= d[1];
	rho += d[2]; vy = d[2];
	rho += d[3]; vy -= d[3];
	rho += d[4]; vx += d[4]; vy += d[4];
	rho += d[5]; vx -= d[5]; vy -= d[5];
	rho += d[6]; vx += d[6]; vy -= d[6];
	rho += d[7]; vx -= d[7]; vy += d[7];
	rho += d[8]; vx += d[8]; vz = d[8];
	rho += d[9]; vx -= d[9]; vz -= d[9];
	rho += d[10]; vx += d[10]; vz -= d[10];
	rho += d[11]; vx -= d[11]; vz += d[11];
	rho += d[12]; vy += d[12]; vz += d[12];
	rho += d[13]; vy -= d[13]; vz -= d[13];
	rho += d[14]; vy += d[14]; vz -= d[14];
	rho += d[15]; vy -= d[15]; vz += d[15];
	rho += d[16]; vz += d[16];
	rho += d[17]; vz -= d[17];
	rho += d[18];
}

/* moments in the order of the shipped shared-memory beta path: lbm_beta.cl:268-483 */
template <typename T>
__device__ __forceinline__ void moments_shipped(const T (&d)[19], T &rho, T &vx, T &vy, T &vz)
{
	rho = d[3]; vy = -d[3];
	rho += d[2]; vy += d[2];
	rho += d[0]; vx = d[0];
	rho += d[1]; vx -= d[1];
	rho += d[4]; vx += d[4]; vy += d[4];
	rho += d[5]; vx -= d[5]; vy -= d[5];
	rho += d[6]; vx += d[6]; vy -= d[6];
	rho += d[7]; vx -= d[7]; vy += d[7];
	rho += d[8]; vx += d[8]; vz = d[8];
	rho += d[9]; vx -= d[9]; vz -= d[9];
	rho += d[10]; vx += d[10]; vz -= d[10];
	rho += d[11]; vx -= d[11]; vz += d[11];
	rho += d[13]; vy -= d[13]; vz -= d[13];
	rho += d[12]; vy += d[12]; vz += d[12];
	rho += d[15]; vy -= d[15]; vz += d[15];
	rho += d[14]; vy += d[14]; vz -= d[14];
	rho += d[17]; vz -= d[17];
	rho += d[16]; vz += d[16];
	rho += d[18];
}

template <typename T>
__device__ __forceinline__ void swap_pairs(T (&d)[19])
{
#pragma unroll
	for (int i = 0; i < 18; i += 2) { T t = d[i + 1]; d[i + 1] = d[i]; d[i] = t; }
}

/* ---------------------------------------------------------------- alpha cell update
 * lbm_alpha.cl:177-497.  On return d[i] holds the value the reference stores for
 * direction i (it lands in slot i^1), `rho` the value it stores as density
 * (`#define tmp rho`, :173).  Returns false when the reference writes nothing (obstacle,
 * ghost, unknown flag): the caller then leaves the cell's slots untouched. */
template <typename T, bool SMAG>
__device__ __forceinline__ bool alpha_cell(T (&d)[19], int flag, const StepParams<T> &P,
		T &rho, T &vx, T &vy, T &vz)
{
	moments_linear(d, rho, vx, vy, vz);
	if (flag == FLAG_FLUID) {
		const T vel2 = vx * vx + vy * vy + vz * vz;
		const T p = rho - (T)(3.0f / 2.0f) * vel2;
		T w = P.inv_tau;
		if (SMAG) {
			T eq[19];
			equilibria(eq, vx, vy, vz, p);
			w = smagorinsky_inv_tau(d, eq, rho, P.tau, P.smag_k);
		}
		T v2, vv;
		rho = P.gx * (T)(1.0f / 18.0f) * rho;
		v2 = vx * vx;
		d[1] += w * (Eq<T>::a1(vx, v2, p) - d[1]); d[1] -= rho;
		d[0] += w * (Eq<T>::a0(vx, v2, p) - d[0]); d[0] += rho;
		rho = P.gy * (T)(-1.0f / 18.0f) * rho;
		v2 = vy * vy;
		d[3] += w * (Eq<T>::a1(vy, v2, p) - d[3]); d[3] -= rho;
		d[2] += w * (Eq<T>::a0(vy, v2, p) - d[2]); d[2] += rho;
		vv = vx + vy; v2 = vv * vv;
		rho = (P.gx - P.gy) * (T)(1.0f / 36.0f) * rho;
		d[5] += w * (Eq<T>::d5(vv, v2, p) - d[5]); d[5] -= rho;
		d[4] += w * (Eq<T>::d4(vv, v2, p) - d[4]); d[4] += rho;
		vv = vx - vy; v2 = vv * vv;
		rho = (P.gx + P.gy) * (T)(1.0f / 36.0f) * rho;
		d[7] += w * (Eq<T>::d5(vv, v2, p) - d[7]); d[7] -= rho;
		d[6] += w * (Eq<T>::d4(vv, v2, p) - d[6]); d[6] += rho;
		vv = vx + vz; v2 = vv * vv;
		rho = (P.gx + P.gz) * (T)(1.0f / 36.0f) * rho;
		d[9] += w * (Eq<T>::d5(vv, v2, p) - d[9]); d[9] -= rho;
		d[8] += w * (Eq<T>::d4(vv, v2, p) - d[8]); d[8] += rho;
		rho = (P.gx - P.gz) * (T)(1.0f / 36.0f) * rho;
		vv = vx - vz; v2 = vv * vv;
		d[11] += w * (Eq<T>::d5(vv, v2, p) - d[11]); d[11] -= rho;
		d[10] += w * (Eq<T>::d4(vv, v2, p) - d[10]); d[10] += rho;
		vv = vy + vz; v2 = vv * vv;
		rho = (P.gz - P.gy) * (T)(1.0f / 36.0f) * rho;
		d[13] += w * (Eq<T>::d5(vv, v2, p) - d[13]); d[13] -= rho;
		d[12] += w * (Eq<T>::d4(vv, v2, p) - d[12]); d[12] += rho;
		vv = vy - vz; v2 = vv * vv;
		rho = (P.gz + P.gy) * (T)(-1.0f / 36.0f) * rho;
		d[15] += w * (Eq<T>::d5(vv, v2, p) - d[15]); d[15] -= rho;
		d[14] += w * (Eq<T>::d4(vv, v2, p) - d[14]); d[14] += rho;
		v2 = vz * vz;
		rho = P.gz * (T)(1.0f / 18.0f) * rho;
		d[17] += w * (Eq<T>::a1(vz, v2, p) - d[17]); d[17] -= rho;
		d[16] += w * (Eq<T>::a0(vz, v2, p) - d[16]); d[16] += rho;
		d[18] += w * (Eq<T>::c18(p) - d[18]);
		return true;
	}
	if (flag == FLAG_LID) {
		vx = P.u_lid; vy = 0; vz = 0;
		rho = 1.0f;
		const T vel2 = vx * vx + vy * vy + vz * vz;
		const T p = rho - (T)(3.0f / 2.0f) * vel2;
		T v2, vv;
		v2 = vx * vx;
		rho = P.gx * (T)(1.0f / 18.0f) * rho;
		d[1] = Eq<T>::a1(vx, v2, p); d[1] -= rho;
		d[0] = Eq<T>::a0(vx, v2, p); d[0] += rho;
		v2 = vy * vy;
		rho = P.gy * (T)(-1.0f / 18.0f) * rho;
		d[3] = Eq<T>::a1(vy, v2, p); d[3] -= rho;
		d[2] = Eq<T>::a0(vy, v2, p); d[2] += rho;
		vv = vx + vy; v2 = vv * vv;
		rho = (P.gx - P.gy) * (T)(1.0f / 36.0f) * rho;
		d[5] = Eq<T>::d5(vv, v2, p); d[5] -= rho;
		d[4] = Eq<T>::d4(vv, v2, p); d[4] += rho;
		vv = vx - vy; v2 = vv * vv;
		rho = (P.gx + P.gy) * (T)(1.0f / 36.0f) * rho;
		d[7] = Eq<T>::d5(vv, v2, p); d[7] -= rho;
		d[6] = Eq<T>::d4(vv, v2, p); d[6] += rho;
		vv = vx + vz; v2 = vv * vv;
		rho = (P.gx + P.gz) * (T)(1.0f / 36.0f) * rho;
		d[9] = Eq<T>::d5(vv, v2, p); d[9] -= rho;
		d[8] = Eq<T>::d4(vv, v2, p); d[8] += rho;
		vv = vx - vz; v2 = vv * vv;
		rho = (P.gx - P.gz) * (T)(1.0f / 36.0f) * rho;
		d[11] = Eq<T>::d5(vv, v2, p); d[11] -= rho;
		d[10] = Eq<T>::d4(vv, v2, p); d[10] += rho;
		vv = vy + vz; v2 = vv * vv;
		rho = (P.gz - P.gy) * (T)(1.0f / 36.0f) * rho;
		d[13] = Eq<T>::d5(vv, v2, p); d[13] -= rho;
		d[12] = Eq<T>::d4(vv, v2, p); d[12] += rho;
		vv = vy - vz; v2 = vv * vv;
		rho = (P.gz + P.gy) * (T)(-1.0f / 36.0f) * rho;
		d[15] = Eq<T>::d5(vv, v2, p); d[15] -= rho;
		d[14] = Eq<T>::d4(vv, v2, p); d[14] += rho;
		v2 = vz * vz;
		rho = P.gz * (T)(1.0f / 18.0f) * rho;
		d[17] = Eq<T>::a1(vz, v2, p); d[17] -= rho;
		d[16] = Eq<T>::a0(vz, v2, p); d[16] += rho;
		d[18] = Eq<T>::c18(p);
		return true;
	}
	if (flag == FLAG_OBSTACLE) { vx = 0.0f; vy = 0.0f; vz = 0.0f; }
	return false;
}

/* ---------------------------------------------------------------- beta cell update
 * lbm_beta.cl:495-656 (+ moments).  d[i] in: pulled populations; out: values to push.
 * `rho` out = what the reference stores as density (`#define dd_param rho`, :493). */
template <typename T, bool SMAG, int ORDER>
__device__ __forceinline__ void beta_cell(T (&d)[19], int flag, const StepParams<T> &P,
		T &rho, T &vx, T &vy, T &vz)
{
	if (ORDER == 0) moments_shipped(d, rho, vx, vy, vz);
	else moments_linear(d, rho, vx, vy, vz);
	if (flag == FLAG_FLUID) {
		const T vel2 = vx * vx + vy * vy + vz * vz;
		T w = P.inv_tau;
		const T rho_sum = rho;
		rho = rho - (T)(3.0f / 2.0f) * vel2;
		if (SMAG) {
			T eq[19];
			equilibria(eq, vx, vy, vz, rho);
			w = smagorinsky_inv_tau(d, eq, rho_sum, P.tau, P.smag_k);
		}
		T v2, vv;
		v2 = vx * vx;
		d[0] += w * (Eq<T>::a0(vx, v2, rho) - d[0]);
		d[1] += w * (Eq<T>::a1(vx, v2, rho) - d[1]);
		v2 = vy * vy;
		d[2] += w * (Eq<T>::a0(vy, v2, rho) - d[2]);
		d[3] += w * (Eq<T>::a1(vy, v2, rho) - d[3]);
		vv = vx + vy; v2 = vv * vv;
		d[4] += w * (Eq<T>::d4(vv, v2, rho) - d[4]);
		d[5] += w * (Eq<T>::d5(vv, v2, rho) - d[5]);
		vv = vx - vy; v2 = vv * vv;
		d[6] += w * (Eq<T>::d4(vv, v2, rho) - d[6]);
		d[7] += w * (Eq<T>::d5(vv, v2, rho) - d[7]);
		vv = vx + vz; v2 = vv * vv;
		d[8] += w * (Eq<T>::d4(vv, v2, rho) - d[8]);
		d[9] += w * (Eq<T>::d5(vv, v2, rho) - d[9]);
		vv = vx - vz; v2 = vv * vv;
		d[10] += w * (Eq<T>::d4(vv, v2, rho) - d[10]);
		d[11] += w * (Eq<T>::d5(vv, v2, rho) - d[11]);
		vv = vy + vz; v2 = vv * vv;
		d[12] += w * (Eq<T>::d4(vv, v2, rho) - d[12]);
		d[13] += w * (Eq<T>::d5(vv, v2, rho) - d[13]);
		vv = vy - vz; v2 = vv * vv;
		d[14] += w * (Eq<T>::d4(vv, v2, rho) - d[14]);
		d[15] += w * (Eq<T>::d5(vv, v2, rho) - d[15]);
		v2 = vz * vz;
		d[16] += w * (Eq<T>::a0(vz, v2, rho) - d[16]);
		d[17] += w * (Eq<T>::a1(vz, v2, rho) - d[17]);
		d[18] += w * (Eq<T>::c18(rho) - d[18]);
	} else if (flag == FLAG_OBSTACLE) {
		vx = 0.0f; vy = 0.0f; vz = 0.0f;
		swap_pairs(d);
	} else if (flag == FLAG_LID) {
		vx = P.u_lid; vy = 0; vz = 0;
		rho = 1.0f;
		const T vel2 = vx * vx + vy * vy + vz * vz;
		rho = rho - (T)(3.0f / 2.0f) * vel2;
		equilibria(d, vx, vy, vz, rho);
	}
}

/* LBM_XFUSE_GRID3D=1 (default): XFUSE launches use a 3-D grid (blocks of a row, rows of the box, z rows), so
 * that a block's place in its row and the row's index come straight from blockIdx; 0: every launch is
 * (blocks of a plane of the box, z rows) and the row comes from a multiply by StepParams::xmagic.
 * Measured (profiles/r2/xfuse_lane_ab.md): the 3-D grid is 0.3-1 % of a step faster. */
#ifndef LBM_XFUSE_GRID3D
#define LBM_XFUSE_GRID3D 1
#endif
template <typename T, bool XG>
__device__ __forceinline__ int box_z(const StepParams<T> &P)
{
	const int row = (int)(XG ? blockIdx.z : blockIdx.y);
	return P.z0 + row + (row >= P.zsplit ? P.zjump : 0);
}

/* thread -> first cell of its VEC-wide group inside the iteration box; false = out of box */
template <typename T, int VEC, bool XG>
__device__ __forceinline__ bool box_cell(const StepParams<T> &P, long long &gid)
{
	if (XG) {
		/* whole rows, gridDim.x blocks each (the last one may be partly empty) */
		const unsigned int x = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
		if (x >= (unsigned int)P.sx) return false;
		gid = (long long)box_z<T, XG>(P) * P.sxy + (long long)(P.y0 + (int)blockIdx.y) * P.sx + x;
		return true;
	}
	const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
	if (t >= (long long)P.nx * P.ny) return false;
	long long off;
	if (P.nx == P.sx) off = (long long)P.y0 * P.sx + t;
	else { const int iy = (int)(t / P.nx), ix = (int)(t - (long long)iy * P.nx); off = (long long)(P.y0 + iy) * P.sx + P.x0 + ix; }
	gid = (long long)box_z<T, XG>(P) * P.sxy + off;
	return true;
}

/* XFUSE lanes: the cell next to the low x ghost face (x = 1) is element E_LO of group T_LO, held by thread T_LO
 * of the first block of a row; the one next to the high face (x = sx-2) is element E_HI of group (sx-2) / VEC,
 * a fixed thread of a fixed block of the row (StepParams::xkey_hi). */
template <int VEC> struct XLane {
	enum { T_LO = VEC == 1 ? 1 : 0,
	       E_LO = VEC == 1 ? 0 : 1,
	       E_HI = VEC == 1 ? 0 : VEC - 2 };
};
/* row of this block inside the box (XFUSE launches: whole rows, xbpr blocks each) */
template <typename T, bool XG>
__device__ __forceinline__ unsigned int x_row(const StepParams<T> &P)
{
	if (XG) return blockIdx.y;
	return P.xmagic ? __umulhi(blockIdx.x, P.xmagic) : blockIdx.x;
}
/* (block in row, thread) as one word: the lanes are the threads whose key is XLane::T_LO (low face) or
 * StepParams::xkey_hi (high face) */
template <typename T, bool XG>
__device__ __forceinline__ unsigned int x_key(const StepParams<T> &P)
{
	const unsigned int in_row = XG ? blockIdx.x : blockIdx.x - x_row<T, XG>(P) * P.xbpr;
	return (in_row << 10) | threadIdx.x;
}
/* index of the lane's cell in the x face [z][y] (a face has < 2^31 cells) */
template <typename T, bool XG>
__device__ __forceinline__ int x_faceidx(const StepParams<T> &P)
{
	return box_z<T, XG>(P) * P.sy + P.y0 + (int)x_row<T, XG>(P);
}

/* ================================================================== ALPHA kernel */
template <typename T, int VEC, bool SMAG, bool STORE, bool XFUSE>
__global__ void LBM_LB_ALPHA lbm_alpha_kernel(const StepParams<T> P)
{
	constexpr bool XG = XFUSE && LBM_XFUSE_GRID3D;
	long long gid;
	if (!box_cell<T, VEC, XG>(P, gid)) return;

	int flag[VEC];
	FlagIO<VEC>::load(P.flags + gid, flag);
	bool any_write = false;
#pragma unroll
	for (int e = 0; e < VEC; e++) any_write |= (flag[e] == FLAG_FLUID) | (flag[e] == FLAG_LID);
	bool all_ghost = true;
#pragma unroll
	for (int e = 0; e < VEC; e++) all_ghost &= (flag[e] == FLAG_GHOST);
	/* XFUSE PULL, issued first so that it overlaps the 19 slot loads: what the x neighbour's beta step
	 * streamed into the cell next to the face sits in my receive block (the reference's
	 * setDensityDistribution(..., norm) would have scattered it into these slots).  Into registers of
	 * their own: a second load into a slot's register would wait for the first (the scoreboard is per warp). */
	const unsigned int key = XFUSE ? x_key<T, XG>(P) : 0u;
	const bool lane = XFUSE && (key == (unsigned int)XLane<VEC>::T_LO || key == P.xkey_hi);
	int side = 0;
	T px[5];                                     /* only read when pulled */
	bool pulled = false, xlane = false;
	if (XFUSE && lane) {
		side = key == P.xkey_hi ? 1 : 0;
		const T *st = P.xpull[side];
		xlane = st || P.xstage[side];
		if (st) {
			const int f = x_faceidx<T, XG>(P);                 /* f + offset < 5 * face cells < 2^31 (host) */
			pulled = true;
#pragma unroll
			for (int k = 0; k < 5; k++) px[k] = __ldcg(st + (f + P.xoff[2][k]));
		}
	}
	/* lbm_alpha.cl:31-32: ghost cells are skipped; obstacle cells write nothing (:305-343).  A group
	 * that holds the cell next to an x face still pulls / ships that cell's slots. */
	if (all_ghost && !xlane) return;
	if (!any_write && !STORE && !xlane) return;

	T v[19][VEC];
	T *base = P.dd + gid;
#pragma unroll
	for (int i = 0; i < 19; i++) VecIO<T, VEC>::load(base + (long long)i * P.ns, v[i]);

	if (XFUSE && pulled) {
		if (side == 0) {
#pragma unroll
			for (int k = 0; k < 5; k++) v[2 * k + (k ? 2 : 0)][XLane<VEC>::E_LO] = px[k];       /* low face: slots 0,4,6,8,10 */
		} else {
#pragma unroll
			for (int k = 0; k < 5; k++) v[2 * k + (k ? 3 : 1)][XLane<VEC>::E_HI] = px[k];       /* high face: slots 1,5,7,9,11 */
		}
	}

	T orho[VEC], ovx[VEC], ovy[VEC], ovz[VEC];
#pragma unroll
	for (int e = 0; e < VEC; e++) {
		T d[19];
#pragma unroll
		for (int i = 0; i < 19; i++) d[i] = v[i][e];
		const bool wrote = alpha_cell<T, SMAG>(d, flag[e], P, orho[e], ovx[e], ovy[e], ovz[e]);
		if (!wrote) {
			/* the reference leaves slot j = old d[j]; the store below writes d[j^1] to slot j */
#pragma unroll
			for (int i = 0; i < 19; i++) d[i] = v[i][e];
			swap_pairs(d);
		}
#pragma unroll
		for (int i = 0; i < 19; i++) v[i][e] = d[i];
	}
	/* `pulled`: a cell the step does not update (obstacle, ghost rim) keeps what it pulled in dd, as
	 * if it had been unpacked there */
	if (any_write || pulled) {
#pragma unroll
		for (int i = 0; i < 18; i++) VecIO<T, VEC>::store(base + (long long)i * P.ns, v[i ^ 1]);
		VecIO<T, VEC>::store(base + 18LL * P.ns, v[18]);
	}
	if (XFUSE && lane) {
		/* PUSH: slot j holds v[j^1] after this step.  Ghost cells (rims of the y/z faces) are left to
		 * the rim pass that follows the y/z unpack (halo_xrim_flag_kernel). */
		T *st = P.xstage[side];
		const int xflag = side == 0 ? flag[XLane<VEC>::E_LO] : flag[XLane<VEC>::E_HI];
		if (st && xflag != FLAG_GHOST) {
			const int f = x_faceidx<T, XG>(P);
			if (side == 0) {
#pragma unroll
				for (int k = 0; k < 5; k++) st[f + P.xoff[2][k]] = v[2 * k + (k ? 3 : 1)][XLane<VEC>::E_LO];      /* slots 0,4,6,8,10 <- v[1,5,7,9,11] */
			} else {
#pragma unroll
				for (int k = 0; k < 5; k++) st[f + P.xoff[2][k]] = v[2 * k + (k ? 2 : 0)][XLane<VEC>::E_HI];      /* slots 1,5,7,9,11 <- v[0,4,6,8,10] */
			}
		}
	}
	if (STORE) {
#pragma unroll
		for (int e = 0; e < VEC; e++) {
			if (flag[e] == FLAG_GHOST) continue;
			if (P.store_v) {
				P.velocity[gid + e] = ovx[e];
				P.velocity[P.n + gid + e] = ovy[e];
				P.velocity[2 * P.n + gid + e] = ovz[e];
			}
			if (P.store_r) P.density[gid + e] = orho[e];
		}
	}
}

/* ================================================================== BETA kernels
 * Two kernels share one thread->cell mapping and one block-uniform predicate:
 *   lbm_beta_kernel          vectorised, wrap-free path; blocks that need wrapping exit;
 *   lbm_beta_general_kernel  scalar path with the reference's linear periodic wrap
 *                            (wrap.h:112-127) and work-group x-shift; only the blocks the
 *                            fast kernel skipped do work (launched on the outer planes only).
 * Keeping them apart keeps the wrap bookkeeping out of the hot kernel's register budget. */
template <typename T, int VEC, bool XG>
__device__ __forceinline__ bool beta_block_is_general(const StepParams<T> &P)
{
	if (P.wg > 0) return true;
	long long o0, o1;                            /* first and last cell of the block, inside its plane */
	if (XG) {
		const long long row = (long long)(P.y0 + (int)blockIdx.y) * P.sx;
		const long long x0 = (long long)blockIdx.x * blockDim.x * VEC;
		long long x1 = x0 + (long long)blockDim.x * VEC - 1;
		if (x1 > P.sx - 1) x1 = P.sx - 1;
		o0 = row + x0; o1 = row + x1;
	} else {
		const long long t0 = (long long)blockIdx.x * blockDim.x * VEC;
		long long t1 = t0 + (long long)blockDim.x * VEC - 1;
		const long long tmax = (long long)P.nx * P.ny - 1;
		if (t1 > tmax) t1 = tmax;
		if (P.nx == P.sx) { o0 = (long long)P.y0 * P.sx + t0; o1 = (long long)P.y0 * P.sx + t1; }
		else { o0 = (long long)(P.y0 + t0 / P.nx) * P.sx; o1 = (long long)(P.y0 + t1 / P.nx) * P.sx + P.sx - 1; }
	}
	const long long zb = (long long)box_z<T, XG>(P) * P.sxy;
	const long long reach = P.sxy + P.sx + 1;
	return (zb + o0 < reach) || (zb + o1 + reach >= P.n);
}

/* How the 18 pull/push addresses are formed: from the constant-bank offset table
 * (StepParams::boff) or from DY/DZ/ns arithmetic.  Same addresses, different instruction
 * schedule; measured per instantiation (profiles/r1_sweep_beta_offtab.log).
 * LBM_BETA_OFFTAB: 0 never, 1 always, 2 (default) the measured choice. */
#ifndef LBM_BETA_OFFTAB
#define LBM_BETA_OFFTAB 2
#endif
template <typename T, bool SMAG, bool STORE>
__device__ __forceinline__ constexpr bool beta_uses_offset_table()
{
	/* fp32 Smagorinsky without output stores: 92 -> 80 registers, +5.6 % (measured).  With output
	 * stores the table costs registers (95 -> 127), fp32 BGK likewise (96 -> 114, -4.5 % measured),
	 * fp64 stays at 3 resident blocks either way (+-0 measured): arithmetic there. */
	return LBM_BETA_OFFTAB == 1 || (LBM_BETA_OFFTAB == 2 && SMAG && !STORE && sizeof(T) == 4);
}

template <typename T, int VEC, bool SMAG, bool STORE, int ORDER, bool XFUSE>
__global__ void LBM_LB_BETA lbm_beta_kernel(const StepParams<T> P)
{
	constexpr bool XG = XFUSE && LBM_XFUSE_GRID3D;
	long long gid;
	if (!box_cell<T, VEC, XG>(P, gid)) return;
	if (beta_block_is_general<T, VEC, XG>(P)) return;

	const long long DY = P.sx, DZ = P.sxy;
	/* XFUSE PULL, issued first so that it overlaps the slot loads: location (slot j, c + e_j) in the ghost
	 * column next to the x = 1 (x = sx-2) cell holds what the neighbour's alpha step left there -- it sits
	 * in my receive block at face index row(c) + e_y + e_z * sy (in range: this path is a plane + a row
	 * away from the array ends); it is read as d[j^1].  Into registers of their own: a second load into a
	 * slot's register would wait for the first (the scoreboard is per warp). */
	const unsigned int key = XFUSE ? x_key<T, XG>(P) : 0u;
	const bool lane = XFUSE && (key == (unsigned int)XLane<VEC>::T_LO || key == P.xkey_hi);
	int side = 0;
	T px[5];                                     /* only read when pulled */
	bool pulled = false;
	if (XFUSE && lane) {
		side = key == P.xkey_hi ? 1 : 0;
		const T *st = P.xpull[side];
		if (st) {
			const int f = x_faceidx<T, XG>(P);                 /* f + offset < 5 * face cells < 2^31 (host) */
			pulled = true;
			if (side == 0) {                                   /* slots 1,5,7,9,11: (-1,0,0) (-1,-1,0) (-1,1,0) (-1,0,-1) (-1,0,1) */
#pragma unroll
				for (int k = 0; k < 5; k++) px[k] = __ldcg(st + (f + P.xoff[0][k]));
			} else {                                           /* slots 0,4,6,8,10: mirrored */
#pragma unroll
				for (int k = 0; k < 5; k++) px[k] = __ldcg(st + (f + P.xoff[1][k]));
			}
		}
	}
	int flag[VEC];
	FlagIO<VEC>::load(P.flags + gid, flag);

	/* location (slot j, cell c + e_j) is read as d[j^1] and written as d[j] */
	T *base = P.dd + gid;
	T *loc[18];
	if (beta_uses_offset_table<T, SMAG, STORE>()) {
#pragma unroll
		for (int i = 0; i < 18; i++) loc[i] = base + P.boff[i];
	} else {
		loc[0] = base + 1;            loc[1] = base - 1;
		loc[2] = base + DY;           loc[3] = base - DY;
		loc[4] = base + 1 + DY;       loc[5] = base - 1 - DY;
		loc[6] = base + 1 - DY;       loc[7] = base - 1 + DY;
		loc[8] = base + 1 + DZ;       loc[9] = base - 1 - DZ;
		loc[10] = base + 1 - DZ;      loc[11] = base - 1 + DZ;
		loc[12] = base + DY + DZ;     loc[13] = base - DY - DZ;
		loc[14] = base + DY - DZ;     loc[15] = base - DY + DZ;
		loc[16] = base + DZ;          loc[17] = base - DZ;
#pragma unroll
		for (int i = 0; i < 18; i++) loc[i] += (long long)i * P.ns;
	}

	T v[19][VEC];
#pragma unroll
	for (int i = 0; i < 18; i++) {
		const bool shifted = (i < 2) || (i >= 4 && i < 12);     /* e_x != 0 */
		if (shifted) VecIO<T, VEC>::load_shifted(loc[i], v[i ^ 1]);
		else VecIO<T, VEC>::load(loc[i], v[i ^ 1]);
	}
	VecIO<T, VEC>::load(base + 18LL * P.ns, v[18]);

	if (XFUSE && pulled) {
		if (side == 0) {
#pragma unroll
			for (int k = 0; k < 5; k++) v[2 * k + (k ? 2 : 0)][XLane<VEC>::E_LO] = px[k];       /* read as d[j^1], j = 1,5,7,9,11 */
		} else {
#pragma unroll
			for (int k = 0; k < 5; k++) v[2 * k + (k ? 3 : 1)][XLane<VEC>::E_HI] = px[k];       /* j = 0,4,6,8,10 */
		}
	}

	T orho[VEC], ovx[VEC], ovy[VEC], ovz[VEC];
#pragma unroll
	for (int e = 0; e < VEC; e++) {
		T d[19];
#pragma unroll
		for (int i = 0; i < 19; i++) d[i] = v[i][e];
		beta_cell<T, SMAG, ORDER>(d, flag[e], P, orho[e], ovx[e], ovy[e], ovz[e]);
#pragma unroll
		for (int i = 0; i < 19; i++) v[i][e] = d[i];
	}
#pragma unroll
	for (int i = 0; i < 18; i++) {
		const bool shifted = (i < 2) || (i >= 4 && i < 12);
		if (shifted) VecIO<T, VEC>::store_shifted(loc[i], v[i]);
		else VecIO<T, VEC>::store(loc[i], v[i]);
	}
	VecIO<T, VEC>::store(base + 18LL * P.ns, v[18]);

	if (XFUSE && lane) {
		/* PUSH: the x = 1 (x = sx-2) cell is the only writer of the e_x = -1 (+1) slots of the ghost
		 * column next to it; the same values go to the neighbour, same face index as above. */
		T *st = P.xstage[side];
		if (st) {
			const int f = x_faceidx<T, XG>(P);
			if (side == 0) {
#pragma unroll
				for (int k = 0; k < 5; k++) st[f + P.xoff[0][k]] = v[2 * k + (k ? 3 : 1)][XLane<VEC>::E_LO];      /* slots 1,5,7,9,11 */
			} else {
#pragma unroll
				for (int k = 0; k < 5; k++) st[f + P.xoff[1][k]] = v[2 * k + (k ? 2 : 0)][XLane<VEC>::E_HI];      /* slots 0,4,6,8,10 */
			}
		}
	}

	if (STORE) {
#pragma unroll
		for (int e = 0; e < VEC; e++) {
			if (flag[e] == FLAG_GHOST) continue;
			if (P.store_v) {
				P.velocity[gid + e] = ovx[e];
				P.velocity[P.n + gid + e] = ovy[e];
				P.velocity[2 * P.n + gid + e] = ovz[e];
			}
			if (P.store_r) P.density[gid + e] = orho[e];
		}
	}
}

template <typename T, int VEC, bool SMAG, bool STORE, int ORDER, bool XFUSE>
__global__ void lbm_beta_general_kernel(const StepParams<T> P)
{
	long long gid;
	constexpr bool XG = XFUSE && LBM_XFUSE_GRID3D;
	if (!box_cell<T, VEC, XG>(P, gid)) return;
	if (!beta_block_is_general<T, VEC, XG>(P)) return;
	const long long DY = P.sx, DZ = P.sxy;
	const int gx0 = XFUSE ? (int)(gid % P.sx) : 0;
#pragma unroll 1
	for (int e = 0; e < VEC; e++) {
		const long long c = gid + e;
		const int flag = P.flags[c];
		long long xm = c - 1, xp = c + 1;
		if (P.wg > 0) {
			const int lid = (int)(c % P.wg);
			if (lid == 0) xm = c + P.wg - 1;
			if (lid == P.wg - 1) xp = c - (P.wg - 1);
		}
		long long L[18];
		L[0] = xp;            L[1] = xm;
		L[2] = c + DY;        L[3] = c - DY;
		L[4] = xp + DY;       L[5] = xm - DY;
		L[6] = xp - DY;       L[7] = xm + DY;
		L[8] = xp + DZ;       L[9] = xm - DZ;
		L[10] = xp - DZ;      L[11] = xm + DZ;
		L[12] = c + DY + DZ;  L[13] = c - DY - DZ;
		L[14] = c + DY - DZ;  L[15] = c - DY + DZ;
		L[16] = c + DZ;       L[17] = c - DZ;
#pragma unroll
		for (int i = 0; i < 18; i++) {
			while (L[i] < 0) L[i] += P.n;
			while (L[i] >= P.n) L[i] -= P.n;
			L[i] += (long long)i * P.ns;
		}
		T d[19];
#pragma unroll
		for (int i = 0; i < 18; i++) d[i ^ 1] = P.dd[L[i]];
		d[18] = P.dd[18LL * P.ns + c];
		/* XFUSE works by LOCATION here (the wrap and the work-group quirk move cells around): whatever
		 * this cell reads out of / writes into an x ghost column with a neighbour behind it comes from
		 * my receive block / also goes to the neighbour's */
		const bool xedge = XFUSE && (P.wg > 0 || gx0 + e <= 1 || gx0 + e >= P.sx - 2);
		if (XFUSE && xedge) {
#pragma unroll
			for (int k = 0; k < 5; k++) {
				const int jm = 2 * k + (k ? 3 : 1), jp = 2 * k + (k ? 2 : 0);      /* kXMinus[k], kXPlus[k] */
				if (P.xpull[0]) {
					const long long loc = L[jm] - (long long)jm * P.ns;
					if (loc % P.sx == 0) d[jm ^ 1] = __ldcg(P.xpull[0] + k * P.xface_n + loc / P.sx);
				}
				if (P.xpull[1]) {
					const long long loc = L[jp] - (long long)jp * P.ns;
					if (loc % P.sx == P.sx - 1) d[jp ^ 1] = __ldcg(P.xpull[1] + k * P.xface_n + loc / P.sx);
				}
			}
		}
		T rho, vx, vy, vz;
		beta_cell<T, SMAG, ORDER>(d, flag, P, rho, vx, vy, vz);
#pragma unroll
		for (int i = 0; i < 18; i++) P.dd[L[i]] = d[i];
		P.dd[18LL * P.ns + c] = d[18];
		if (XFUSE && xedge) {
#pragma unroll
			for (int k = 0; k < 5; k++) {
				const int jm = 2 * k + (k ? 3 : 1), jp = 2 * k + (k ? 2 : 0);
				if (P.xstage[0]) {
					const long long loc = L[jm] - (long long)jm * P.ns;
					if (loc % P.sx == 0) P.xstage[0][k * P.xface_n + loc / P.sx] = d[jm];
				}
				if (P.xstage[1]) {
					const long long loc = L[jp] - (long long)jp * P.ns;
					if (loc % P.sx == P.sx - 1) P.xstage[1][k * P.xface_n + loc / P.sx] = d[jp];
				}
			}
		}
		if (STORE && flag != FLAG_GHOST) {
			if (P.store_v) { P.velocity[c] = vx; P.velocity[P.n + c] = vy; P.velocity[2 * P.n + c] = vz; }
			if (P.store_r) P.density[c] = rho;
		}
	}
}

/* ================================================================== INIT kernel
 * lbm_init.cl:32-237 */
template <typename T>
__global__ void lbm_init_kernel(T *dd, int *flags, T *velocity, T *density,
		long long n, long long ns, int sx, int sy, int sz, int b0, int b1, int b2, int b3, int b4, int b5,
		int store_v, int store_r)
{
	const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= n) return;
	const int x = (int)(gid % sx);
	const int y = (int)((gid / sx) % sy);
	const int z = (int)(gid / ((long long)sx * sy));
	int flag = FLAG_FLUID;
	if (x == 0) flag = b0;
	else if (x == sx - 1) flag = b1;
	else if (y == 0) flag = b2;
	else if (y == sy - 1) flag = b3;
	else if (z == 0) flag = b4;
	else if (z == sz - 1) flag = b5;
	T eq[19];
	const T rho = 1.0f;
	const T vx = 0, vy = 0, vz = 0;
	const T p = rho - (T)(3.0f / 2.0f) * (vx * vx);       /* :133-134 */
	equilibria(eq, vx, vy, vz, p);
#pragma unroll
	for (int i = 0; i < 19; i++) dd[(long long)i * ns + gid] = eq[i];
	flags[gid] = flag;
	if (store_v) { velocity[gid] = vx; velocity[n + gid] = vy; velocity[2 * n + gid] = vz; }
	if (store_r) density[gid] = rho;
}

/* ================================================================== rect copies
 * copy_buffer_rect.cl:12-52 generalised: all components in ONE launch (the reference
 * enqueues one launch per slot, src/CLbmSolver.hpp:704-711), 64-bit offsets (the
 * reference's int offsets overflow at 512^3), optional per-component selection. */
struct RectCopy {
	long long src_comp_stride, dst_comp_stride;
	int so[3], ss[3];        /* src origin, src array size  */
	int dorg[3], ds[3];      /* dst origin, dst array size  */
	int block[3];
	int ncomp;               /* components to copy */
	int src_comp[19];        /* component index on the src side */
	int dst_comp[19];        /* component index on the dst side */
};

template <typename T>
__global__ void rect_copy_kernel(const T *__restrict__ src, T *__restrict__ dst, const RectCopy R)
{
	const long long cells = (long long)R.block[0] * R.block[1] * R.block[2];
	const long long total = cells * R.ncomp;
	for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
			t += (long long)gridDim.x * blockDim.x) {
		const int c = (int)(t / cells);
		long long r = t - (long long)c * cells;
		const int k = (int)(r / ((long long)R.block[0] * R.block[1]));
		r -= (long long)k * R.block[0] * R.block[1];
		const int j = (int)(r / R.block[0]);
		const int i = (int)(r - (long long)j * R.block[0]);
		const long long s = (long long)R.src_comp[c] * R.src_comp_stride + (R.so[0] + i)
				+ (long long)(R.so[1] + j) * R.ss[0] + (long long)(R.so[2] + k) * R.ss[0] * R.ss[1];
		const long long d = (long long)R.dst_comp[c] * R.dst_comp_stride + (R.dorg[0] + i)
				+ (long long)(R.dorg[1] + j) * R.ds[0] + (long long)(R.dorg[2] + k) * R.ds[0] * R.ds[1];
		dst[d] = src[s];
	}
}


/* ================================================================== peer-memory halo
 * One-sided halo exchange over NVLink peer memory (same process or CUDA-IPC mapped):
 *   halo_push_kernel       packs the selected slots of the (up to two) faces of one axis and stores
 *                          them STRAIGHT INTO the neighbours' staging buffers (peer stores: pack +
 *                          send fused); the last block of a face publishes a sequence number in the
 *                          neighbour's flag word;
 *   XFUSE step kernels     the same for x faces from inside lbm_alpha/beta_kernel (the owning threads
 *                          hold the values in registers), completed by halo_xrim_flag_kernel;
 *   halo_wait_kernel       one thread per face waits (acquire, system scope) until the local flag
 *                          reaches the next sequence number;
 *   halo_unpack_kernel     unpacks the local staging buffers into the dd rects.
 * All sequence numbers are counted in device memory: a captured CUDA graph can be replayed.
 * Replaces storeDensityDistribution -> MPI_Isend/Irecv/Waitall -> setDensityDistribution
 * (reference src/CController.hpp:265-383). */
/* One halo face of a fused launch.  dd side: [slot][z][y][x] rect inside the sub-domain with
 * slot stride dd_stride; staging side: dense [selected slot][z][y][x].  vec = elements moved
 * per thread and step (4/2 when rows are whole, aligned multiples -- y and z faces; 1 for x faces). */
struct HaloFace {
	void *dd;                       /* sub-domain populations (local) */
	void *staging;                  /* push: the NEIGHBOUR's receive block (peer memory); unpack: mine */
	unsigned int *block_counter;    /* push only: blocks of this launch that are done (local) */
	unsigned int *sync_count;       /* push: syncs of this kind I have pushed on this face; wait: syncs I have
	                                   pulled.  DEVICE-resident, so a replayed CUDA graph keeps counting */
	volatile unsigned int *flag;    /* push: the neighbour's flag word; wait: mine */
	long long dd_stride;
	int origin[3], size[3];         /* rect in the sub-domain */
	int ss[2];                      /* sub-domain Sx, Sy */
	int ncomp;
	int dd_comp[19];                /* slot on the dd side   */
	int st_comp[19];                /* position on the staging side */
	int vec;
};
struct HaloAxis { HaloFace f[2]; };

/* element e (in units of vec) of a face -> offsets on both sides; 32-bit arithmetic */
template <typename T>
__device__ __forceinline__ void halo_offsets(const HaloFace &F, unsigned int e, long long &dd_off, long long &st_off)
{
	const unsigned int xw = (unsigned int)F.size[0] / (unsigned int)F.vec;
	const unsigned int i = e % xw; unsigned int r = e / xw;
	const unsigned int j = r % (unsigned int)F.size[1]; r /= (unsigned int)F.size[1];
	const unsigned int k = r % (unsigned int)F.size[2];
	const unsigned int c = r / (unsigned int)F.size[2];
	const long long cells = (long long)F.size[0] * F.size[1] * F.size[2];
	dd_off = (long long)F.dd_comp[c] * F.dd_stride + (F.origin[0] + (long long)i * F.vec)
			+ (long long)(F.origin[1] + j) * F.ss[0] + (long long)(F.origin[2] + k) * F.ss[0] * F.ss[1];
	st_off = (long long)F.st_comp[c] * cells + ((long long)k * F.size[1] + j) * F.size[0] + (long long)i * F.vec;
}

template <typename T, int VEC> struct HaloVec;
template <> struct HaloVec<float, 4> { typedef float4 type; };
template <> struct HaloVec<float, 2> { typedef float2 type; };
template <> struct HaloVec<float, 1> { typedef float type; };
template <> struct HaloVec<double, 2> { typedef double2 type; };
template <> struct HaloVec<double, 1> { typedef double type; };
template <> struct HaloVec<double, 4> { typedef double2 type; };   /* unused */

template <typename T, int VEC>
__device__ __forceinline__ void halo_copy(const HaloFace &F, bool to_staging)
{
	typedef typename HaloVec<T, VEC>::type V;
	const unsigned int total = (unsigned int)(((long long)F.size[0] * F.size[1] * F.size[2] * F.ncomp) / VEC);
	T *dd = (T *)F.dd, *st = (T *)F.staging;
	for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
		long long a, b;
		halo_offsets<T>(F, e, a, b);
		if (to_staging) *reinterpret_cast<V *>(st + b) = *reinterpret_cast<const V *>(dd + a);
		else *reinterpret_cast<V *>(dd + a) = __ldcg(reinterpret_cast<const V *>(st + b));   /* L2: never a stale L1 line */
	}
}

/* last block of a face: everything this launch stored into the neighbour's block is visible before
 * the neighbour sees the new sequence number */
__device__ __forceinline__ void halo_publish(const HaloFace &F)
{
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system();
		const unsigned int done = atomicAdd(F.block_counter, 1u);
		if (done == gridDim.x - 1) {
			*F.block_counter = 0;                 /* ready for the next push of this face */
			const unsigned int seq = *F.sync_count + 1u;
			*F.sync_count = seq;
			__threadfence_system();
			*F.flag = seq;
		}
	}
}

/* push: both faces of one axis in ONE launch (blockIdx.y = face).  Packs the face and stores it
 * straight into the neighbour's receive block over NVLink; the last block of a face raises the
 * neighbour's flag.  ONE thread per block fences (system scope, cumulative over the block's peer
 * stores it observed through the barrier): a fence of scope >= cluster invalidates the SM's L1
 * (CCTL.IVALL), which the step kernel running next to this one depends on. */
template <typename T>
__global__ void halo_push_kernel(const HaloAxis A)
{
	const HaloFace &F = A.f[blockIdx.y];
	if (F.vec == 4) halo_copy<T, sizeof(T) == 4 ? 4 : 2>(F, true);
	else if (F.vec == 2) halo_copy<T, 2>(F, true);
	else halo_copy<T, 1>(F, true);
	halo_publish(F);
}

__device__ __forceinline__ unsigned int ld_acquire_sys(const volatile unsigned int *p)
{
	unsigned int v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_sys(volatile unsigned int *p, unsigned int v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

/* wait until the neighbour's push has raised my flag to the next sequence number (acquire at system
 * scope).  The expected number is counted in device memory.  A neighbour that never arrives (crashed
 * rank) must not hang the GPU: after timeout_ns the wait gives up and leaves a mark the host finds
 * in lbmWait. */
__device__ __forceinline__ void halo_wait_face(const HaloFace &F, unsigned long long timeout_ns, unsigned int *error_word)
{
	const unsigned int seq = *F.sync_count + 1u;
	*F.sync_count = seq;
	unsigned long long t0 = 0;
	unsigned int spins = 0;
	/* sequence numbers only grow; signed distance tolerates wrap-around */
	while ((int)(ld_acquire_sys(F.flag) - seq) < 0) {
		__nanosleep(20);
		if ((++spins & 1023u) == 0) {
			unsigned long long now;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			if (t0 == 0) t0 = now;
			else if (now - t0 > timeout_ns) { atomicExch(error_word, 1u); break; }
		}
	}
}

/* x faces whose bulk went out of the step kernels (XFUSE): the rim pass.  The cells of an x face that lie
 * in the two outermost layers of y or z hold values the y/z phases of THIS sync delivered (edge
 * populations on their way to the diagonal neighbour) or are ghost cells the step kernel skipped; they
 * are re-read from dd -- final by now: this kernel runs behind the step kernels and the y/z unpack -- and
 * stored over whatever the step kernel sent for them.  Then the flag goes up.  F.origin[0] = column.
 * blockIdx.y = face; blockIdx.x < rim_blocks: the rim lines, one element per thread (the kernel is the
 * exposed tail of a step: every gather is a DRAM round trip, so they all go out at once); the block that
 * finishes last counts the sync up and releases the flag.  do_rim == 0 (no y/z neighbours, nothing to
 * forward; rim_blocks == 1): the flag only.
 * nwait > 0: blockIdx.x == rim_blocks of each face is a one-thread waiter for the neighbour's flag (W = my
 * receive side), running NEXT TO the rim pass: the whole exposed x tail of a step is this one launch. */
template <typename T>
__global__ void halo_xrim_flag_kernel(const HaloAxis A, const HaloAxis W, int nwait, int rim_blocks, int do_rim,
		unsigned long long timeout_ns, unsigned int *error_word)
{
	if ((int)blockIdx.x == rim_blocks) {
		if (threadIdx.x == 0 && (int)blockIdx.y < nwait) halo_wait_face(W.f[blockIdx.y], timeout_ns, error_word);
		return;
	}
	const HaloFace &F = A.f[blockIdx.y];
	if (do_rim) {
		const int sy = F.size[1], sz = F.size[2];
		const unsigned int line_cells = 4u * (unsigned int)(sy + sz);
		const unsigned int total = line_cells * (unsigned int)F.ncomp;
		const T *dd = (const T *)F.dd;
		T *st = (T *)F.staging;
		for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (unsigned int)rim_blocks * blockDim.x) {
			const unsigned int c = t / line_cells, q = t - c * line_cells;
			int y, z;
			if (q < 4u * (unsigned int)sz) {                       /* rows y = 0, 1, sy-2, sy-1 */
				const int w = (int)(q / (unsigned int)sz); z = (int)(q - (unsigned int)w * sz);
				y = w < 2 ? w : sy - 4 + w;
			} else {                                               /* rows z = 0, 1, sz-2, sz-1 */
				const unsigned int r = q - 4u * (unsigned int)sz;
				const int w = (int)(r / (unsigned int)sy); y = (int)(r - (unsigned int)w * sy);
				z = w < 2 ? w : sz - 4 + w;
			}
			if (y < 0 || y >= sy || z < 0 || z >= sz) continue;    /* faces thinner than 4 */
			const long long face = (long long)z * sy + y;
			st[(long long)F.st_comp[c] * sy * sz + face] =
				dd[(long long)F.dd_comp[c] * F.dd_stride + F.origin[0] + (long long)y * F.ss[0] + (long long)z * F.ss[0] * F.ss[1]];
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		bool last = true;
		if (rim_blocks > 1) {
			__threadfence_system();                  /* this block's peer stores (barrier) before its count */
			last = atomicAdd(F.block_counter, 1u) == (unsigned int)rim_blocks - 1u;
			if (last) { *F.block_counter = 0; __threadfence_system(); }      /* ready for the next sync of this face */
		}
		if (last) {
			const unsigned int seq = *F.sync_count + 1u;
			*F.sync_count = seq;
			/* release at system scope, cumulative over the rim blocks' stores (counted above) and over
			 * the step kernels' peer stores (stream order) */
			st_release_sys(F.flag, seq);
		}
	}
}

/* wait: ONE thread per face (not a grid of spinning blocks next to the step kernel) */
__global__ void halo_wait_kernel(const HaloAxis A, int nfaces, unsigned long long timeout_ns, unsigned int *error_word)
{
	if ((int)threadIdx.x < nfaces) halo_wait_face(A.f[threadIdx.x], timeout_ns, error_word);
}

/* unpack: my receive block(s) of one axis -> the dd rects; runs behind halo_wait_kernel in stream
 * order, so nothing spins here */
template <typename T>
__global__ void halo_unpack_kernel(const HaloAxis A)
{
	const HaloFace &F = A.f[blockIdx.y];
	if (F.vec == 4) halo_copy<T, sizeof(T) == 4 ? 4 : 2>(F, false);
	else if (F.vec == 2) halo_copy<T, 2>(F, false);
	else halo_copy<T, 1>(F, false);
}

/* ================================================================== checksum
 * device-side variant of CLbmSolver::getVelocityChecksum: sum over FLUID cells of
 * (ux+uy)+uz, accumulated in double with warp shuffles, one atomic per block. */
template <typename T>
__global__ void checksum_kernel(const T *__restrict__ velocity, const int *__restrict__ flags,
		long long n, double *out)
{
	double acc = 0.0;
	for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < n;
			a += (long long)gridDim.x * blockDim.x)
		if (flags[a] == FLAG_FLUID)
			acc += (double)((velocity[a] + velocity[n + a]) + velocity[2 * n + a]);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
	__shared__ double warp_sums[32];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (lane == 0) warp_sums[wid] = acc;
	__syncthreads();
	if (wid == 0) {
		acc = (lane < (blockDim.x + 31) / 32) ? warp_sums[lane] : 0.0;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
		if (lane == 0) atomicAdd(out, acc);
	}
}

} /* namespace lbm */
