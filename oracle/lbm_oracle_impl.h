/*
 * lbm_oracle_impl.h -- CPU restatement of the reference's D3Q19 kernels.
 * TEST INFRASTRUCTURE ONLY; included twice by lbm_oracle.c (T = float, T = double).
 *
 * Every function cites the reference lines it restates.  Floating-point expressions are
 * written in the reference's operation order and must be compiled with
 * -ffp-contract=off so that they evaluate in strict IEEE source order (no FMA).
 * The restatement is pinned bit-for-bit against the reference's own kernel sources
 * executed on the CPU (oracle/_ref, tests/test_oracle_vs_ref.py) and against the
 * committed golden vectors in tests/golden/.
 *
 * Layout (reference src/cl_programs/lbm_header.h:15-27): dd[f*N + gid], gid = x + y*Sx +
 * z*Sx*Sy, slot order
 *   0 (1,0,0) 1 (-1,0,0) 2 (0,1,0) 3 (0,-1,0) 4 (1,1,0) 5 (-1,-1,0) 6 (1,-1,0) 7 (-1,1,0)
 *   8 (1,0,1) 9 (-1,0,-1) 10 (1,0,-1) 11 (-1,0,1) 12 (0,1,1) 13 (0,-1,-1) 14 (0,1,-1)
 *   15 (0,-1,1) 16 (0,0,1) 17 (0,0,-1) 18 (0,0,0)
 */

/* equilibrium macros: lbm_header.h:68-94.  The constants are *float* quotients cast to T
 * (so T = double also uses float-rounded 1/18, 1/36, 1/3), exactly as the reference. */
#define W18  ((T)(1.0f/18.0f))
#define W36  ((T)(1.0f/36.0f))
#define W3   ((T)(1.0f/3.0f))
#define EQ_A0(v, v2, p)  (W18*((p) + (T)(3.0f)*(v) + (T)(9.0f/2.0f)*(v2)))
#define EQ_A1(v, v2, p)  (W18*((p) + (T)(-3.0f)*(v) + (T)(9.0f/2.0f)*(v2)))
#define EQ_4(v, v2, p)   (W36*((p) + (T)(3.0f)*(v) + (T)(9.0f/2.0f)*(v2)))
#define EQ_5(v, v2, p)   (W36*((p) + (T)(-3.0f)*(v) + (T)(9.0f/2.0f)*(v2)))
#define EQ_18(p)         (W3*(p))

/* ------------------------------------------------------------------------------------
 * init_kernel, lbm_init.cl:32-237: flag from position and bc[6] (x wins on edges, :53-68),
 * all 19 slots := equilibrium of rho = 1, u = 0, velocity := 0, density := 1.
 */
void FN(lbmo_init)(T *dd, int *flags, T *velocity, T *density, const int *bc,
		int sx, int sy, int sz, int store_velocity, int store_density)
{
	const long n = (long)sx * sy * sz;
	long gid;
	#pragma omp parallel for schedule(static)
	for (gid = 0; gid < n; gid++) {
		const int x = (int)(gid % sx), y = (int)((gid / sx) % sy), z = (int)(gid / ((long)sx * sy));
		int flag = LBMO_FLAG_FLUID;
		if (x == 0) flag = bc[0];
		else if (x == sx - 1) flag = bc[1];
		else if (y == 0) flag = bc[2];
		else if (y == sy - 1) flag = bc[3];
		else if (z == 0) flag = bc[4];
		else if (z == sz - 1) flag = bc[5];

		const T vx = 0, vy = 0, vz = 0;
		T rho = 1.0f;
		T vela2, vv, p;
		vela2 = vx*vx;
		p = rho - (T)(3.0f/2.0f)*(vela2);
		dd[0*n + gid] = EQ_A0(vx, vela2, p);
		dd[1*n + gid] = EQ_A1(vx, vela2, p);
		vela2 = vy*vy;
		dd[2*n + gid] = EQ_A0(vy, vela2, p);
		dd[3*n + gid] = EQ_A1(vy, vela2, p);
		vv = vx+vy; vela2 = vv*vv;
		dd[4*n + gid] = EQ_4(vv, vela2, p);
		dd[5*n + gid] = EQ_5(vv, vela2, p);
		vv = vx-vy; vela2 = vv*vv;
		dd[6*n + gid] = EQ_4(vv, vela2, p);
		dd[7*n + gid] = EQ_5(vv, vela2, p);
		vv = vx+vz; vela2 = vv*vv;
		dd[8*n + gid] = EQ_4(vv, vela2, p);
		dd[9*n + gid] = EQ_5(vv, vela2, p);
		vv = vx-vz; vela2 = vv*vv;
		dd[10*n + gid] = EQ_4(vv, vela2, p);
		dd[11*n + gid] = EQ_5(vv, vela2, p);
		vv = vy+vz; vela2 = vv*vv;
		dd[12*n + gid] = EQ_4(vv, vela2, p);
		dd[13*n + gid] = EQ_5(vv, vela2, p);
		vv = vy-vz; vela2 = vv*vv;
		dd[14*n + gid] = EQ_4(vv, vela2, p);
		dd[15*n + gid] = EQ_5(vv, vela2, p);
		vela2 = vz*vz;
		dd[16*n + gid] = EQ_A0(vz, vela2, p);
		dd[17*n + gid] = EQ_A1(vz, vela2, p);
		dd[18*n + gid] = EQ_18(p);
		flags[gid] = flag;
		if (store_velocity) { velocity[gid] = vx; velocity[n + gid] = vy; velocity[2*n + gid] = vz; }
		if (store_density) density[gid] = rho;
	}
}

/*
 * Smagorinsky eddy viscosity -- NEW FEATURE, no reference counterpart (the reference is
 * plain BGK).  This function *defines* the operation order the CUDA
 * kernels reproduce:  neq_i = dd_i - feq_i;  Pi_ab = sum_i e_ia e_ib neq_i;
 * |Pi| = sqrt(Pxx^2+Pyy^2+Pzz^2 + 2(Pxy^2+Pxz^2+Pyz^2));
 * tau_eff = 0.5*(tau + sqrt(tau^2 + smag_k*|Pi|/rho)),  smag_k = 18*sqrt(2)*C_s^2;
 * returns 1/tau_eff.
 */
static inline T FN(smag_inv_tau)(const T *dd, const T *eq, T rho, T tau, T smag_k)
{
	T q[19];
	int i;
	for (i = 0; i < 19; i++) q[i] = dd[i] - eq[i];
	T pxx = q[0]; pxx += q[1]; pxx += q[4]; pxx += q[5]; pxx += q[6]; pxx += q[7];
	pxx += q[8]; pxx += q[9]; pxx += q[10]; pxx += q[11];
	T pyy = q[2]; pyy += q[3]; pyy += q[4]; pyy += q[5]; pyy += q[6]; pyy += q[7];
	pyy += q[12]; pyy += q[13]; pyy += q[14]; pyy += q[15];
	T pzz = q[8]; pzz += q[9]; pzz += q[10]; pzz += q[11]; pzz += q[12]; pzz += q[13];
	pzz += q[14]; pzz += q[15]; pzz += q[16]; pzz += q[17];
	T pxy = q[4]; pxy += q[5]; pxy -= q[6]; pxy -= q[7];
	T pxz = q[8]; pxz += q[9]; pxz -= q[10]; pxz -= q[11];
	T pyz = q[12]; pyz += q[13]; pyz -= q[14]; pyz -= q[15];
	T diag = pxx*pxx; diag += pyy*pyy; diag += pzz*pzz;
	T off = pxy*pxy; off += pxz*pxz; off += pyz*pyz;
	T pi_norm = SQRT(diag + (T)2.0f*off);
	T t = tau*tau + (smag_k*pi_norm)/rho;
	T tau_eff = (T)0.5f*(tau + SQRT(t));
	return (T)1.0f/tau_eff;
}

/* ------------------------------------------------------------------------------------
 * lbm_kernel_alpha, lbm_alpha.cl:12-498: purely local; reads slot f at gid, writes the
 * post-collision value of direction i into the slot of the OPPOSITE direction (:185-301).
 * `#define tmp rho` (:173): every gravity term overwrites rho and feeds the next one;
 * the stored "density" is the end of that chain.
 */
void FN(lbmo_alpha)(T *dd, const int *flags, T *velocity, T *density,
		int sx, int sy, int sz, T inv_tau, T gx, T gy, T gz, T u_lid,
		T tau, T smag_k, int store_velocity, int store_density)
{
	const long n = (long)sx * sy * sz;
	long gid;
	#pragma omp parallel for schedule(static)
	for (gid = 0; gid < n; gid++) {
		const int flag = flags[gid];
		if (flag == LBMO_FLAG_GHOST) continue;                       /* :31-32 */
		T d[19];
		int i;
		for (i = 0; i < 19; i++) d[i] = dd[(long)i*n + gid];
		T rho, vx, vy, vz;
		/* :61-156, summation strictly in slot order */
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

		T vel2, vela2, vv, p;
		T *o = dd + gid;      /* o[k*n] = slot k at this cell */
		switch (flag) {
		case LBMO_FLAG_FLUID: {                                      /* :179-303 */
			vel2 = vx*vx + vy*vy + vz*vz;
			p = rho - (T)(3.0f/2.0f)*(vel2);
			T w = inv_tau;
			if (smag_k != (T)0) {
				T eq[19];
				vela2 = vx*vx; eq[0] = EQ_A0(vx, vela2, p); eq[1] = EQ_A1(vx, vela2, p);
				vela2 = vy*vy; eq[2] = EQ_A0(vy, vela2, p); eq[3] = EQ_A1(vy, vela2, p);
				vv = vx+vy; vela2 = vv*vv; eq[4] = EQ_4(vv, vela2, p); eq[5] = EQ_5(vv, vela2, p);
				vv = vx-vy; vela2 = vv*vv; eq[6] = EQ_4(vv, vela2, p); eq[7] = EQ_5(vv, vela2, p);
				vv = vx+vz; vela2 = vv*vv; eq[8] = EQ_4(vv, vela2, p); eq[9] = EQ_5(vv, vela2, p);
				vv = vx-vz; vela2 = vv*vv; eq[10] = EQ_4(vv, vela2, p); eq[11] = EQ_5(vv, vela2, p);
				vv = vy+vz; vela2 = vv*vv; eq[12] = EQ_4(vv, vela2, p); eq[13] = EQ_5(vv, vela2, p);
				vv = vy-vz; vela2 = vv*vv; eq[14] = EQ_4(vv, vela2, p); eq[15] = EQ_5(vv, vela2, p);
				vela2 = vz*vz; eq[16] = EQ_A0(vz, vela2, p); eq[17] = EQ_A1(vz, vela2, p);
				eq[18] = EQ_18(p);
				w = FN(smag_inv_tau)(d, eq, rho, tau, smag_k);
			}
			rho = gx*(T)(1.0f/18.0f)*rho;                        /* tmp aliases rho */
			vela2 = vx*vx;
			d[1] += w*(EQ_A1(vx, vela2, p) - d[1]); d[1] -= rho; o[0*n] = d[1];
			d[0] += w*(EQ_A0(vx, vela2, p) - d[0]); d[0] += rho; o[1*n] = d[0];
			rho = gy*(T)(-1.0f/18.0f)*rho;
			vela2 = vy*vy;
			d[3] += w*(EQ_A1(vy, vela2, p) - d[3]); d[3] -= rho; o[2*n] = d[3];
			d[2] += w*(EQ_A0(vy, vela2, p) - d[2]); d[2] += rho; o[3*n] = d[2];
			vv = vx+vy; vela2 = vv*vv;
			rho = (gx - gy)*(T)(1.0f/36.0f)*rho;
			d[5] += w*(EQ_5(vv, vela2, p) - d[5]); d[5] -= rho; o[4*n] = d[5];
			d[4] += w*(EQ_4(vv, vela2, p) - d[4]); d[4] += rho; o[5*n] = d[4];
			vv = vx-vy; vela2 = vv*vv;
			rho = (gx + gy)*(T)(1.0f/36.0f)*rho;
			d[7] += w*(EQ_5(vv, vela2, p) - d[7]); d[7] -= rho; o[6*n] = d[7];
			d[6] += w*(EQ_4(vv, vela2, p) - d[6]); d[6] += rho; o[7*n] = d[6];
			vv = vx+vz; vela2 = vv*vv;
			rho = (gx + gz)*(T)(1.0f/36.0f)*rho;
			d[9] += w*(EQ_5(vv, vela2, p) - d[9]); d[9] -= rho; o[8*n] = d[9];
			d[8] += w*(EQ_4(vv, vela2, p) - d[8]); d[8] += rho; o[9*n] = d[8];
			rho = (gx - gz)*(T)(1.0f/36.0f)*rho;
			vv = vx-vz; vela2 = vv*vv;
			d[11] += w*(EQ_5(vv, vela2, p) - d[11]); d[11] -= rho; o[10*n] = d[11];
			d[10] += w*(EQ_4(vv, vela2, p) - d[10]); d[10] += rho; o[11*n] = d[10];
			vv = vy+vz; vela2 = vv*vv;
			rho = (gz - gy)*(T)(1.0f/36.0f)*rho;
			d[13] += w*(EQ_5(vv, vela2, p) - d[13]); d[13] -= rho; o[12*n] = d[13];
			d[12] += w*(EQ_4(vv, vela2, p) - d[12]); d[12] += rho; o[13*n] = d[12];
			vv = vy-vz; vela2 = vv*vv;
			rho = (gz + gy)*(T)(-1.0f/36.0f)*rho;
			d[15] += w*(EQ_5(vv, vela2, p) - d[15]); d[15] -= rho; o[14*n] = d[15];
			d[14] += w*(EQ_4(vv, vela2, p) - d[14]); d[14] += rho; o[15*n] = d[14];
			vela2 = vz*vz;
			rho = gz*(T)(1.0f/18.0f)*rho;
			d[17] += w*(EQ_A1(vz, vela2, p) - d[17]); d[17] -= rho; o[16*n] = d[17];
			d[16] += w*(EQ_A0(vz, vela2, p) - d[16]); d[16] += rho; o[17*n] = d[16];
			d[18] += w*(EQ_18(p) - d[18]); o[18*n] = d[18];
			break;
		}
		case LBMO_FLAG_OBSTACLE:                                     /* :305-343: nothing written */
			vx = 0.0f; vy = 0.0f; vz = 0.0f;
			break;
		case LBMO_FLAG_LID:                                          /* :345-480 */
			vx = u_lid; vy = 0; vz = 0;
			rho = 1.0f;
			vel2 = vx*vx + vy*vy + vz*vz;
			p = rho - (T)(3.0f/2.0f)*(vel2);
			vela2 = vx*vx;
			rho = gx*(T)(1.0f/18.0f)*rho;
			d[1] = EQ_A1(vx, vela2, p); d[1] -= rho; o[0*n] = d[1];
			d[0] = EQ_A0(vx, vela2, p); d[0] += rho; o[1*n] = d[0];
			vela2 = vy*vy;
			rho = gy*(T)(-1.0f/18.0f)*rho;
			d[3] = EQ_A1(vy, vela2, p); d[3] -= rho; o[2*n] = d[3];
			d[2] = EQ_A0(vy, vela2, p); d[2] += rho; o[3*n] = d[2];
			vv = vx+vy; vela2 = vv*vv;
			rho = (gx - gy)*(T)(1.0f/36.0f)*rho;
			d[5] = EQ_5(vv, vela2, p); d[5] -= rho; o[4*n] = d[5];
			d[4] = EQ_4(vv, vela2, p); d[4] += rho; o[5*n] = d[4];
			vv = vx-vy; vela2 = vv*vv;
			rho = (gx + gy)*(T)(1.0f/36.0f)*rho;
			d[7] = EQ_5(vv, vela2, p); d[7] -= rho; o[6*n] = d[7];
			d[6] = EQ_4(vv, vela2, p); d[6] += rho; o[7*n] = d[6];
			vv = vx+vz; vela2 = vv*vv;
			rho = (gx + gz)*(T)(1.0f/36.0f)*rho;
			d[9] = EQ_5(vv, vela2, p); d[9] -= rho; o[8*n] = d[9];
			d[8] = EQ_4(vv, vela2, p); d[8] += rho; o[9*n] = d[8];
			vv = vx-vz; vela2 = vv*vv;
			rho = (gx - gz)*(T)(1.0f/36.0f)*rho;
			d[11] = EQ_5(vv, vela2, p); d[11] -= rho; o[10*n] = d[11];
			d[10] = EQ_4(vv, vela2, p); d[10] += rho; o[11*n] = d[10];
			vv = vy+vz; vela2 = vv*vv;
			rho = (gz - gy)*(T)(1.0f/36.0f)*rho;
			d[13] = EQ_5(vv, vela2, p); d[13] -= rho; o[12*n] = d[13];
			d[12] = EQ_4(vv, vela2, p); d[12] += rho; o[13*n] = d[12];
			vv = vy-vz; vela2 = vv*vv;
			rho = (gz + gy)*(T)(-1.0f/36.0f)*rho;
			d[15] = EQ_5(vv, vela2, p); d[15] -= rho; o[14*n] = d[15];
			d[14] = EQ_4(vv, vela2, p); d[14] += rho; o[15*n] = d[14];
			vela2 = vz*vz;
			rho = gz*(T)(1.0f/18.0f)*rho;
			d[17] = EQ_A1(vz, vela2, p); d[17] -= rho; o[16*n] = d[17];
			d[16] = EQ_A0(vz, vela2, p); d[16] += rho; o[17*n] = d[16];
			d[18] = EQ_18(p); o[18*n] = d[18];
			break;
		default:
			break;
		}
		if (store_velocity) { velocity[gid] = vx; velocity[n + gid] = vy; velocity[2*n + gid] = vz; }  /* :486-492 */
		if (store_density) density[gid] = rho;                                                        /* :494-497 */
	}
}

/* ------------------------------------------------------------------------------------
 * lbm_kernel_beta, lbm_beta.cl:13-802.  dd_i is pulled from slot opp(i) of the cell at
 * c - e_i (periodic wrap on the LINEAR index, wrap.h:112-127, deltas lbm_header.h:44-51),
 * collided, and pushed to slot i of the cell at c + e_i -- i.e. every (slot, cell) location
 * is read and written by exactly one work-item.
 *
 * order = 0: accumulation order of the kernel AS SHIPPED (USE_SHARED_MEMORY 1, :256-483)
 * order = 1: accumulation order of the reference's alternative path (:53-164)
 * wg  > 0 : emulate the shared-memory path's work-group x-shift; only observable when
 *           wg % sx == 0 (:221-234): then the work-items with lid 0 / wg-1 exchange with
 *           the far end of their own work-group instead of the linear neighbour.
 *           wg = 0: plain linear neighbours (what the order-1 path always does).
 */
void FN(lbmo_beta)(T *dd, const int *flags, T *velocity, T *density,
		int sx, int sy, int sz, T inv_tau, T gx, T gy, T gz, T u_lid,
		T tau, T smag_k, int store_velocity, int store_density, int order, int wg)
{
	const long n = (long)sx * sy * sz;
	const long DY = sx, DZ = (long)sx * sy;
	const int quirk = (wg > 0) && (wg % sx == 0) && (n % wg == 0);
	(void)gx; (void)gy; (void)gz;                                     /* GRAVITATION 0, :2,658-702 */
	/*
	 * Unlike the OpenCL NDRange this loop must not let one cell's writes be seen by
	 * another cell's reads -- but no location is shared between cells, so any order
	 * (and any parallel schedule) gives the same result.
	 */
	long gid;
	#pragma omp parallel for schedule(static)
	for (gid = 0; gid < n; gid++) {
		const int flag = flags[gid];
		long xm = gid - 1, xp = gid + 1;
		if (quirk) {
			const long lid = gid % wg;
			if (lid == 0) xm = gid + wg - 1;
			if (lid == wg - 1) xp = gid - (wg - 1);
		}
#define WRAPI(a) ((((a) % n) + n) % n)
		/* location L[j] = index of (slot j, cell c + e_j): read as dd_opp(j), written as dd_j */
		long L[19];
		L[0] = WRAPI(xp);            L[1] = WRAPI(xm);
		L[2] = WRAPI(gid + DY);      L[3] = WRAPI(gid - DY);
		L[4] = WRAPI(xp + DY);       L[5] = WRAPI(xm - DY);
		L[6] = WRAPI(xp - DY);       L[7] = WRAPI(xm + DY);
		L[8] = WRAPI(xp + DZ);       L[9] = WRAPI(xm - DZ);
		L[10] = WRAPI(xp - DZ);      L[11] = WRAPI(xm + DZ);
		L[12] = WRAPI(gid + DY + DZ); L[13] = WRAPI(gid - DY - DZ);
		L[14] = WRAPI(gid + DY - DZ); L[15] = WRAPI(gid - DY + DZ);
		L[16] = WRAPI(gid + DZ);     L[17] = WRAPI(gid - DZ);
		L[18] = gid;
#undef WRAPI
		T d[19];
		int i;
		for (i = 0; i < 18; i++) d[i ^ 1] = dd[(long)i*n + L[i]];
		d[18] = dd[18*n + gid];

		T rho, vx, vy, vz;
		if (order == 0) {                                            /* :268-483 */
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
		} else {                                                     /* :53-164 */
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

		T vel2, vela2, vv, t;
		switch (flag) {
		case LBMO_FLAG_FLUID: {                                      /* :497-559; dd_param aliases rho (:493) */
			vel2 = vx*vx + vy*vy + vz*vz;
			T w = inv_tau;
			const T rho_sum = rho;
			rho = rho - (T)(3.0f/2.0f)*(vel2);
			if (smag_k != (T)0) {
				T eq[19];
				vela2 = vx*vx; eq[0] = EQ_A0(vx, vela2, rho); eq[1] = EQ_A1(vx, vela2, rho);
				vela2 = vy*vy; eq[2] = EQ_A0(vy, vela2, rho); eq[3] = EQ_A1(vy, vela2, rho);
				vv = vx+vy; vela2 = vv*vv; eq[4] = EQ_4(vv, vela2, rho); eq[5] = EQ_5(vv, vela2, rho);
				vv = vx-vy; vela2 = vv*vv; eq[6] = EQ_4(vv, vela2, rho); eq[7] = EQ_5(vv, vela2, rho);
				vv = vx+vz; vela2 = vv*vv; eq[8] = EQ_4(vv, vela2, rho); eq[9] = EQ_5(vv, vela2, rho);
				vv = vx-vz; vela2 = vv*vv; eq[10] = EQ_4(vv, vela2, rho); eq[11] = EQ_5(vv, vela2, rho);
				vv = vy+vz; vela2 = vv*vv; eq[12] = EQ_4(vv, vela2, rho); eq[13] = EQ_5(vv, vela2, rho);
				vv = vy-vz; vela2 = vv*vv; eq[14] = EQ_4(vv, vela2, rho); eq[15] = EQ_5(vv, vela2, rho);
				vela2 = vz*vz; eq[16] = EQ_A0(vz, vela2, rho); eq[17] = EQ_A1(vz, vela2, rho);
				eq[18] = EQ_18(rho);
				w = FN(smag_inv_tau)(d, eq, rho_sum, tau, smag_k);
			}
			vela2 = vx*vx;
			d[0] += w*(EQ_A0(vx, vela2, rho) - d[0]);
			d[1] += w*(EQ_A1(vx, vela2, rho) - d[1]);
			vela2 = vy*vy;
			d[2] += w*(EQ_A0(vy, vela2, rho) - d[2]);
			d[3] += w*(EQ_A1(vy, vela2, rho) - d[3]);
			vv = vx+vy; vela2 = vv*vv;
			d[4] += w*(EQ_4(vv, vela2, rho) - d[4]);
			d[5] += w*(EQ_5(vv, vela2, rho) - d[5]);
			vv = vx-vy; vela2 = vv*vv;
			d[6] += w*(EQ_4(vv, vela2, rho) - d[6]);
			d[7] += w*(EQ_5(vv, vela2, rho) - d[7]);
			vv = vx+vz; vela2 = vv*vv;
			d[8] += w*(EQ_4(vv, vela2, rho) - d[8]);
			d[9] += w*(EQ_5(vv, vela2, rho) - d[9]);
			vv = vx-vz; vela2 = vv*vv;
			d[10] += w*(EQ_4(vv, vela2, rho) - d[10]);
			d[11] += w*(EQ_5(vv, vela2, rho) - d[11]);
			vv = vy+vz; vela2 = vv*vv;
			d[12] += w*(EQ_4(vv, vela2, rho) - d[12]);
			d[13] += w*(EQ_5(vv, vela2, rho) - d[13]);
			vv = vy-vz; vela2 = vv*vv;
			d[14] += w*(EQ_4(vv, vela2, rho) - d[14]);
			d[15] += w*(EQ_5(vv, vela2, rho) - d[15]);
			vela2 = vz*vz;
			d[16] += w*(EQ_A0(vz, vela2, rho) - d[16]);
			d[17] += w*(EQ_A1(vz, vela2, rho) - d[17]);
			d[18] += w*(EQ_18(rho) - d[18]);
			break;
		}
		case LBMO_FLAG_OBSTACLE:                                     /* :561-580 bounce back */
			vx = 0.0f; vy = 0.0f; vz = 0.0f;
			for (i = 0; i < 18; i += 2) { t = d[i+1]; d[i+1] = d[i]; d[i] = t; }
			break;
		case LBMO_FLAG_LID:                                          /* :582-653 */
			vx = u_lid; vy = 0; vz = 0;
			rho = 1.0f;
			vel2 = vx*vx + vy*vy + vz*vz;
			rho = rho - (T)(3.0f/2.0f)*(vel2);
			vela2 = vx*vx;
			d[0] = EQ_A0(vx, vela2, rho); d[1] = EQ_A1(vx, vela2, rho);
			vela2 = vy*vy;
			d[2] = EQ_A0(vy, vela2, rho); d[3] = EQ_A1(vy, vela2, rho);
			vv = vx+vy; vela2 = vv*vv;
			d[4] = EQ_4(vv, vela2, rho); d[5] = EQ_5(vv, vela2, rho);
			vv = vx-vy; vela2 = vv*vv;
			d[6] = EQ_4(vv, vela2, rho); d[7] = EQ_5(vv, vela2, rho);
			vv = vx+vz; vela2 = vv*vv;
			d[8] = EQ_4(vv, vela2, rho); d[9] = EQ_5(vv, vela2, rho);
			vv = vx-vz; vela2 = vv*vv;
			d[10] = EQ_4(vv, vela2, rho); d[11] = EQ_5(vv, vela2, rho);
			vv = vy+vz; vela2 = vv*vv;
			d[12] = EQ_4(vv, vela2, rho); d[13] = EQ_5(vv, vela2, rho);
			vv = vy-vz; vela2 = vv*vv;
			d[14] = EQ_4(vv, vela2, rho); d[15] = EQ_5(vv, vela2, rho);
			vela2 = vz*vz;
			d[16] = EQ_A0(vz, vela2, rho); d[17] = EQ_A1(vz, vela2, rho);
			d[18] = EQ_18(rho);
			break;
		default:                                                     /* ghost layer: pass through, :654-655 */
			break;
		}
		for (i = 0; i < 18; i++) dd[(long)i*n + L[i]] = d[i];          /* :712-754 / :758-784 */
		dd[18*n + gid] = d[18];
		if (flag == LBMO_FLAG_GHOST) continue;                        /* :787-788 */
		if (store_velocity) { velocity[gid] = vx; velocity[n + gid] = vy; velocity[2*n + gid] = vz; }
		if (store_density) density[gid] = rho;
	}
}

/* ------------------------------------------------------------------------------------
 * copy_buffer_rect, copy_buffer_rect.cl:12-52 (64-bit offsets here; the reference's int
 * offsets overflow for f*N at 512^3).
 */
void FN(lbmo_copy_rect)(const T *src, long src_off, const int so[3], const int ss[3],
		T *dst, long dst_off, const int dorg[3], const int ds[3], const int block[3])
{
	int i, j, k;
	for (k = 0; k < block[2]; k++)
		for (j = 0; j < block[1]; j++) {
			const long s0 = src_off + so[0] + (long)(so[1] + j) * ss[0] + (long)(so[2] + k) * ss[0] * ss[1];
			const long d0 = dst_off + dorg[0] + (long)(dorg[1] + j) * ds[0] + (long)(dorg[2] + k) * ds[0] * ds[1];
			for (i = 0; i < block[0]; i++) dst[d0 + i] = src[s0 + i];
		}
}

/* CLbmSolver::getVelocityChecksum, src/CLbmSolver.hpp:1103-1123: float accumulator,
 * index order, FLUID cells only. */
float FN(lbmo_checksum)(const T *velocity, const int *flags, long n)
{
	float checksum = 0;
	long a;
	for (a = 0; a < n; a++)
		if (flags[a] == LBMO_FLAG_FLUID)
			checksum += velocity[a] + velocity[n + a] + velocity[2*n + a];
	return checksum;
}

#undef W18
#undef W36
#undef W3
#undef EQ_A0
#undef EQ_A1
#undef EQ_4
#undef EQ_5
#undef EQ_18
