/* A plain C99 consumer of include/lbm_b200.h: what a cgo / JNI / ctypes-free binding sees.
 * Links against liblbm_b200.so, walks the no-compute entry points and prints what it observed.
 * With a GPU present it creates a 16^3 solver, steps it twice and destroys it. */
#include <stdio.h>
#include <string.h>
#include "../../include/lbm_b200.h"

int main(void)
{
	int ndev = -1, rc;
	lbm_t h = NULL;
	lbm_desc d;
	uint32_t mask = 0;
	const int dir[3] = { 1, 0, 0 };
	printf("version %d\n", lbmGetVersion());
	rc = lbmGetDeviceCount(&ndev);
	printf("devices rc=%d n=%d\n", rc, ndev);
	rc = lbmHaloSlotMask(LBM_SYNC_ALPHA, dir, LBM_HALO_SLOTS_MINIMAL, &mask);
	printf("mask rc=%d 0x%05x\n", rc, (unsigned)mask);
	memset(&d, 0, sizeof d);
	d.struct_size = sizeof d;
	d.dtype = LBM_F32;
	d.size[0] = d.size[1] = d.size[2] = 16;
	for (rc = 0; rc < 6; rc++) d.bc[rc] = LBM_FLAG_OBSTACLE;
	d.inv_tau = 1.5; d.tau = 1.0 / 1.5;
	rc = lbmCreate(&h, &d);
	printf("create rc=%d handle=%s\n", rc, h ? "yes" : "no");
	if (rc != LBM_OK) {
		printf("error: %s\n", lbmGetLastErrorString(NULL));
		printf("null-handle step rc=%d\n", lbmStep(NULL));
		return 0;
	}
	rc = lbmStep(h); rc |= lbmStep(h); rc |= lbmWait(h);
	{
		uint64_t c = 0;
		lbmGetStepCounter(h, &c);
		printf("stepped rc=%d counter=%llu\n", rc, (unsigned long long)c);
	}
	printf("destroy rc=%d\n", lbmDestroy(h));
	return 0;
}
