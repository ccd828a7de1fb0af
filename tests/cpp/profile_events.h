/* the event list both profile writers are fed with (tests/golden/make_profile_golden.sh feeds the
 * REFERENCE's CProfiler, tests/cpp/profile_writer.cpp this repository's): name, start ns, end ns */
#ifndef TEST_PROFILE_EVENTS_H
#define TEST_PROFILE_EVENTS_H
static const struct { const char *name; unsigned long long start, end; } kProfileEvents[] = {
	{ "init_kernel", 0ull, 187392ull },
	{ "copy_buffer_rect", 201000ull, 205984ull },
	{ "lbm_kernel_beta", 1000000ull, 1428331ull },
	{ "lbm_kernel_alpha", 1430000ull, 1838100ull },
	{ "lbm_kernel_beta", 1838200ull, 123456789012ull },
	{ "copy_buffer_rect", 123456789012ull, 123456789013ull },
	{ "lbm_kernel_alpha", 200000000000ull, 200000000000ull },
};
static const int kProfileEventCount = (int)(sizeof(kProfileEvents) / sizeof(kProfileEvents[0]));
#define PROFILE_TEST_NPROC 8
#define PROFILE_TEST_UID 5
#endif
