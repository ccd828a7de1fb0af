/* tests/test_profiler.py: this repository's CProfiler fed with tests/cpp/profile_events.h; the
 * output must be byte-identical to tests/golden/profile_8_5.ini (the reference's writer). */
#include "../../turbulent_lbm_multigpu_b200/host/CProfiler.hpp"
#include "profile_events.h"

/* the C ABI is not linked here: the writer alone is under test */
extern "C" {
int lbmProfileEventCount(lbm_t, uint64_t *, uint64_t *) { return 1; }
int lbmProfileGetEvent(lbm_t, uint64_t, char *, size_t, uint64_t *, uint64_t *) { return 1; }
int lbmProfileClear(lbm_t) { return 1; }
}

int main(int argc, char **argv)
{
	if (argc < 2) return 2;
	CProfiler p;
	for (int i = 0; i < kProfileEventCount; i++)
		p.addDeviceKernel(kProfileEvents[i].name, kProfileEvents[i].start, kProfileEvents[i].end);
	p.saveProfile(argv[1], PROFILE_TEST_NPROC, PROFILE_TEST_UID);
	std::cout << p.size() << " " << p.countOverlapping() << std::endl;
	return 0;
}
