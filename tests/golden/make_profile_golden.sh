#!/bin/bash
# Regenerates tests/golden/profile_8_5.ini with the REFERENCE's own profiler classes
# (/root/reference/src/libtools/CProfilerEvent.hpp, CProfiler.hpp, CProfilerEvent.cpp, compiled
# where they lie) fed with tests/cpp/profile_events.h through a mock clGetEventProfilingInfo, and
# the [METADATA] block written exactly as src/CController.hpp:503-519 does.  Only this container
# has /root/reference.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=/root/reference/src
TMP=$(mktemp -d)
cat > $TMP/main.cpp <<'EOC'
#include <cstring>
#include <CL/cl.h>
struct _cl_event { cl_ulong start, end; };
extern "C" cl_int clGetEventProfilingInfo(cl_event e, cl_profiling_info what, size_t n, void *out, size_t *)
{
	cl_ulong v = what == CL_PROFILING_COMMAND_START ? e->start : e->end;
	std::memcpy(out, &v, n);
	return CL_SUCCESS;
}
#define CL_CHECK_ERROR(x) (x)
#include "libtools/CProfiler.hpp"
#include "profile_events.h"
int main(int, char **argv)
{
	CProfiler prof;
	for (int i = 0; i < kProfileEventCount; i++) {
		_cl_event ev = { kProfileEvents[i].start, kProfileEvents[i].end };
		cl_event e = &ev;
		prof.addProfilerEvent(new CProfilerEvent(kProfileEvents[i].name, &e));
	}
	/* src/CController.hpp:503-519 */
	std::ofstream prof_file(argv[1], std::ios::out | std::ios::app);
	prof_file << "[METADATA]" << std::endl;
	prof_file << "TOTAL_NUM_PROC : " << PROFILE_TEST_NPROC << std::endl;
	prof_file << "CURRENT_PROC_ID : " << PROFILE_TEST_UID << std::endl;
	prof_file << std::endl;
	prof_file.close();
	prof.saveEvents(argv[1]);
	return 0;
}
EOC
g++ -O1 -w -I$REF -I$REF/include -I$HERE/../cpp $TMP/main.cpp $REF/libtools/CProfilerEvent.cpp -o $TMP/refprof
rm -f $HERE/profile_8_5.ini
$TMP/refprof $HERE/profile_8_5.ini
rm -rf $TMP
ls -la $HERE/profile_8_5.ini
