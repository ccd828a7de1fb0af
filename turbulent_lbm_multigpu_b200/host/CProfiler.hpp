/*
 * CProfiler.hpp -- per-kernel device timeline of one rank, written in the schema the reference's
 * profile.py reads (reference src/libtools/CProfilerEvent.hpp:17-93, src/libtools/CProfiler.hpp:16-50,
 * file header src/CController.hpp:503-519).
 *
 * The reference builds with PROFILE=1, enables CL_QUEUE_PROFILING_ENABLE, waits after every
 * enqueue and turns each cl_event into a CProfilerEvent.  Here the events come from the C ABI
 * (lbmProfileEnable / lbmProfileGetEvent: CUDA events around every launch, nothing blocks), the
 * switch is the run-time variable LBM_B200_PROFILE, and ranks are threads, so there is one
 * CProfiler per controller instead of a process-wide singleton.
 */
#ifndef LBM_B200_HOST_CPROFILER_HPP
#define LBM_B200_HOST_CPROFILER_HPP

#include <cstdint>
#include <fstream>
#include <iostream>
#include <map>
#include <string>

#include "../../include/lbm_b200.h"

typedef enum {
	EVENT_TYPE_UNKNOWN = 0,
	EVENT_TYPE_DEVICE_KERNEL,
	EVENT_TYPE_HOST_FUNCTION,
} EVENT_TYPE;

class CProfilerEvent
{
	int _uuid;
	EVENT_TYPE _type;
	std::string _event_id;
	uint64_t _start;   /* nanoseconds */
	uint64_t _end;     /* nanoseconds */
	float _duration;   /* milliseconds */

public:
	CProfilerEvent(int uuid, const std::string &event_id, uint64_t start_ns, uint64_t end_ns,
			EVENT_TYPE type = EVENT_TYPE_DEVICE_KERNEL)
		: _uuid(uuid), _type(type), _event_id(event_id), _start(start_ns), _end(end_ns),
		  _duration((float)((end_ns - start_ns) / 1000000.0))
	{
		if (event_id.empty()) throw "CProfilerEvent: ID of the event is unknown!";
	}

	int getUuid() const { return _uuid; }
	std::string getEventId() const { return _event_id; }
	uint64_t getEventStartTime() const { return _start; }
	uint64_t getEventEndTime() const { return _end; }
	float getEventDuration() const { return _duration; }
	bool overlap(const CProfilerEvent &o) const { return _start < o._end && _end > o._start; }

	void printEvent(std::ostream &prof_file) const
	{
		prof_file << "[EVENT" << _uuid << "]" << std::endl;
		prof_file << "TYPE : " << ((_type == EVENT_TYPE_DEVICE_KERNEL) ? "DEVICE_KERNEL" : "HOST_FUNCTION") << std::endl;
		prof_file << "NAME : " << _event_id << std::endl;
		prof_file << "# start/end in nanoseconds" << std::endl;
		prof_file << "START : " << _start << std::endl;
		prof_file << "END : " << _end << std::endl;
		prof_file << "# duration in milliseconds" << std::endl;
		prof_file << "DURATION : " << _duration << std::endl;
		prof_file << std::endl;
	}
};

class CProfiler
{
	typedef std::multimap<int, CProfilerEvent *> event_map;
	event_map _event_container;
	int _event_counter;

public:
	CProfiler() : _event_counter(0) {}
	~CProfiler() { clear(); }
	CProfiler(const CProfiler &) = delete;
	CProfiler &operator=(const CProfiler &) = delete;

	void clear()
	{
		for (event_map::iterator it = _event_container.begin(); it != _event_container.end(); ++it) delete it->second;
		_event_container.clear();
		_event_counter = 0;
	}

	void addProfilerEvent(CProfilerEvent *e) { _event_container.insert(event_map::value_type(e->getUuid(), e)); }

	/* uuids count from 1 like the reference's static event_counter */
	void addDeviceKernel(const std::string &name, uint64_t start_ns, uint64_t end_ns)
	{
		addProfilerEvent(new CProfilerEvent(++_event_counter, name, start_ns, end_ns));
	}

	/* drain the timeline a solver handle has recorded since lbmProfileEnable / lbmProfileClear */
	int collect(lbm_t handle)
	{
		uint64_t n = 0;
		if (int rc = lbmProfileEventCount(handle, &n, NULL)) return rc;
		for (uint64_t i = 0; i < n; i++) {
			char name[128];
			uint64_t t0 = 0, t1 = 0;
			if (int rc = lbmProfileGetEvent(handle, i, name, sizeof(name), &t0, &t1)) return rc;
			addDeviceKernel(name, t0, t1);
		}
		return lbmProfileClear(handle);
	}

	size_t size() const { return _event_container.size(); }

	size_t countOverlapping() const
	{
		size_t c = 0;
		for (event_map::const_iterator a = _event_container.begin(); a != _event_container.end(); ++a) {
			event_map::const_iterator b = a;
			for (++b; b != _event_container.end(); ++b)
				if (a->second->overlap(*b->second)) c++;
		}
		return c;
	}

	void saveEvents(const std::string &file_name) const
	{
		std::ofstream prof_file(file_name.c_str(), std::ios::out | std::ios::app);
		if (prof_file.is_open()) {
			for (event_map::const_iterator it = _event_container.begin(); it != _event_container.end(); ++it)
				it->second->printEvent(prof_file);
		} else std::cout << "Unable to open file: " << file_name << std::endl;
	}

	/* the whole file of one rank: [METADATA] block of src/CController.hpp:511-516, then the events */
	void saveProfile(const std::string &file_name, int total_num_proc, int current_proc_id) const
	{
		{
			std::ofstream prof_file(file_name.c_str(), std::ios::out | std::ios::app);
			if (prof_file.is_open()) {
				prof_file << "[METADATA]" << std::endl;
				prof_file << "TOTAL_NUM_PROC : " << total_num_proc << std::endl;
				prof_file << "CURRENT_PROC_ID : " << current_proc_id << std::endl;
				prof_file << std::endl;
			} else std::cout << "Unable to open file: " << file_name << std::endl;
		}
		saveEvents(file_name);
	}
};

#endif
