"""CProfiler / CProfilerEvent -- per-kernel device timeline of one rank in the schema the
reference's profile.py reads (reference src/libtools/CProfilerEvent.hpp:17-93,
src/libtools/CProfiler.hpp:16-50; file header src/CController.hpp:503-519).

The events come from the C ABI (lbmProfileEnable / lbmProfileGetEvent: CUDA events around every
kernel launch of a solver handle); this module only keeps and formats them.
"""
from __future__ import annotations

import os
import struct

EVENT_TYPE_UNKNOWN, EVENT_TYPE_DEVICE_KERNEL, EVENT_TYPE_HOST_FUNCTION = 0, 1, 2
PROFILE_OUTPUT_DIR = "output/profile"           # reference src/common.h:27


def _ostream_float(v):
    """std::ostream << cl_float at the default precision: the value rounded to binary32
    (CProfilerEvent.hpp:26,37), printed with 6 significant digits (%g)."""
    return "%g" % struct.unpack("f", struct.pack("f", float(v)))[0]


class CProfilerEvent:
    def __init__(self, uuid, event_id, start_ns, end_ns, event_type=EVENT_TYPE_DEVICE_KERNEL):
        if not event_id:
            raise ValueError("CProfilerEvent: ID of the event is unknown!")
        self._uuid, self._type, self._event_id = int(uuid), event_type, str(event_id)
        self._start, self._end = int(start_ns), int(end_ns)
        self._duration = (self._end - self._start) / 1000000.0      # milliseconds

    def getUuid(self):
        return self._uuid

    def getEventId(self):
        return self._event_id

    def getEventStartTime(self):
        return self._start

    def getEventEndTime(self):
        return self._end

    def getEventDuration(self):
        return self._duration

    def overlap(self, other):
        """profile.py:31-34"""
        return self._start < other._end and self._end > other._start

    def format(self):
        """CProfilerEvent::printEvent (CProfilerEvent.hpp:81-91)"""
        return ("[EVENT%d]\nTYPE : %s\nNAME : %s\n# start/end in nanoseconds\nSTART : %d\nEND : %d\n"
                "# duration in milliseconds\nDURATION : %s\n\n" % (
                    self._uuid, "DEVICE_KERNEL" if self._type == EVENT_TYPE_DEVICE_KERNEL else "HOST_FUNCTION",
                    self._event_id, self._start, self._end, _ostream_float(self._duration)))


class CProfiler:
    def __init__(self):
        self._events = []

    def clear(self):
        self._events = []

    def addProfilerEvent(self, event):
        self._events.append(event)

    def addDeviceKernel(self, name, start_ns, end_ns):
        self.addProfilerEvent(CProfilerEvent(len(self._events) + 1, name, start_ns, end_ns))

    def collect(self, solver):
        """drain the timeline the solver's handle has recorded (and give it a new time zero)"""
        for name, t0, t1 in solver.profileEvents():
            self.addDeviceKernel(name, t0, t1)
        solver.profileClear()

    def events(self):
        return list(self._events)

    def __len__(self):
        return len(self._events)

    def overlappingEvents(self):
        """profile.py:49-58"""
        ev = sorted(self._events, key=lambda e: e.getUuid())
        return [(a, b) for i, a in enumerate(ev) for b in ev[i + 1:] if a.overlap(b)]

    def saveEvents(self, file_name):
        with open(file_name, "a") as f:
            for e in sorted(self._events, key=lambda e: e.getUuid()):
                f.write(e.format())

    def saveProfile(self, file_name, total_num_proc, current_proc_id):
        d = os.path.dirname(file_name)
        if d:
            os.makedirs(d, exist_ok=True)
        with open(file_name, "a") as f:
            f.write("[METADATA]\nTOTAL_NUM_PROC : %d\nCURRENT_PROC_ID : %d\n\n" % (total_num_proc, current_proc_id))
        self.saveEvents(file_name)


def profile_file_name(total_num_proc, uid, base="."):
    return os.path.join(base, PROFILE_OUTPUT_DIR, "profile_%d_%d.ini" % (total_num_proc, uid))


__all__ = ["CProfiler", "CProfilerEvent", "profile_file_name", "PROFILE_OUTPUT_DIR"]
