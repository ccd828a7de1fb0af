"""Text dumps of the device arrays in the reference's format (src/CLbmSolver.hpp:981-1100:
debugChar, debugFloat, debug_print, debugDD).  Pure formatting on host arrays; the Python twin of
host/CLbmDebug.hpp (tests/test_host_logic.py compares the two byte for byte)."""
from __future__ import annotations

import numpy as np


def _wrapped(items, wrap):
    out = []
    for i, t in enumerate(items):
        if i % wrap == 0:
            out.append("\n%d: " % (i // wrap))
        out.append(t + " ")
    return "".join(out)


def debugFloat(values, wrap_size=20):
    """src/CLbmSolver.hpp:1000-1029 (precision 4, fixed)."""
    return _wrapped(("%.4f" % float(v) for v in np.asarray(values).ravel()), wrap_size)


def debugChar(values, wrap_size=20):
    """src/CLbmSolver.hpp:981-998: every BYTE as a signed integer."""
    raw = np.ascontiguousarray(values).view(np.int8).ravel()
    return _wrapped((str(int(v)) for v in raw), wrap_size)


def debug_print(dd, velocity, density, flags):
    """src/CLbmSolver.hpp:1032-1057."""
    return ("DENSITY DISTRIBUTIONS:" + debugFloat(dd, 16) + "\n"
            + "\nVELOCITY:" + debugFloat(velocity, 4 * 3) + "\n"
            + "\nDENSITY:" + debugFloat(density, 4) + "\n"
            + "\nFLAGS:" + debugChar(np.asarray(flags, np.int32), 4 * 4) + "\n")


def debugDD(dd, cells, dd_id=0, wrap_size=16, empty_line=16):
    """src/CLbmSolver.hpp:1062-1101."""
    dd = np.asarray(dd).ravel()
    start, end = cells * dd_id, cells * (dd_id + 1)
    out = []
    for i in range(start, end):
        if empty_line != wrap_size and i % empty_line == 0 and i != start:
            out.append("\n")
        if i % wrap_size == 0:
            if i != start:
                out.append("\n")
            out.append("%d: " % (i // wrap_size))
        out.append("%.4f " % float(dd[i]))
    out.append("\n")
    return "".join(out)
