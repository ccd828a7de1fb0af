/*
 * CLbmDebug.hpp -- text dumps of the device arrays in the reference's format
 * (src/CLbmSolver.hpp:981-1100: debugChar, debugFloat, debug_print, debugDD).  Free functions on
 * host arrays so that the formatting is testable without a device; CLbmSolver<T>::debug_print /
 * debugDD download the arrays through the C ABI and call these.
 */
#ifndef LBM_B200_HOST_CLBMDEBUG_HPP
#define LBM_B200_HOST_CLBMDEBUG_HPP

#include <cstddef>
#include <iomanip>
#include <iostream>

namespace lbm_debug {

/* RAII for the stream state the reference sets and restores (precision 4, fixed) */
struct FixedFour {
	std::ostream &os;
	std::streamsize precision;
	explicit FixedFour(std::ostream &o) : os(o), precision(o.precision())
	{
		os.precision(4);
		os.setf(std::ios::fixed, std::ios::floatfield);
	}
	~FixedFour()
	{
		os.precision(precision);
		os << std::resetiosflags(std::ios::fixed);
	}
};

/* rows of `wrap` entries, each opened by "\n<row>: " (src/CLbmSolver.hpp:988-995, 1013-1021) */
template <typename V, typename Shown>
void wrapped(std::ostream &os, const V *v, size_t count, size_t wrap)
{
	for (size_t i = 0; i < count; i++) {
		if (i % wrap == 0) os << std::endl << (i / wrap) << ": ";
		os << (Shown)v[i] << " ";
	}
}

/* debugFloat: every element of a T array */
template <typename T>
void debugFloat(std::ostream &os, const T *v, size_t count, size_t wrap = 20)
{
	FixedFour f(os);
	wrapped<T, T>(os, v, count, wrap);
}

/* debugChar: every BYTE of the array as an integer (the reference dumps the int32 flags this
 * way, so each flag shows up as four numbers) */
inline void debugChar(std::ostream &os, const void *bytes, size_t byte_count, size_t wrap = 20)
{
	wrapped<char, int>(os, (const char *)bytes, byte_count, wrap);
}

/* debug_print body (src/CLbmSolver.hpp:1032-1057); dd: 19*cells, velocity: 3*cells */
template <typename T>
void debugPrint(std::ostream &os, const T *dd, const T *velocity, const T *density, const int *flags, size_t cells)
{
	os << "DENSITY DISTRIBUTIONS:";
	debugFloat(os, dd, 19 * cells, 16);
	os << std::endl;
	os << std::endl << "VELOCITY:";
	debugFloat(os, velocity, 3 * cells, 4 * 3);
	os << std::endl;
	os << std::endl << "DENSITY:";
	debugFloat(os, density, cells, 4);
	os << std::endl;
	os << std::endl << "FLAGS:";
	debugChar(os, flags, cells * sizeof(int), 4 * 4);
	os << std::endl;
}

/* debugDD (src/CLbmSolver.hpp:1062-1101): one slot of the population array; row labels count
 * from the start of the whole array, a blank line every `empty_line` values */
template <typename T>
void debugDD(std::ostream &os, const T *dd, size_t cells, size_t dd_id = 0, size_t wrap_size = 16, size_t empty_line = 16)
{
	{
		FixedFour f(os);
		const size_t start = cells * dd_id, end = cells * (dd_id + 1);
		for (size_t i = start; i < end; i++) {
			if (empty_line != wrap_size && i % empty_line == 0 && i != start) os << std::endl;
			if (i % wrap_size == 0) {
				if (i != start) os << std::endl;
				os << (i / wrap_size) << ": ";
			}
			os << dd[i] << " ";
		}
	}
	os << std::endl;
}

}  /* namespace lbm_debug */

#endif
