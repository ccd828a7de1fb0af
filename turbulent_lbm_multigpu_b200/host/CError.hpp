/*
 * CError.hpp -- the error convention of the solver facade: `error << "text"; if (error()) ...`
 * (reference src/lib/CError.hpp; consumer check at src/CController.hpp:218-221).
 */
#ifndef LBM_B200_HOST_CERROR_HPP
#define LBM_B200_HOST_CERROR_HPP

#include <ostream>
#include <sstream>
#include <string>

class CError {
	std::ostringstream _text;
	bool _set = false;

public:
	CError() {}
	CError(const CError &o) : _set(o._set) { _text << o._text.str(); }
	CError &operator=(const CError &o) { _text.str(o._text.str()); _set = o._set; return *this; }

	template <typename V>
	CError &operator<<(const V &v) { _text << v; _set = true; return *this; }
	CError &operator<<(std::ostream &(*manip)(std::ostream &)) { _text << manip; _set = true; return *this; }

	bool operator()() const { return _set; }           /* true: an error was recorded */
	std::string getString() { std::string s = _text.str(); _text.str(""); _set = false; return s; }
	std::string peek() const { return _text.str(); }
};

#endif
