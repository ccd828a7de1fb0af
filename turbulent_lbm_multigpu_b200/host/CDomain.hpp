/*
 * CDomain.hpp -- descriptor of a (sub-)domain: id, size in cells INCLUDING ghost layers,
 * origin inside the global grid, physical edge lengths.  Interface of the reference's
 * src/CDomain.hpp:17-57 (getOrigin/getSize/getUid/getLength, both constructors).
 */
#ifndef LBM_B200_HOST_CDOMAIN_HPP
#define LBM_B200_HOST_CDOMAIN_HPP

#include "CVector.hpp"

template <typename T>
class CDomain {
	int _uid;
	CVector<3, int> _cells, _origin;
	CVector<3, T> _edge;

public:
	CDomain(int UID, CVector<3, int> size, CVector<3, int> origin_cell, CVector<3, T> length)
		: _uid(UID), _cells(size), _origin(origin_cell), _edge(length) {}
	/* default edge length 0.05 m per axis, origin 0 (src/CDomain.hpp:30-34) */
	CDomain(int UID, CVector<3, int> size)
		: _uid(UID), _cells(size), _origin(0, 0, 0), _edge((T)0.05, (T)0.05, (T)0.05) {}

	CVector<3, int> getOrigin() const { return _origin; }
	CVector<3, int> getSize() const { return _cells; }
	int getUid() const { return _uid; }
	CVector<3, T> getLength() const { return _edge; }
};

#endif
