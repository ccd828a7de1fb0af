/*
 * CComm.hpp -- one halo face of a sub-domain: neighbour rank, the rect that is sent, the rect
 * that is received into, and the unit normal pointing INTO this sub-domain.  Interface of the
 * reference's src/CComm.hpp:8-79; the values come from CManager::initSimulation
 * (src/CManager.hpp:122-199) and are passed unchanged to lbmCommAddFace (include/lbm_b200.h).
 */
#ifndef LBM_B200_HOST_CCOMM_HPP
#define LBM_B200_HOST_CCOMM_HPP

#include "CVector.hpp"

template <typename T>
class CComm {
	typedef CVector<3, int> V;
	int _dst;
	V _ssize, _rsize, _sorigin, _rorigin, _dir;

public:
	CComm(int dstID, V send_size, V recv_size, V send_origin, V recv_origin, V comm_direction)
		: _dst(dstID), _ssize(send_size), _rsize(recv_size), _sorigin(send_origin), _rorigin(recv_origin),
		  _dir(comm_direction) {}

	int getDstId() const { return _dst; }
	V getSendSize() const { return _ssize; }
	V getRecvSize() const { return _rsize; }
	V getSendOrigin() const { return _sorigin; }
	V getRecvOrigin() const { return _rorigin; }
	V getCommDirection() const { return _dir; }

	void setDstId(int v) { _dst = v; }
	void setSendSize(V v) { _ssize = v; }
	void setRecvSize(V v) { _rsize = v; }
	void setSendOrigin(V v) { _sorigin = v; }
	void setRecvOrigin(V v) { _rorigin = v; }
	void setCommDirection(V v) { _dir = v; }

	/* axis (0,1,2) the face is normal to */
	int axis() const { return _dir[0] != 0 ? 0 : (_dir[1] != 0 ? 1 : 2); }
};

#endif
