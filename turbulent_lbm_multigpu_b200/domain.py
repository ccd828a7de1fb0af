"""CDomain / CComm value types (reference src/CDomain.hpp:17-57, src/CComm.hpp:8-79)."""
from __future__ import annotations

from dataclasses import dataclass


def _v3(v):
    v = tuple(v)
    if len(v) != 3:
        raise ValueError("expected 3 components")
    return v


class CDomain:
    """{uid, size, origin_cell, length} of a (sub-)domain; size includes ghost layers."""

    def __init__(self, UID, size, origin_cell=(0, 0, 0), length=(0.05, 0.05, 0.05)):
        self._UID = int(UID)
        self._size = tuple(int(s) for s in _v3(size))
        self._origin_cell = tuple(int(s) for s in _v3(origin_cell))
        self._length = _v3(length)

    def getOrigin(self):
        return self._origin_cell

    def getSize(self):
        return self._size

    def getUid(self):
        return self._UID

    def getLength(self):
        return self._length


@dataclass
class CComm:
    """Halo descriptor: destination rank, send/recv rect and the unit normal that points
    INTO this sub-domain (values: reference src/CManager.hpp:122-199)."""
    _dstID: int
    _send_size: tuple
    _recv_size: tuple
    _send_origin: tuple
    _recv_origin: tuple
    _comm_direction: tuple

    def getDstId(self):
        return self._dstID

    def setDstId(self, v):
        self._dstID = int(v)

    def getSendSize(self):
        return self._send_size

    def setSendSize(self, v):
        self._send_size = _v3(v)

    def getRecvSize(self):
        return self._recv_size

    def setRecvSize(self, v):
        self._recv_size = _v3(v)

    def getSendOrigin(self):
        return self._send_origin

    def setSendOrigin(self, v):
        self._send_origin = _v3(v)

    def getRecvOrigin(self):
        return self._recv_origin

    def setRecvOrigin(self, v):
        self._recv_origin = _v3(v)

    def getCommDirection(self):
        return self._comm_direction

    def setCommDirection(self, v):
        self._comm_direction = _v3(v)

    @property
    def axis(self):
        return next(a for a in range(3) if self._comm_direction[a] != 0)

    def as_tuple(self):
        return (self._dstID, tuple(self._send_size), tuple(self._recv_size), tuple(self._send_origin),
                tuple(self._recv_origin), tuple(self._comm_direction))
