/*
 * CRankWorld.hpp -- what replaces MPI_COMM_WORLD: the ranks of a run are host threads of ONE
 * process, one per sub-domain / GPU (the image has no MPI; all 8 GPUs of a B200 box are
 * peers of one process).  It provides exactly what the reference takes from MPI on this path
 * (src/CController.hpp:299-311,361-373,449; src/main.cpp:338-352):
 *   barrier()            rendezvous of all ranks
 *   send()/recv()        tagged point-to-point of host buffers (host-staged sync mode)
 *   reduceMax()          MPI_Reduce(MAX) of the wall time
 *   publish()/lookup     registry of solver handles and face ids, so that neighbouring
 *                        sub-domains can map each other's halo blocks (lbmCommConnectLocal)
 */
#ifndef LBM_B200_HOST_CRANKWORLD_HPP
#define LBM_B200_HOST_CRANKWORLD_HPP

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/lbm_b200.h"

class CRankWorld {
	int _n;
	std::mutex _m;
	std::condition_variable _cv;
	int _arrived;
	unsigned long _generation;
	std::map<std::tuple<int, int, int>, std::deque<std::vector<char> > > _mail;   /* (src,dst,tag) */
	std::vector<lbm_t> _handles;
	std::map<std::tuple<int, int, int, int>, int> _faces;                       /* (rank,dst,axis,sign) -> face id */
	bool _failed;

public:
	std::mutex print_mutex;    /* keeps the per-rank stdout blocks readable */

	explicit CRankWorld(int nranks) : _n(nranks), _arrived(0), _generation(0), _handles(nranks, (lbm_t)0),
		_failed(false) {}

	int size() const { return _n; }

	/* a rank that cannot continue marks the world failed so that nobody waits for it forever */
	void fail() { std::lock_guard<std::mutex> g(_m); _failed = true; _cv.notify_all(); }
	bool failed() { std::lock_guard<std::mutex> g(_m); return _failed; }

	bool barrier()
	{
		std::unique_lock<std::mutex> g(_m);
		if (_failed) return false;
		const unsigned long gen = _generation;
		if (++_arrived == _n) { _arrived = 0; _generation++; _cv.notify_all(); return true; }
		_cv.wait(g, [&] { return _generation != gen || _failed; });
		return !_failed;
	}

	void send(int src, int dst, int tag, const void *buf, size_t bytes)
	{
		std::vector<char> msg((const char *)buf, (const char *)buf + bytes);
		std::lock_guard<std::mutex> g(_m);
		_mail[std::make_tuple(src, dst, tag)].push_back(std::move(msg));
		_cv.notify_all();
	}

	bool recv(int src, int dst, int tag, void *buf, size_t bytes)
	{
		std::unique_lock<std::mutex> g(_m);
		std::deque<std::vector<char> > &q = _mail[std::make_tuple(src, dst, tag)];
		_cv.wait(g, [&] { return !q.empty() || _failed; });
		if (q.empty()) return false;
		std::memcpy(buf, q.front().data(), bytes < q.front().size() ? bytes : q.front().size());
		q.pop_front();
		return true;
	}

	/* every rank calls it and gets the maximum (MPI_Reduce(MAX) of src/CController.hpp:449) */
	double reduceMax(double v)
	{
		{ std::lock_guard<std::mutex> g(_m); _contrib.push_back(v); }
		barrier();                              /* everyone has contributed */
		double m;
		{
			std::lock_guard<std::mutex> g(_m);
			m = _contrib[0];
			for (size_t i = 1; i < _contrib.size(); i++) if (_contrib[i] > m) m = _contrib[i];
		}
		barrier();                              /* everyone has read */
		{ std::lock_guard<std::mutex> g(_m); _contrib.clear(); }
		barrier();
		return m;
	}

	void publishHandle(int rank, lbm_t h) { std::lock_guard<std::mutex> g(_m); _handles[rank] = h; }
	lbm_t handle(int rank) { std::lock_guard<std::mutex> g(_m); return _handles[rank]; }
	void publishFace(int rank, int dst, int axis, int sign, int face_id)
	{
		std::lock_guard<std::mutex> g(_m);
		_faces[std::make_tuple(rank, dst, axis, sign)] = face_id;
	}
	int faceId(int rank, int dst, int axis, int sign)
	{
		std::lock_guard<std::mutex> g(_m);
		std::map<std::tuple<int, int, int, int>, int>::iterator it = _faces.find(std::make_tuple(rank, dst, axis, sign));
		return it == _faces.end() ? -1 : it->second;
	}

private:
	std::vector<double> _contrib;
};

#endif
