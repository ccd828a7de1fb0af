/*
 * CController.hpp -- per-rank controller: owns the solver and the halo-face list, runs the
 * time loop.  Public surface of the reference's src/CController.hpp (run, computeNextStep,
 * syncAlpha, syncBeta, addCommunication, setGeometry, getSolver, getDomain, getUid).
 *
 * What changed underneath:
 *   - the OpenCL platform/context/queue bring-up (src/CController.hpp:83-226) is one lbmCreate;
 *   - a rank is a host thread of this process driving one GPU (CRankWorld instead of MPI);
 *   - syncAlpha/syncBeta no longer stage through the host.  Three interchangeable modes:
 *       SYNC_P2P   one-sided NVLink peer stores + device-side flags, fused with the split
 *                  (shell | interior) step in ONE library call per step (lbmCommStep): the
 *                  exchange hides under the interior kernel.  Needs one GPU per rank.
 *       SYNC_COPY  one peer-copy kernel per face (pack + transfer + unpack fused,
 *                  lbmHaloCopyPeer) between rank barriers; also valid when several ranks
 *                  share a GPU.
 *       SYNC_HOST  the reference's algorithm verbatim (storeDensityDistribution -> send/recv
 *                  of host buffers -> setDensityDistribution(+norm), one CComm after the
 *                  other), kept as the behavioural reference.
 */
#ifndef LBM_B200_HOST_CCONTROLLER_HPP
#define LBM_B200_HOST_CCONTROLLER_HPP

#include <sys/stat.h>

#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <list>
#include <sstream>
#include <string>
#include <vector>

#include "CComm.hpp"
#include "CConfiguration.hpp"
#include "CDomain.hpp"
#include "CLbmSolver.hpp"
#include "CLbmVisualizationVTK.hpp"
#include "CProfiler.hpp"
#include "CRankWorld.hpp"
#include "common.h"

enum LbmSyncMode { SYNC_AUTO = 0, SYNC_P2P = 1, SYNC_COPY = 2, SYNC_HOST = 3 };

template <typename T>
class CController {
	typedef Singleton<CConfiguration<T> > ConfigSingleton;

	int _UID;
	CDomain<T> _domain;
	int _BC[3][2];
	std::vector<CComm<T> *> _comm_container;
	std::vector<int> _face_ids;

	CCL::CContext cContext;
	CCL::CDevice cDevice;
	CCL::CCommandQueue cCommandQueue;
	CLbmSolver<T> *cLbmPtr;

	CRankWorld *_world;
	LbmSyncMode _sync;
	bool _connected;

	int initLBMSolver()
	{
		CConfiguration<T> *cfg = ConfigSingleton::Instance();
		const int ndev = CCL::CContext::deviceCount();
		if (ndev <= 0) {
			std::cerr << "no CUDA device available: liblbm_b200 has no CPU fallback" << std::endl;
			return -1;
		}
		if (cfg->device_nr < 0 || cfg->device_nr >= ndev) {
			std::cerr << "invalid device number - use option \"-d -1\" to list all devices" << std::endl;
			return -1;
		}
		/* one GPU per rank, starting at the configured device (the reference always takes
		 * device 0 of its node: one rank per node, src/CController.hpp:173-182) */
		cDevice = CCL::CDevice((cfg->device_nr + (_UID < 0 ? 0 : _UID)) % ndev);
		const bool store = cfg->do_visualization || cfg->debug_mode || cfg->do_validate;
		cLbmPtr = new CLbmSolver<T>(_UID, cCommandQueue, cContext, cDevice, _BC, _domain, cfg->gravitation,
				cfg->viscosity, cfg->computation_kernel_count, store, store, cfg->timestep,
				cfg->drivenCavityVelocity, cfg->lbm_opencl_number_of_threads_list,
				cfg->lbm_opencl_number_of_registers_list, cfg->smagorinsky_constant, beta_order);
		if (cLbmPtr->error()) {
			std::cout << cLbmPtr->error.getString();
			return -1;
		}
		cLbmPtr->wait();
		if (_world && _UID >= 0) _world->publishHandle(_UID, cLbmPtr->handle());
		return 0;
	}

	LbmSyncMode resolvedSync() const
	{
		if (_sync != SYNC_AUTO) return _sync;
		if (!_world || _world->size() <= 1) return SYNC_COPY;
		return _world->size() <= CCL::CContext::deviceCount() ? SYNC_P2P : SYNC_COPY;
	}

	/* P2P: register my faces, then map the neighbours' receive blocks (once, collectively) */
	void connectFaces()
	{
		if (_connected) return;
		_connected = true;
		if (_comm_container.empty() || resolvedSync() != SYNC_P2P) return;
		lbm_t h = cLbmPtr->handle();
		for (size_t i = 0; i < _comm_container.size(); i++) {
			CComm<T> *c = _comm_container[i];
			int fid = -1;
			CVector<3, int> so = c->getSendOrigin(), ro = c->getRecvOrigin(), sz = c->getSendSize(), dir = c->getCommDirection();
			check(lbmCommAddFace(h, c->getDstId(), so.data, ro.data, sz.data, dir.data, halo_slots, &fid));
			_face_ids.push_back(fid);
			_world->publishFace(_UID, c->getDstId(), c->axis(), dir[c->axis()], fid);
		}
		/* a decomposition that cuts x (then every rank has an x neighbour): z, y, x phase order --
		 * x faces after the interior kernel instead of an x shell (include/lbm_b200.h) */
		if (axis_order >= 0) check(lbmCommSetAxisOrder(h, axis_order));
		else if (!getenv("LBM_B200_AXIS_ORDER")) {
			bool cuts_x = false;
			for (size_t i = 0; i < _comm_container.size(); i++) cuts_x |= _comm_container[i]->axis() == 0;
			if (cuts_x) check(lbmCommSetAxisOrder(h, LBM_AXIS_ORDER_ZYX));
		}
		_world->barrier();
		for (size_t i = 0; i < _comm_container.size(); i++) {
			CComm<T> *c = _comm_container[i];
			const int a = c->axis(), sign = c->getCommDirection()[a];
			const int peer_fid = _world->faceId(c->getDstId(), _UID, a, -sign);
			check(lbmCommConnectLocal(h, _face_ids[i], _world->handle(c->getDstId()), peer_fid));
		}
		_world->barrier();
	}

	void check(int status)
	{
		if (status != LBM_OK) {
			cLbmPtr->error << "liblbm_b200 status " << status << ": " << lbmGetLastErrorString(cLbmPtr->handle()) << std::endl;
			std::cerr << cLbmPtr->error.peek();
			if (_world) _world->fail();
		}
	}

	/* SYNC_HOST: src/CController.hpp:265-320 (alpha) / :322-383 (beta) */
	void syncHost(bool beta)
	{
		for (size_t i = 0; i < _comm_container.size(); i++) {
			CComm<T> *c = _comm_container[i];
			CVector<3, int> send_size = beta ? c->getRecvSize() : c->getSendSize();
			CVector<3, int> recv_size = beta ? c->getSendSize() : c->getRecvSize();
			CVector<3, int> send_origin = beta ? c->getRecvOrigin() : c->getSendOrigin();
			CVector<3, int> recv_origin = beta ? c->getSendOrigin() : c->getRecvOrigin();
			const size_t ns = (size_t)send_size.elements() * cLbmPtr->SIZE_DD_HOST;
			const size_t nr = (size_t)recv_size.elements() * cLbmPtr->SIZE_DD_HOST;
			std::vector<T> send_buffer(ns), recv_buffer(nr);
			cLbmPtr->storeDensityDistribution(send_buffer.data(), send_origin, send_size);
			const int tag = beta ? MPI_TAG_BETA_SYNC : MPI_TAG_ALPHA_SYNC;
			_world->send(_UID, c->getDstId(), tag, send_buffer.data(), ns * sizeof(T));
			if (!_world->recv(c->getDstId(), _UID, tag, recv_buffer.data(), nr * sizeof(T))) return;
			if (beta) cLbmPtr->setDensityDistribution(recv_buffer.data(), recv_origin, recv_size, c->getCommDirection());
			else cLbmPtr->setDensityDistribution(recv_buffer.data(), recv_origin, recv_size);
			cLbmPtr->wait();
		}
	}

	/* SYNC_COPY: per axis phase, every rank copies its faces straight into the neighbour */
	void syncCopy(bool beta)
	{
		const int kind = beta ? LBM_SYNC_BETA : LBM_SYNC_ALPHA;
		CVector<3, int> S = _domain.getSize();
		for (int axis = 0; axis < 3; axis++) {
			cLbmPtr->wait();
			if (!_world->barrier()) return;            /* everyone finished the step / the previous phase */
			for (size_t i = 0; i < _comm_container.size(); i++) {
				CComm<T> *c = _comm_container[i];
				if (c->axis() != axis) continue;
				/* the neighbour's descriptor of the same face (src/CManager.hpp:122-199) */
				CVector<3, int> back_dir = c->getCommDirection() * -1;
				CVector<3, int> back_send(0, 0, 0), back_recv(0, 0, 0);
				back_send[axis] = back_dir[axis] > 0 ? 1 : S[axis] - 2;
				back_recv[axis] = back_dir[axis] > 0 ? 0 : S[axis] - 1;
				uint32_t mask = 0, minimal = 0;
				lbmHaloSlotMask(kind, back_dir.data, halo_slots, &mask);
				CVector<3, int> so, dorg, sz = c->getSendSize();
				if (beta) {         /* my ghost layer -> the neighbour's outermost real layer */
					lbmHaloSlotMask(kind, back_dir.data, LBM_HALO_SLOTS_MINIMAL, &minimal);
					mask &= minimal;
					so = c->getRecvOrigin(); dorg = back_send;
				} else {            /* my outermost real layer -> the neighbour's ghost layer */
					so = c->getSendOrigin(); dorg = back_recv;
				}
				check(lbmHaloCopyPeer(cLbmPtr->handle(), so.data, _world->handle(c->getDstId()), dorg.data, sz.data, mask, NULL));
			}
			cLbmPtr->wait();
		}
		_world->barrier();
	}

public:
	float vector_checksum;
	double seconds, mlups;          /* filled by run() */
	CProfiler profiler;             /* per rank (the reference's ProfilerSingleton is per MPI process) */
	int halo_slots;                 /* LBM_HALO_SLOTS_MINIMAL (5 per face) | _REFERENCE (19) */
	int axis_order;                 /* LBM_AXIS_ORDER_*; -1 (default): z,y,x when the decomposition cuts x */
	int beta_order;

	CController(int UID, CDomain<T> domain, int BC[3][2], CRankWorld *world = NULL, LbmSyncMode sync = SYNC_AUTO,
			int p_beta_order = LBM_BETA_ORDER_SHIPPED)
		: _UID(UID), _domain(domain), cLbmPtr(NULL), _world(world), _sync(sync), _connected(false),
		  vector_checksum(0), seconds(0), mlups(0), halo_slots(LBM_HALO_SLOTS_MINIMAL), axis_order(-1), beta_order(p_beta_order)
	{
		for (int a = 0; a < 3; a++) for (int s = 0; s < 2; s++) _BC[a][s] = BC[a][s];
		if (initLBMSolver() == -1) {
			if (_world) _world->fail();
			throw "Initialization of LBM Solver failed!";
		}
	}

	~CController()
	{
		delete cLbmPtr;
		for (size_t i = 0; i < _comm_container.size(); i++) delete _comm_container[i];
	}

	void syncAlpha()
	{
		connectFaces();
		if (_comm_container.empty() && resolvedSync() != SYNC_COPY) return;
		switch (resolvedSync()) {
		case SYNC_HOST: syncHost(false); break;
		case SYNC_P2P: lbmStreamWaitStream(cLbmPtr->handle(), 1); check(lbmCommSync(cLbmPtr->handle(), LBM_SYNC_ALPHA)); lbmStreamWaitStream(cLbmPtr->handle(), 0); break;
		default: if (_world && _world->size() > 1) syncCopy(false); break;
		}
	}

	void syncBeta()
	{
		connectFaces();
		if (_comm_container.empty() && resolvedSync() != SYNC_COPY) return;
		switch (resolvedSync()) {
		case SYNC_HOST: syncHost(true); break;
		case SYNC_P2P: lbmStreamWaitStream(cLbmPtr->handle(), 1); check(lbmCommSync(cLbmPtr->handle(), LBM_SYNC_BETA)); lbmStreamWaitStream(cLbmPtr->handle(), 0); break;
		default: if (_world && _world->size() > 1) syncCopy(true); break;
		}
	}

	/* src/CController.hpp:385-391 */
	void computeNextStep()
	{
		connectFaces();
		if (resolvedSync() == SYNC_P2P && !_comm_container.empty()) {
			cLbmPtr->simulationStepWithHalo();      /* shell | exchange || interior, one call */
			return;
		}
		cLbmPtr->simulationStep();
		if (cLbmPtr->simulation_step_counter & 1) syncBeta();
		else syncAlpha();
	}

	/* src/CController.hpp:396-522 */
	int run()
	{
		CConfiguration<T> *cfg = ConfigSingleton::Instance();
		CVector<3, int> domain_size = _domain.getSize();
		int loops = cfg->loops;
		if (loops < 0) loops = 100;
		vector_checksum = 0;
		double floats_per_cell = 19.0 * 2.0 + 1.0;
		if (cfg->do_visualization || cfg->debug_mode) floats_per_cell += 3;

		/* optional per-step VTK dump (src/CController.hpp:419-436): output/vtk/OUTPUT.<uid>.<step>.vtk */
		CLbmVisualizationVTK<T> *cLbmVisualization = NULL;
		if (cfg->do_visualization) {
			mkdir("output", 0755);
			mkdir(VTK_OUTPUT_DIR, 0755);
			cLbmVisualization = new CLbmVisualizationVTK<T>(_UID, std::string("./") + VTK_OUTPUT_DIR + "/OUTPUT");
			cLbmVisualization->setup(cLbmPtr);
		}

		connectFaces();
		cLbmPtr->wait();
		if (_world) _world->barrier();
		const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
		for (int i = 0; i < loops; i++) {
			computeNextStep();
			if (cLbmVisualization) cLbmVisualization->render(i);
			if (cLbmPtr->error()) { std::cerr << cLbmPtr->error.getString(); if (_world) _world->fail(); return EXIT_FAILURE; }
		}
		cLbmPtr->wait();
		delete cLbmVisualization;
		seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

		const double gtime = (_world && _UID >= 0) ? _world->reduceMax(seconds) : seconds;
		if (_UID <= 0 && getenv("LBM_B200_BENCHMARK")) {
			const double gfps = (double)loops / gtime;
			const double gmlups = gfps * (double)cfg->domain_size.elements() * 0.000001;
			std::ostringstream name;
			name << "./" << BENCHMARK_OUTPUT_DIR << "/benchmark_" << cfg->subdomain_num.elements() << ".ini";
			std::ofstream f(name.str().c_str(), std::ios::out | std::ios::app);
			if (f.is_open()) {
				f << "CUBE_X : " << cfg->domain_size[0] << std::endl << "CUBE_Y : " << cfg->domain_size[1] << std::endl
				  << "CUBE_Z : " << cfg->domain_size[2] << std::endl << "SECONDS : " << gtime << std::endl
				  << "FPS : " << gfps << std::endl << "MLUPS : " << gmlups << std::endl
				  << "BANDWIDTH : " << gmlups * floats_per_cell * (double)sizeof(T) << std::endl << std::endl;
			} else std::cout << "Unable to open file";
		}

		const double fps = (double)loops / seconds;
		mlups = fps * (double)cLbmPtr->domain_cells.elements() * 0.000001;
		if (cfg->debug_mode) vector_checksum = cLbmPtr->getVelocityChecksum();
		{
			std::unique_lock<std::mutex> guard;
			if (_world) guard = std::unique_lock<std::mutex>(_world->print_mutex);
			std::cout << std::endl;
			std::cout << "Cube: " << domain_size << std::endl;
			std::cout << "Seconds: " << seconds << std::endl;
			std::cout << "FPS: " << fps << std::endl;
			std::cout << "MLUPS: " << mlups << std::endl;
			std::cout << "Bandwidth: " << (mlups * floats_per_cell * (double)sizeof(T)) << " MB/s (RW, bidirectional)" << std::endl;
			if (cfg->debug_mode) {
				std::streamsize ss = std::cout.precision();
				std::cout.precision(8);
				std::cout.setf(std::ios::fixed, std::ios::floatfield);
				std::cout << "Checksum: " << (vector_checksum * 1000.0f) << std::endl;
				std::cout.precision(ss);
				std::cout << std::resetiosflags(std::ios::fixed);
			}
			std::cout << "done." << std::endl;
		}

		/* PROFILE block of src/CController.hpp:503-519: output/profile/profile_<np>_<uid>.ini.  The
		 * reference's compile-time switch is the run-time variable LBM_B200_PROFILE (read by lbmCreate). */
		uint64_t prof_events = 0;
		if (lbmProfileEventCount(cLbmPtr->handle(), &prof_events, NULL) == LBM_OK && prof_events > 0) {
			mkdir("output", 0755);
			mkdir(PROFILE_OUTPUT_DIR, 0755);
			std::ostringstream profile_file_name;
			profile_file_name << "./" << PROFILE_OUTPUT_DIR << "/" << "profile_" << cfg->subdomain_num.elements()
			                  << "_" << _UID << ".ini";
			if (profiler.collect(cLbmPtr->handle()) != LBM_OK)
				std::cerr << "profile: " << lbmGetLastErrorString(cLbmPtr->handle()) << std::endl;
			profiler.saveProfile(profile_file_name.str(), cfg->subdomain_num.elements(), _UID);
		}
		return EXIT_SUCCESS;
	}

	void addCommunication(CComm<T> *comm) { _comm_container.push_back(comm); }

	/* lid: flags := VELOCITY_INJECTION on y = Sy-2, x in [1,Sx-2], z in [1,Sz-2] (src/CController.hpp:531-546) */
	void setGeometry()
	{
		CVector<3, int> S = _domain.getSize();
		CVector<3, int> origin(1, S[1] - 2, 1);
		CVector<3, int> size(S[0] - 2, 1, S[2] - 2);
		std::vector<int> src((size_t)size.elements(), FLAG_VELOCITY_INJECTION);
		cLbmPtr->setFlags(src.data(), origin, size);
	}

	CLbmSolver<T> *getSolver() const { return cLbmPtr; }
	void setSolver(CLbmSolver<T> *s) { cLbmPtr = s; }
	CDomain<T> getDomain() const { return _domain; }
	int getUid() const { return _UID; }
	const std::vector<CComm<T> *> &getComms() const { return _comm_container; }
	LbmSyncMode syncMode() const { return resolvedSync(); }
};

#endif
