/*
 * CManager.hpp -- decomposition of the global domain into sub-domains and their wiring.
 * Public surface and semantics of the reference's src/CManager.hpp: sub-domain size =
 * domain_size / subdomain_num INCLUDING the ghost layers (:57-63), rank id =
 * nx + ny*NX + nz*NX*NY (:87-101), faces toward a neighbour are GHOST_LAYER, outer faces
 * OBSTACLE (:103-117), up to six CComm in the order x-,x+,y-,y+,z-,z+ (:122-199), the lid
 * geometry on the ranks at the top of y (:200-202).
 */
#ifndef LBM_B200_HOST_CMANAGER_HPP
#define LBM_B200_HOST_CMANAGER_HPP

#include "CController.hpp"

template <typename T>
class CManager {
	CDomain<T> _domain;
	CVector<3, int> _subdomain_size, _subdomain_nums;
	CVector<3, T> _subdomain_length;
	CController<T> *_lbm_controller;
	CRankWorld *_world;
	LbmSyncMode _sync;
	int _beta_order;

public:
	CManager(CDomain<T> domain, CVector<3, int> subdomainNums, CRankWorld *world = NULL, LbmSyncMode sync = SYNC_AUTO,
			int beta_order = LBM_BETA_ORDER_SHIPPED)
		: _domain(domain), _lbm_controller(NULL), _world(world), _sync(sync), _beta_order(beta_order)
	{
		this->setSubdomainNums(subdomainNums);
	}

	~CManager() { delete _lbm_controller; }

	CDomain<T> getDomain() const { return _domain; }
	void setDomain(CDomain<T> grid) { _domain = grid; }
	CVector<3, int> getSubdomainNums() const { return _subdomain_nums; }
	CVector<3, int> getSubdomainSize() const { return _subdomain_size; }

	void setSubdomainNums(CVector<3, int> subdomainNums)
	{
		CVector<3, int> D = _domain.getSize();
		for (int a = 0; a < 3; a++)
			if (subdomainNums[a] <= 0 || D[a] % subdomainNums[a] != 0)
				throw "Number of subdomains does not match with the grid size!";
		CVector<3, T> L = _domain.getLength();
		for (int a = 0; a < 3; a++) {
			_subdomain_size[a] = D[a] / subdomainNums[a];
			_subdomain_length[a] = L[a] / (T)subdomainNums[a];
		}
		_subdomain_nums = subdomainNums;
	}

	/* position of a rank in the grid of sub-domains, x fastest */
	CVector<3, int> rankCoords(int id) const
	{
		CVector<3, int> c;
		c[0] = id % _subdomain_nums[0]; id /= _subdomain_nums[0];
		c[1] = id % _subdomain_nums[1]; id /= _subdomain_nums[1];
		c[2] = id;
		return c;
	}

	/* BC table and CComm list of a rank -- pure host logic, usable without a GPU */
	void layout(int my_rank, int BC[3][2], std::vector<CComm<T> > &comms, CVector<3, int> &origin) const
	{
		const int id = my_rank < 0 ? 0 : my_rank;
		const CVector<3, int> c = rankCoords(id), S = _subdomain_size;
		const int stride[3] = { 1, _subdomain_nums[0], _subdomain_nums[0] * _subdomain_nums[1] };
		comms.clear();
		for (int a = 0; a < 3; a++) {
			BC[a][0] = c[a] == 0 ? FLAG_OBSTACLE : FLAG_GHOST_LAYER;
			BC[a][1] = c[a] == _subdomain_nums[a] - 1 ? FLAG_OBSTACLE : FLAG_GHOST_LAYER;
			origin[a] = c[a] * S[a];
		}
		for (int a = 0; a < 3; a++) {
			CVector<3, int> face = S;
			face[a] = 1;
			for (int side = 0; side < 2; side++) {
				if (BC[a][side] != FLAG_GHOST_LAYER) continue;
				CVector<3, int> send_origin(0, 0, 0), recv_origin(0, 0, 0), dir(0, 0, 0);
				send_origin[a] = side == 0 ? 1 : S[a] - 2;      /* my outermost real layer */
				recv_origin[a] = side == 0 ? 0 : S[a] - 1;      /* my ghost layer */
				dir[a] = side == 0 ? 1 : -1;                    /* points into this sub-domain */
				comms.push_back(CComm<T>(id + (side == 0 ? -stride[a] : stride[a]), face, face, send_origin, recv_origin, dir));
			}
		}
	}

	void initSimulation(int my_rank)
	{
		int BC[3][2];
		std::vector<CComm<T> > comms;
		CVector<3, int> origin;
		layout(my_rank, BC, comms, origin);
		const int id = my_rank < 0 ? 0 : my_rank;
		CDomain<T> subdomain(id, _subdomain_size, origin, _subdomain_length);
		_lbm_controller = new CController<T>(id, subdomain, BC, _world, _sync, _beta_order);
		for (size_t i = 0; i < comms.size(); i++) _lbm_controller->addCommunication(new CComm<T>(comms[i]));
		if (rankCoords(id)[1] == _subdomain_nums[1] - 1) _lbm_controller->setGeometry();
	}

	void startSimulation()
	{
		if (!_lbm_controller) throw "CManager: Initialize the simulation before starting it!";
		_lbm_controller->run();
	}

	CController<T> *getController() const { return _lbm_controller; }
	void setController(CController<T> *c) { _lbm_controller = c; }
};

#endif
