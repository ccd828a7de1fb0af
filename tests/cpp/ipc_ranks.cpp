/*
 * ipc_ranks.cpp -- one PROCESS per sub-domain, driven from C++ through the C ABI only (no Python,
 * no torch, no MPI): what a maintainer of the reference would write around its MPI ranks
 * (src/main.cpp:415-423, src/CController.hpp:299-311).
 *
 *   fork()            stands in for mpirun: N children, each owning one sub-domain on GPU (rank % ndev)
 *   socketpair()      stands in for the MPI control plane: it only carries the 64-byte CUDA-IPC handles
 *                     of the halo receive blocks, the start barrier and the validation blocks
 *   data plane        lbmCommStep: one-sided NVLink / peer-memory stores, device-side flags
 *
 * Acceptance is the reference's validate criterion (src/main.cpp:309-408): the interior velocity block
 * of every rank equals, bit for bit, the matching block of a single-domain run of size D - 2 (n - 1),
 * which the parent computes after the children have finished.
 *
 * usage: ipc_ranks X Y Z NX NY NZ steps [zyx] [smagorinsky] [double]   exit code 0 = "0 failed cells"
 */
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <sys/socket.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include "../../turbulent_lbm_multigpu_b200/host/CManager.hpp"

struct FaceMsg { int rank, dst, axis, sign; unsigned char handle[64]; };

static bool write_all(int fd, const void *p, size_t n)
{
	const char *c = (const char *)p;
	while (n) { ssize_t k = write(fd, c, n); if (k <= 0) return false; c += k; n -= (size_t)k; }
	return true;
}
static bool read_all(int fd, void *p, size_t n)
{
	char *c = (char *)p;
	while (n) { ssize_t k = read(fd, c, n); if (k <= 0) return false; c += k; n -= (size_t)k; }
	return true;
}

template <typename T>
static CLbmSolver<T> *make_solver(int uid, int device, int BC[3][2], CDomain<T> &dom, T cs)
{
	static CCL::CCommandQueue queue;
	static CCL::CContext context;
	CCL::CDevice *dev = new CCL::CDevice(device);
	CVector<3, T> g(0, (T)-9.81, 0);
	CVector<4, T> lid((T)100, 0, 0, (T)1);
	std::list<int> a, b;
	return new CLbmSolver<T>(uid, queue, context, *dev, BC, dom, g, (T)0.001308, 128, true, false, (T)-1, lid, a, b, cs);
}

template <typename T>
static void set_lid(CLbmSolver<T> *s, CVector<3, int> S)
{
	/* CController::setGeometry, src/CController.hpp:531-546 */
	CVector<3, int> origin(1, S[1] - 2, 1), size(S[0] - 2, 1, S[2] - 2);
	std::vector<int> src((size_t)size.elements(), FLAG_VELOCITY_INJECTION);
	s->setFlags(src.data(), origin, size);
}

template <typename T>
static int child(int rank, int fd, CVector<3, int> D, CVector<3, int> nums, int steps, bool zyx, T cs)
{
	int ndev = CCL::CContext::deviceCount();
	if (ndev <= 0) { fprintf(stderr, "rank %d: no CUDA device\n", rank); return 3; }
	CVector<3, int> origin0(0, 0, 0);
	CVector<3, T> L((T)0.1, (T)0.1, (T)0.1);
	CManager<T> manager(CDomain<T>(-1, D, origin0, L), nums);
	int BC[3][2];
	std::vector<CComm<T> > comms;
	CVector<3, int> origin;
	manager.layout(rank, BC, comms, origin);
	CVector<3, int> S = manager.getSubdomainSize();
	CVector<3, T> subL;
	for (int a = 0; a < 3; a++) subL[a] = L[a] / (T)nums[a];
	CDomain<T> sub(rank, S, origin, subL);
	CLbmSolver<T> *s = make_solver<T>(rank, rank % ndev, BC, sub, cs);
	if (s->error()) { fprintf(stderr, "rank %d: %s\n", rank, s->error.getString().c_str()); return 4; }
	if (manager.rankCoords(rank)[1] == nums[1] - 1) set_lid(s, S);
	lbm_t h = s->handle();
	if (zyx && lbmCommSetAxisOrder(h, LBM_AXIS_ORDER_ZYX) != LBM_OK) return 5;

	/* register the faces, publish their IPC handles, map the neighbours' */
	std::vector<int> fids(comms.size());
	int nfaces = (int)comms.size();
	if (!write_all(fd, &nfaces, sizeof(nfaces))) return 6;
	for (size_t i = 0; i < comms.size(); i++) {
		CComm<T> &c = comms[i];
		CVector<3, int> so = c.getSendOrigin(), ro = c.getRecvOrigin(), sz = c.getSendSize(), dir = c.getCommDirection();
		if (lbmCommAddFace(h, c.getDstId(), so.data, ro.data, sz.data, dir.data, LBM_HALO_SLOTS_MINIMAL, &fids[i]) != LBM_OK) {
			fprintf(stderr, "rank %d: %s\n", rank, lbmGetLastErrorString(h)); return 7;
		}
		FaceMsg m;
		m.rank = rank; m.dst = c.getDstId();
		m.axis = dir[0] ? 0 : dir[1] ? 1 : 2; m.sign = dir[m.axis];
		if (lbmCommGetIpcHandle(h, fids[i], m.handle) != LBM_OK) { fprintf(stderr, "rank %d: %s\n", rank, lbmGetLastErrorString(h)); return 8; }
		if (!write_all(fd, &m, sizeof(m))) return 9;
	}
	int total = 0;
	if (!read_all(fd, &total, sizeof(total))) return 10;
	std::vector<FaceMsg> table((size_t)total);
	if (total && !read_all(fd, table.data(), sizeof(FaceMsg) * (size_t)total)) return 11;
	for (size_t i = 0; i < comms.size(); i++) {
		CComm<T> &c = comms[i];
		CVector<3, int> dir = c.getCommDirection();
		const int axis = dir[0] ? 0 : dir[1] ? 1 : 2;
		const FaceMsg *peer = NULL;                /* the neighbour's face that looks back at me */
		for (int k = 0; k < total; k++)
			if (table[k].rank == c.getDstId() && table[k].dst == rank && table[k].axis == axis && table[k].sign == -dir[axis]) peer = &table[k];
		if (!peer) { fprintf(stderr, "rank %d: no peer face\n", rank); return 12; }
		if (lbmCommConnectIpc(h, fids[i], peer->handle) != LBM_OK) { fprintf(stderr, "rank %d: %s\n", rank, lbmGetLastErrorString(h)); return 13; }
	}
	char go = 0;
	if (!write_all(fd, "c", 1) || !read_all(fd, &go, 1)) return 14;       /* everybody connected */

	for (int i = 0; i < steps; i++)
		if (lbmCommStep(h) != LBM_OK) { fprintf(stderr, "rank %d: %s\n", rank, lbmGetLastErrorString(h)); return 15; }
	if (lbmWait(h) != LBM_OK) { fprintf(stderr, "rank %d: %s\n", rank, lbmGetLastErrorString(h)); return 16; }

	/* interior block: origin (1,1,1), size S-2 (src/main.cpp:332-337) */
	CVector<3, int> inner(S[0] - 2, S[1] - 2, S[2] - 2), one(1, 1, 1);
	std::vector<T> vel((size_t)inner.elements() * 3);
	s->storeVelocity(vel.data(), one, inner);
	if (s->error()) { fprintf(stderr, "rank %d: %s\n", rank, s->error.getString().c_str()); return 17; }
	if (!write_all(fd, vel.data(), vel.size() * sizeof(T))) return 18;
	if (!read_all(fd, &go, 1)) return 19;        /* nobody unmaps halo blocks while a neighbour may still write */
	delete s;
	return 0;
}

template <typename T>
static int run(CVector<3, int> D, CVector<3, int> nums, int steps, bool zyx, T cs)
{
	const int n = nums.elements();
	std::vector<int> fds((size_t)n);
	std::vector<pid_t> pids((size_t)n);
	for (int r = 0; r < n; r++) {
		int sv[2];
		if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) { perror("socketpair"); return 2; }
		pid_t pid = fork();                       /* before this process touches CUDA */
		if (pid < 0) { perror("fork"); return 2; }
		if (pid == 0) {
			close(sv[0]);
			for (int k = 0; k < r; k++) close(fds[(size_t)k]);
			_exit(child<T>(r, sv[1], D, nums, steps, zyx, cs));
		}
		close(sv[1]);
		fds[(size_t)r] = sv[0];
		pids[(size_t)r] = pid;
	}
	bool ok = true;
	/* rendezvous: gather every face handle, broadcast the table */
	std::vector<FaceMsg> table;
	for (int r = 0; r < n && ok; r++) {
		int nf = 0;
		ok = read_all(fds[(size_t)r], &nf, sizeof(nf));
		for (int k = 0; k < nf && ok; k++) { FaceMsg m; ok = read_all(fds[(size_t)r], &m, sizeof(m)); table.push_back(m); }
	}
	int total = (int)table.size();
	for (int r = 0; r < n && ok; r++)
		ok = write_all(fds[(size_t)r], &total, sizeof(total)) && (!total || write_all(fds[(size_t)r], table.data(), sizeof(FaceMsg) * table.size()));
	char c = 0;
	for (int r = 0; r < n && ok; r++) ok = read_all(fds[(size_t)r], &c, 1);
	for (int r = 0; r < n && ok; r++) ok = write_all(fds[(size_t)r], "g", 1);

	CVector<3, int> S(D[0] / nums[0], D[1] / nums[1], D[2] / nums[2]);
	CVector<3, int> inner(S[0] - 2, S[1] - 2, S[2] - 2);
	std::vector<std::vector<T> > blocks((size_t)n);
	for (int r = 0; r < n && ok; r++) {
		blocks[(size_t)r].resize((size_t)inner.elements() * 3);
		ok = read_all(fds[(size_t)r], blocks[(size_t)r].data(), blocks[(size_t)r].size() * sizeof(T));
	}
	for (int r = 0; r < n; r++) write_all(fds[(size_t)r], "d", 1);
	int failed_ranks = 0;
	for (int r = 0; r < n; r++) {
		int st = 0;
		waitpid(pids[(size_t)r], &st, 0);
		if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { fprintf(stderr, "rank %d exited with %d\n", r, WIFEXITED(st) ? WEXITSTATUS(st) : -1); failed_ranks++; }
	}
	if (!ok || failed_ranks) { printf("validation: not run (%d ranks failed)\n", failed_ranks); return 1; }

	/* the single domain of the validate mode (src/main.cpp:358-387), computed here, after the ranks */
	CVector<3, int> V(D[0] - 2 * (nums[0] - 1), D[1] - 2 * (nums[1] - 1), D[2] - 2 * (nums[2] - 1)), origin0(0, 0, 0);
	CVector<3, T> L((T)0.1, (T)0.1, (T)0.1), subL;
	for (int a = 0; a < 3; a++) subL[a] = L[a] / (T)nums[a];
	/* same cell length (src/main.cpp:358-361), hence the same tau / u_lid (src/CLbmSkeleton.hpp:157) */
	CVector<3, T> VL;
	for (int a = 0; a < 3; a++) VL[a] = (T)V[a] * (L[a] / (T)D[a]);
	int BC[3][2] = { { FLAG_OBSTACLE, FLAG_OBSTACLE }, { FLAG_OBSTACLE, FLAG_OBSTACLE }, { FLAG_OBSTACLE, FLAG_OBSTACLE } };
	CDomain<T> vdom(0, V, origin0, VL);
	CLbmSolver<T> *single = make_solver<T>(0, 0, BC, vdom, cs);
	if (single->error()) { fprintf(stderr, "single: %s\n", single->error.getString().c_str()); return 1; }
	/* bit-identical parametrisation is part of the criterion: take the sub-domain's numbers */
	{
		CDomain<T> sub(0, S, origin0, subL);
		CVector<4, T> lid((T)100, 0, 0, (T)1);
		CLbmSkeleton<T> p(sub, lid);
		CVector<3, T> g(0, (T)-9.81, 0);
		p.init(g, (T)0.001308, (T)1.0);
		if ((double)p.inv_tau != (double)single->inv_tau || (double)p.drivenCavityVelocity[0] != (double)((CLbmSkeleton<T> *)single)->drivenCavityVelocity[0]) {
			fprintf(stderr, "validation domain parametrised differently (inv_tau %.9g vs %.9g)\n", (double)single->inv_tau, (double)p.inv_tau);
			return 1;
		}
	}
	set_lid(single, V);
	for (int i = 0; i < steps; i++) single->simulationStep();
	single->wait();
	long long failed = 0, checked = 0;
	double maxv = 0;
	for (int r = 0; r < n; r++) {
		int id = r;
		const int cx = id % nums[0]; id /= nums[0];
		const int cy = id % nums[1]; id /= nums[1];
		const int cz = id;
		CVector<3, int> o(1 + cx * inner[0], 1 + cy * inner[1], 1 + cz * inner[2]);
		std::vector<T> ref((size_t)inner.elements() * 3);
		single->storeVelocity(ref.data(), o, inner);
		for (size_t k = 0; k < ref.size(); k++) {
			checked++;
			if (memcmp(&ref[k], &blocks[(size_t)r][k], sizeof(T)) != 0) failed++;
			if ((double)ref[k] > maxv) maxv = (double)ref[k];
		}
	}
	delete single;
	printf("validation: %lld failed cells of %lld (%d processes, max velocity %g)\n", failed, checked, n, maxv);
	return failed == 0 && maxv > 0 ? 0 : 1;
}

int main(int argc, char **argv)
{
	if (argc < 8) { fprintf(stderr, "usage: %s X Y Z NX NY NZ steps [zyx] [smagorinsky] [double]\n", argv[0]); return 2; }
	CVector<3, int> D(atoi(argv[1]), atoi(argv[2]), atoi(argv[3])), nums(atoi(argv[4]), atoi(argv[5]), atoi(argv[6]));
	const int steps = atoi(argv[7]);
	signal(SIGPIPE, SIG_IGN);                 /* a rank that died is reported, not a reason to die too */
	bool zyx = false, dbl = false;
	double cs = 0;
	for (int i = 8; i < argc; i++) {
		if (!strcmp(argv[i], "zyx")) zyx = true;
		else if (!strcmp(argv[i], "double")) dbl = true;
		else cs = atof(argv[i]);
	}
	try {
		return dbl ? run<double>(D, nums, steps, zyx, cs) : run<float>(D, nums, steps, zyx, (float)cs);
	} catch (const char *msg) {
		fprintf(stderr, "%s\n", msg);
		return 2;
	}
}
