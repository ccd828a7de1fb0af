/*
 * main.cpp -- the `lbm_b200` driver: the reference's command line (src/main.cpp:96-435) in front
 * of the B200 library.  Same flags (-x -y -z -X -Y -Z -S -l -k -d -r -n -m -p -G -t -v -c ...),
 * same two run modes (normal: :409-430, validate: :309-408), same stdout block per rank.
 * One process drives all sub-domains: a rank is a host thread with its own GPU (CRankWorld),
 * so `mpirun -np N ./lbm_opencl ...` becomes `./lbm_b200 ...`.
 *
 * Long options are additions: --double (T = double), --smagorinsky C, --sync p2p|copy|host,
 * --beta-order shipped|linear, --dump-layout, --dump-params, --dump-velocity FILE (rank blocks
 * of the interior velocity, for tests), --help.
 */
#include <getopt.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <exception>
#include <vector>

#include "CManager.hpp"

#define VALIDATION_RANK 0

struct Options {
	bool use_double = false, dump_layout = false, dump_params = false, debug = false;
	std::string dump_velocity;
	LbmSyncMode sync = SYNC_AUTO;
	int beta_order = LBM_BETA_ORDER_SHIPPED;
};

static void usage(const char *argv0)
{
	std::cout << "usage: " << argv0 << std::endl
		<< "		[-x resolution_x, default: 32] [-y resolution_y] [-z resolution_z] [-S resolution of all axes]" << std::endl
		<< "		[-X subdomains in x, default: 1] [-Y subdomains in y] [-Z subdomains in z]" << std::endl
		<< "		[-n domain length x, default: 0.1] [-m domain length y] [-p domain length z]" << std::endl
		<< "		[-G gravitation in down direction, default: -9.81]" << std::endl
		<< "		[-r viscosity, default: 0.001308]" << std::endl
		<< "		[-l loops, default: 100]" << std::endl
		<< "		[-v]	(debug mode on, be verbose: stores velocity/density, prints the checksum)" << std::endl
		<< "		[-d device_num]	(-1: list available devices, or first device of the run)" << std::endl
		<< "		[-k work group size of the reference kernels, default: 128]	(selects the shipped beta kernel's x-shift only)" << std::endl
		<< "		[-t timestep]	(default: -1 for automatic detection)" << std::endl
		<< "		[-g]	(write output/vtk/OUTPUT.<rank>.<step>.vtk every step, default: disabled)" << std::endl
		<< "		[-R list] [-T list] [-s]	(accepted for compatibility, ignored)" << std::endl
		<< "		[-c conf.xml]	read the configuration file (replaces all of the above)" << std::endl
		<< "		[--double] [--smagorinsky C_s] [--sync p2p|copy|host] [--beta-order shipped|linear]" << std::endl
		<< "		[--validate] [--dump-layout] [--dump-params] [--dump-velocity FILE]" << std::endl;
}

template <typename T>
static void print_hex(const char *name, T v)
{
	if (sizeof(T) == 4) { float f = (float)v; unsigned u; std::memcpy(&u, &f, 4); std::printf("%s %.9g 0x%08x\n", name, (double)f, u); }
	else { double d = (double)v; unsigned long long u; std::memcpy(&u, &d, 8); std::printf("%s %.17g 0x%016llx\n", name, d, u); }
}

/* one rank of the decomposed run (what a reference MPI process does, src/main.cpp:409-430) */
template <typename T>
static void rank_main(int rank, CDomain<T> domain, CVector<3, int> nums, CRankWorld *world, const Options *opt,
		std::vector<T> *validation_data, CVector<3, int> *validation_size, int *status)
{
	try {
		CManager<T> manager(domain, nums, world, opt->sync, opt->beta_order);
		manager.initSimulation(rank);
		manager.startSimulation();
		CController<T> *controller = manager.getController();
		if (validation_data && (rank == VALIDATION_RANK || !opt->dump_velocity.empty())) {
			/* interior block of this rank: origin (1,1,1), size S-2 (src/main.cpp:332-337) */
			CVector<3, int> S = controller->getDomain().getSize();
			CVector<3, int> inner(S[0] - 2, S[1] - 2, S[2] - 2), origin(1, 1, 1);
			validation_data->resize((size_t)inner.elements() * 3);
			controller->getSolver()->storeVelocity(validation_data->data(), origin, inner);
			*validation_size = inner;
		}
		*status = controller->getSolver()->error() ? 1 : 0;
		world->barrier();           /* nobody unmaps halo blocks while a neighbour may still write */
	} catch (const char *msg) {
		std::cerr << "rank " << rank << ": " << msg << std::endl;
		world->fail();
		*status = 1;
	} catch (const std::exception &e) {
		/* bad_alloc from the validation / recv buffers, system_error ...: release the peers from their
		 * barriers and device-side flag waits instead of std::terminate */
		std::cerr << "rank " << rank << ": " << e.what() << std::endl;
		world->fail();
		*status = 1;
	} catch (...) {
		std::cerr << "rank " << rank << ": unknown exception" << std::endl;
		world->fail();
		*status = 1;
	}
}

template <typename T>
static int run(CConfiguration<T> *cfg, const Options &opt)
{
	typedef Singleton<CConfiguration<T> > ConfigSingleton;
	cfg->debug_mode = opt.debug;
	if (opt.debug) cfg->printMe();
	if (cfg->device_nr == -1) {
		std::cout << "CUDA devices: " << CCL::CContext::deviceCount() << std::endl;
		return 0;
	}
	CVector<3, int> origin(0, 0, 0);
	CDomain<T> domain(-1, cfg->domain_size, origin, cfg->domain_length);
	const int nranks = cfg->subdomain_num.elements();

	if (opt.dump_layout || opt.dump_params) {
		CManager<T> manager(domain, cfg->subdomain_num);
		if (opt.dump_layout) {
			CVector<3, int> S = manager.getSubdomainSize();
			std::printf("subdomain_size %d %d %d\n", S[0], S[1], S[2]);
			for (int r = 0; r < nranks; r++) {
				int BC[3][2];
				std::vector<CComm<T> > comms;
				CVector<3, int> o;
				manager.layout(r, BC, comms, o);
				std::printf("rank %d origin %d %d %d bc %d %d %d %d %d %d ncomm %d\n", r, o[0], o[1], o[2], BC[0][0], BC[0][1],
						BC[1][0], BC[1][1], BC[2][0], BC[2][1], (int)comms.size());
				for (size_t i = 0; i < comms.size(); i++) {
					const CComm<T> &c = comms[i];
					CVector<3, int> ss = c.getSendSize(), rs = c.getRecvSize(), so = c.getSendOrigin(), ro = c.getRecvOrigin(), d = c.getCommDirection();
					std::printf("  comm dst %d send_size %d %d %d recv_size %d %d %d send_origin %d %d %d recv_origin %d %d %d dir %d %d %d\n",
							c.getDstId(), ss[0], ss[1], ss[2], rs[0], rs[1], rs[2], so[0], so[1], so[2], ro[0], ro[1], ro[2], d[0], d[1], d[2]);
				}
			}
		}
		if (opt.dump_params) {
			/* the parametrisation of one sub-domain, bit patterns included (no GPU needed) */
			CVector<3, T> L = domain.getLength();
			CVector<3, T> subL(L[0] / (T)cfg->subdomain_num[0], L[1] / (T)cfg->subdomain_num[1], L[2] / (T)cfg->subdomain_num[2]);
			CDomain<T> sub(0, manager.getSubdomainSize(), origin, subL);
			CLbmSkeleton<T> sk(sub, cfg->drivenCavityVelocity);
			sk.init(cfg->gravitation, cfg->viscosity, (T)1.0);
			print_hex("d_cell_length", sk.d_cell_length);
			print_hex("d_timestep", sk.d_timestep);
			print_hex("tau", sk.tau);
			print_hex("inv_tau", sk.inv_tau);
			print_hex("inv_trt_tau", sk.inv_trt_tau);
			print_hex("gravitation_x", sk.gravitation[0]);
			print_hex("gravitation_y", sk.gravitation[1]);
			print_hex("gravitation_z", sk.gravitation[2]);
			print_hex("u_lid", sk.drivenCavityVelocity[0]);
			print_hex("d_reynolds", sk.d_reynolds);
			std::printf("tau_ok %d\n", sk.error() ? 0 : 1);
		}
		return 0;
	}

	const bool validate = cfg->do_validate;
	CRankWorld world(nranks);
	std::vector<std::thread> threads;
	std::vector<std::vector<T> > rank_data(nranks);
	std::vector<CVector<3, int> > rank_size(nranks);
	std::vector<int> status(nranks, 0);
	const bool collect = validate || !opt.dump_velocity.empty();
	for (int r = 0; r < nranks; r++)
		threads.push_back(std::thread(rank_main<T>, r, domain, cfg->subdomain_num, &world, &opt,
				collect ? &rank_data[r] : (std::vector<T> *)NULL, &rank_size[r], &status[r]));
	for (int r = 0; r < nranks; r++) threads[r].join();
	for (int r = 0; r < nranks; r++) if (status[r]) return EXIT_FAILURE;

	if (!opt.dump_velocity.empty()) {
		FILE *f = std::fopen(opt.dump_velocity.c_str(), "wb");
		if (!f) { std::cerr << "cannot write " << opt.dump_velocity << std::endl; return EXIT_FAILURE; }
		for (int r = 0; r < nranks; r++) {
			int hdr[4] = { r, rank_size[r][0], rank_size[r][1], rank_size[r][2] };
			std::fwrite(hdr, sizeof(int), 4, f);
			std::fwrite(rank_data[r].data(), sizeof(T), rank_data[r].size(), f);
		}
		std::fclose(f);
	}

	if (validate) {
		/* src/main.cpp:341-406: the single domain every decomposed run must equal */
		const int my_rank = nranks;
		CVector<3, int> local = rank_size[VALIDATION_RANK];
		CVector<3, int> vsize;
		CVector<3, T> vlen;
		for (int a = 0; a < 3; a++) {
			vsize[a] = cfg->domain_size[a] - 2 * (cfg->subdomain_num[a] - 1);
			const T cell = cfg->domain_length[a] / (T)cfg->domain_size[a];
			vlen[a] = vsize[a] * cell;
		}
		CDomain<T> vdomain(-2, vsize, origin, vlen);
		CManager<T> vmanager(vdomain, CVector<3, int>(1, 1, 1), NULL, SYNC_AUTO, opt.beta_order);
		std::cout << my_rank << "--> Compute the results for one domain." << std::endl;
		vmanager.initSimulation(-1);
		vmanager.startSimulation();
		int id = VALIDATION_RANK;
		const int nx = id % cfg->subdomain_num[0]; id /= cfg->subdomain_num[0];
		const int ny = id % cfg->subdomain_num[1]; id /= cfg->subdomain_num[1];
		const int nz = id;
		CVector<3, int> sub_origin(1 + nx * local[0], 1 + ny * local[1], 1 + nz * local[2]);
		std::vector<T> global((size_t)local.elements() * 3);
		vmanager.getController()->getSolver()->storeVelocity(global.data(), sub_origin, local);
		std::cout << "PROC. RANK: " << my_rank << " VALIDATION SIZE: " << vsize << std::endl;
		const double tolerance = 1.0e-15;
		int error_counter = 0;
		const std::vector<T> &mine = rank_data[VALIDATION_RANK];
		for (size_t i = 0; i < global.size(); i++)
			if (std::fabs(global[i] - mine[i]) > tolerance) error_counter++;
		std::cout << "--> PROC. RANK: " << my_rank << " TOLERANCE: " << tolerance
			<< " NUMBER OF FAILED CELLS/TOTAL NUMBER OF CELLS: " << error_counter << "/" << local.elements() << std::endl;
		(void)ConfigSingleton::Instance();
		return error_counter == 0 ? 0 : 2;
	}
	return 0;
}

int main(int argc, char **argv)
{
	/* rank threads that share a GPU keep two streams each, and a flag wait of the one-sided exchange spins
	 * until the neighbour's kernels have run: with the default of 8 hardware queues a neighbour's stream can end
	 * up queued BEHIND such a wait (false dependency) and the run stalls until the wait times out.  Must be set
	 * before the first CUDA call; an explicit setting of the user wins. */
	setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
	/* parsed into double, converted once into the simulation type */
	int domain_size[3] = { 32, 32, 32 }, subdomain_nums[3] = { 1, 1, 1 };
	double domain_length[3] = { 0.1, 0.1, 0.1 }, gravitation_y = -9.81, viscosity = 0.001308, timestep = -1.0, smagorinsky = 0.0;
	int loops = -1, device_nr = 0;
	size_t kernel_count = 128;
	bool do_validate = false, do_visualisation = false, use_config_file = false;
	std::string conf_file;
	Options opt;

	static const struct option long_opts[] = {
		{ "double", no_argument, 0, 1000 }, { "smagorinsky", required_argument, 0, 1001 },
		{ "sync", required_argument, 0, 1002 }, { "beta-order", required_argument, 0, 1003 },
		{ "dump-layout", no_argument, 0, 1004 }, { "dump-params", no_argument, 0, 1005 },
		{ "dump-velocity", required_argument, 0, 1006 }, { "validate", no_argument, 0, 1007 },
		{ "help", no_argument, 0, 1008 }, { 0, 0, 0, 0 } };
	int c;
	while ((c = getopt_long(argc, argv, "x:y:z:d:vr:k:gG:t:sl:R:T:X:Y:Z:S:u:c:n:m:p:", long_opts, NULL)) > 0) {
		switch (c) {
		case 'x': domain_size[0] = atoi(optarg); break;
		case 'y': domain_size[1] = atoi(optarg); break;
		case 'z': domain_size[2] = atoi(optarg); break;
		case 'S': domain_size[0] = domain_size[1] = domain_size[2] = atoi(optarg); break;
		case 'X': subdomain_nums[0] = atoi(optarg); break;
		case 'Y': subdomain_nums[1] = atoi(optarg); break;
		case 'Z': subdomain_nums[2] = atoi(optarg); break;
		case 'n': domain_length[0] = atof(optarg); break;
		case 'm': domain_length[1] = atof(optarg); break;
		case 'p': domain_length[2] = atof(optarg); break;
		case 'l': loops = atoi(optarg); break;
		case 'k': kernel_count = (size_t)atoi(optarg); break;
		case 'd': device_nr = atoi(optarg); break;
		case 'r': viscosity = atof(optarg); break;
		case 'G': gravitation_y = atof(optarg); break;
		case 't': timestep = atof(optarg); break;
		case 'v': opt.debug = true; break;
		case 'g': do_visualisation = true; break;       /* per-step VTK files under output/vtk */
		case 's': case 'R': case 'T': break;
		case 'u': std::cerr << "unit tests live in tests/ (pytest); -u is not supported" << std::endl; return -1;
		case 'c': use_config_file = true; conf_file = optarg; break;
		case 1000: opt.use_double = true; break;
		case 1001: smagorinsky = atof(optarg); break;
		case 1002:
			if (!strcmp(optarg, "p2p")) opt.sync = SYNC_P2P;
			else if (!strcmp(optarg, "copy")) opt.sync = SYNC_COPY;
			else if (!strcmp(optarg, "host")) opt.sync = SYNC_HOST;
			else { usage(argv[0]); return -1; }
			break;
		case 1003: opt.beta_order = !strcmp(optarg, "linear") ? LBM_BETA_ORDER_LINEAR : LBM_BETA_ORDER_SHIPPED; break;
		case 1004: opt.dump_layout = true; break;
		case 1005: opt.dump_params = true; break;
		case 1006: opt.dump_velocity = optarg; break;
		case 1007: do_validate = true; break;
		case 1008: usage(argv[0]); return 0;
		default: usage(argv[0]); return -1;
		}
	}

#define FILL_AND_RUN(T)                                                                              \
	do {                                                                                             \
		CConfiguration<T> *cfg = Singleton<CConfiguration<T> >::Instance();                           \
		try {                                                                                         \
			if (use_config_file) cfg->loadFile(conf_file);                                            \
			else {                                                                                    \
				cfg->domain_size = CVector<3, int>(domain_size[0], domain_size[1], domain_size[2]);   \
				cfg->subdomain_num = CVector<3, int>(subdomain_nums[0], subdomain_nums[1], subdomain_nums[2]); \
				cfg->domain_length = CVector<3, T>((T)domain_length[0], (T)domain_length[1], (T)domain_length[2]); \
				cfg->gravitation = CVector<3, T>((T)0, (T)gravitation_y, (T)0);                       \
				cfg->viscosity = (T)viscosity;                                                        \
				cfg->computation_kernel_count = kernel_count;                                         \
				cfg->device_nr = device_nr;                                                           \
				cfg->do_visualization = do_visualisation;                                             \
				cfg->timestep = (T)timestep;                                                          \
				cfg->loops = loops;                                                                   \
				cfg->do_validate = do_validate;                                                       \
				cfg->drivenCavityVelocity = CVector<4, T>((T)100, (T)0, (T)0, (T)1);                  \
			}                                                                                         \
			if (smagorinsky != 0.0) cfg->smagorinsky_constant = (T)smagorinsky;                       \
			if (do_validate) cfg->do_validate = true;                                                 \
			if (loops >= 0) cfg->loops = loops;                                                       \
			return run<T>(cfg, opt);                                                                  \
		} catch (const char *msg) {                                                                   \
			std::cerr << msg << std::endl;                                                            \
			return EXIT_FAILURE;                                                                      \
		}                                                                                             \
	} while (0)

	if (opt.use_double) FILL_AND_RUN(double);
	FILL_AND_RUN(float);
}
