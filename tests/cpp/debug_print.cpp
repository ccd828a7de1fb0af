/* Prints host/CLbmDebug.hpp's dumps of small synthetic arrays: debug_print [double] */
#include <cstring>
#include <vector>
#include "../../turbulent_lbm_multigpu_b200/host/CLbmDebug.hpp"

template <typename T>
static void go()
{
	const size_t cells = 2 * 3 * 2;
	std::vector<T> dd(19 * cells), vel(3 * cells), rho(cells);
	std::vector<int> fl(cells);
	for (size_t a = 0; a < dd.size(); a++) dd[a] = (T)((int)((unsigned)(a * 2654435761u) >> 9) % 4001 - 2000) / (T)7919;
	for (size_t a = 0; a < vel.size(); a++) vel[a] = (T)((int)((unsigned)(a * 40503u) >> 2) % 2001 - 1000) / (T)100000;
	for (size_t a = 0; a < rho.size(); a++) rho[a] = (T)1 + (T)((int)a - 5) / (T)3000;
	for (size_t a = 0; a < fl.size(); a++) fl[a] = 1 << (int)((a * 7u) % 4);
	fl[3] = -2;   /* bytes above 127 print as negative chars */
	std::cout << 0.123456789 << std::endl;          /* stream state before ... */
	lbm_debug::debugPrint(std::cout, dd.data(), vel.data(), rho.data(), fl.data(), cells);
	std::cout << "DD 5" << std::endl;
	lbm_debug::debugDD(std::cout, dd.data(), cells, 5, 16, 2 * 3);
	lbm_debug::debugDD(std::cout, dd.data(), cells, 0);
	std::cout << 0.123456789 << std::endl;          /* ... and after: restored */
}

int main(int argc, char **argv)
{
	if (argc > 1 && !std::strcmp(argv[1], "double")) go<double>();
	else go<float>();
	return 0;
}
