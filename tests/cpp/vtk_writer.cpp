/* Drives this repository's VTK writer with the mock solver: vtk_writer <out-prefix> [double] */
#include <cstddef>
#include <cstring>
#include "../../turbulent_lbm_multigpu_b200/host/CVector.hpp"
#include "../../turbulent_lbm_multigpu_b200/host/CDomain.hpp"
#include "vtk_mock_solver.hpp"
#include "../../turbulent_lbm_multigpu_b200/host/CLbmVisualizationVTK.hpp"

template <typename T>
static void go(const char *prefix)
{
	CLbmSolver<T> solver(CVector<3, int>(6, 5, 4), CVector<3, int>(12, 0, 8), (T)0.1 / (T)96);
	CLbmVisualizationVTK<T> vis(3, prefix);
	vis.setup(&solver);
	vis.render(7);
}

int main(int argc, char **argv)
{
	if (argc < 2) return 2;
	if (argc > 2 && !std::strcmp(argv[2], "double")) go<double>(argv[1]);
	else go<float>(argv[1]);
	return 0;
}
