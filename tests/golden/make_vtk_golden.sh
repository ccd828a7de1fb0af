#!/bin/bash
# Regenerates tests/golden/vtk_{f32,f64}.3.7.vtk with the REFERENCE's own VTK writer
# (/root/reference/src/libvis/CLbmVisualizationVTK.hpp + VTK_Common.cpp, compiled where they lie)
# fed by tests/cpp/vtk_mock_solver.hpp.  The reference's CLbmSolver.hpp (OpenCL) is kept out by
# pre-defining its include guard; only this container has /root/reference.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=/root/reference/src
TMP=$(mktemp -d)
cat > $TMP/main.cpp <<'EOC'
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include "libmath/CVector.hpp"
#include "CDomain.hpp"
#include "vtk_mock_solver.hpp"
#include "libvis/CLbmVisualizationVTK.hpp"
template <typename T> static void go(const char *prefix)
{
	CLbmSolver<T> solver(CVector<3, int>(6, 5, 4), CVector<3, int>(12, 0, 8), (T)0.1 / (T)96);
	CLbmVisualizationVTK<T> vis(3, prefix);
	vis.setup(&solver);
	vis.render(7);
}
int main(int argc, char **argv)
{
	if (argc > 2 && !std::strcmp(argv[2], "double")) go<double>(argv[1]); else go<float>(argv[1]);
	return 0;
}
EOC
g++ -O1 -ffp-contract=off -w -DCLBMOPENCL_HH -I$REF -I$HERE/../cpp $TMP/main.cpp $REF/libvis/VTK_Common.cpp -o $TMP/refvtk
$TMP/refvtk $HERE/vtk_f32
$TMP/refvtk $HERE/vtk_f64 double
rm -rf $TMP
ls -la $HERE/vtk_f32.3.7.vtk $HERE/vtk_f64.3.7.vtk
