/*
 * CLbmVisualizationVTK.hpp -- per-rank, per-step ASCII legacy-VTK dump of density, flags and
 * velocity: the consumer side of the solver's output path (storeVelocity / storeDensity /
 * storeFlags).  Writes byte-for-byte what the reference's writer produces
 * (src/libvis/CLbmVisualizationVTK.hpp:27-99, VTK_Common.cpp:9-21, VTK_Common.hpp:30-45):
 * file "<name>.<uid>.<step>.vtk", STRUCTURED_GRID of (S+1)^3 points spaced by the cell length
 * from the sub-domain origin, then CELL_DATA density / flag / velocity in cell order, all "%f".
 * tests/test_vtk_output.py pins it against a file produced by the reference's own writer.
 *
 * Generic in the solver type so that it can be exercised without a GPU; the default is the
 * CLbmSolver<T> facade.
 */
#ifndef LBM_B200_HOST_CLBMVISUALIZATIONVTK_HPP
#define LBM_B200_HOST_CLBMVISUALIZATIONVTK_HPP

#include <cstdio>
#include <string>
#include <vector>

#include "CDomain.hpp"
#include "CVector.hpp"

template <typename T> class CLbmSolver;

/* src/libvis/ILbmVisualization.hpp: owns host copies of the three fields */
template <typename T, typename Solver = CLbmSolver<T> >
class ILbmVisualization {
protected:
	std::vector<T> velocity, density;
	std::vector<int> flags;
	Solver *cLbmOpencl;

public:
	ILbmVisualization() : cLbmOpencl(NULL) {}
	virtual ~ILbmVisualization() {}
	virtual void setup(Solver *solver)
	{
		cLbmOpencl = solver;
		const size_t n = (size_t)solver->domain_cells.elements();
		velocity.assign(n * 3, T(0));
		density.assign(n, T(0));
		flags.assign(n, 0);
	}
	virtual void render(int increment = -1) = 0;
};

template <typename T, typename Solver = CLbmSolver<T> >
class CLbmVisualizationVTK : public ILbmVisualization<T, Solver> {
	int _UID;
	std::string _file_name;
	int _timeStepNumber;

public:
	CLbmVisualizationVTK(int UID, std::string file_name) : _UID(UID), _file_name(file_name), _timeStepNumber(-1) {}

	void render(int increment = -1)
	{
		_timeStepNumber = increment;       /* the loop counter names the file */
		Solver *s = this->cLbmOpencl;
		const int nx = s->domain_cells[0], ny = s->domain_cells[1], nz = s->domain_cells[2];
		const size_t cells = (size_t)nx * ny * nz;
		s->storeVelocity(this->velocity.data());
		s->storeDensity(this->density.data());
		s->storeFlags(this->flags.data());

		const std::string path = _file_name + "." + std::to_string(_UID) + "." + std::to_string(_timeStepNumber) + ".vtk";
		FILE *fp = std::fopen(path.c_str(), "w");
		if (!fp) { std::fprintf(stderr, "Failed to open %s", path.c_str()); return; }

		std::fputs("# vtk DataFile Version 3.1\nTurbulent Fluid Simulation on MultiGPU.\nASCII\n\n", fp);
		std::fputs("DATASET STRUCTURED_GRID\n", fp);
		std::fprintf(fp, "DIMENSIONS  %i %i %i \n", nx + 1, ny + 1, nz + 1);
		std::fprintf(fp, "POINTS %i float\n\n", (nx + 1) * (ny + 1) * (nz + 1));
		/* grid points: cubic cells of edge d_cell_length, offset by the sub-domain origin; the
		 * products and sums are evaluated in T like the reference's template */
		const T h = s->d_cell_length;
		const CVector<3, int> org = s->domain.getOrigin();
		const T ox = org[0] * h, oy = org[1] * h, oz = org[2] * h;
		for (int k = 0; k <= nz; k++)
			for (int j = 0; j <= ny; j++)
				for (int i = 0; i <= nx; i++)
					std::fprintf(fp, "%f %f %f\n", ox + (i * h), oy + (j * h), oz + (k * h));

		std::fprintf(fp, "\nCELL_DATA %i \n", (int)cells);
		std::fputs("SCALARS density float 1 \nLOOKUP_TABLE default \n", fp);
		for (size_t a = 0; a < cells; a++) std::fprintf(fp, "%f\n", this->density[a]);
		std::fputs("\nSCALARS flag INT 1 \nLOOKUP_TABLE default \n", fp);
		for (size_t a = 0; a < cells; a++) std::fprintf(fp, "%i\n", this->flags[a]);
		std::fputs("\nVECTORS velocity float\n", fp);
		const T *vx = this->velocity.data(), *vy = vx + cells, *vz = vy + cells;
		for (size_t a = 0; a < cells; a++) std::fprintf(fp, "%f %f %f\n", vx[a], vy[a], vz[a]);
		if (std::fclose(fp)) std::fprintf(stderr, "Failed to close %s", path.c_str());
	}
};

#endif
