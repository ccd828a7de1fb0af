/* A GPU-free stand-in for CLbmSolver<T> with the members the VTK writers touch
 * (domain_cells, d_cell_length, domain, storeVelocity/storeDensity/storeFlags).  Deterministic
 * synthetic fields; shared by tests/cpp/vtk_writer.cpp (this repository's writer) and
 * tests/golden/make_vtk_golden.sh (the reference's writer) so both see identical input. */
#ifndef LBM_TESTS_VTK_MOCK_SOLVER_HPP
#define LBM_TESTS_VTK_MOCK_SOLVER_HPP
template <typename T>
class CLbmSolver {
public:
	CVector<3, int> domain_cells;
	T d_cell_length;
	CDomain<T> domain;
	CLbmSolver(CVector<3, int> size, CVector<3, int> origin, T cell)
		: domain_cells(size), d_cell_length(cell), domain(3, size, origin, CVector<3, T>(cell * size[0], cell * size[1], cell * size[2])) {}
	size_t n() { return (size_t)domain_cells[0] * domain_cells[1] * domain_cells[2]; }
	void storeVelocity(T *dst) { for (size_t a = 0; a < 3 * n(); a++) dst[a] = (T)(((int)((a * 2654435761u) >> 7) % 20001 - 10000) * 1.0e-5); }
	void storeDensity(T *dst) { for (size_t a = 0; a < n(); a++) dst[a] = (T)1.0 + (T)(((int)((a * 40503u) >> 3) % 2001 - 1000) * 1.0e-6); }
	void storeFlags(int *dst) { for (size_t a = 0; a < n(); a++) dst[a] = 1 << (int)((a * 7u) % 4); }
};
#endif
