/*
 * common.h -- constants shared by the host facade (reference src/common.h:19-31,59-67 and the
 * D3Q19 table of src/main.cpp:38-66).
 */
#ifndef LBM_B200_HOST_COMMON_H
#define LBM_B200_HOST_COMMON_H

#include "../../include/lbm_b200.h"
#include "CVector.hpp"

#define FLAG_OBSTACLE            LBM_FLAG_OBSTACLE
#define FLAG_FLUID               LBM_FLAG_FLUID
#define FLAG_VELOCITY_INJECTION  LBM_FLAG_VELOCITY_INJECTION
#define FLAG_GHOST_LAYER         LBM_FLAG_GHOST_LAYER

#define BENCHMARK_OUTPUT_DIR "output/benchmark"
#define PROFILE_OUTPUT_DIR   "output/profile"
#define VTK_OUTPUT_DIR       "output/vtk"
#define LOG_OUTPUT_DIR       "output/log"

#define MPI_TAG_ALPHA_SYNC 0
#define MPI_TAG_BETA_SYNC  1

/* lattice vectors in slot order; slot f^1 is the opposite direction for f < 18 */
static const int lbm_units[19][3] = {
	{ 1, 0, 0 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, -1, 0 },
	{ 1, 1, 0 }, { -1, -1, 0 }, { 1, -1, 0 }, { -1, 1, 0 },
	{ 1, 0, 1 }, { -1, 0, -1 }, { 1, 0, -1 }, { -1, 0, 1 },
	{ 0, 1, 1 }, { 0, -1, -1 }, { 0, 1, -1 }, { 0, -1, 1 },
	{ 0, 0, 1 }, { 0, 0, -1 }, { 0, 0, 0 } };

#endif
