// common.hpp -- hex8 helpers shared by host code (reference: include/common.hpp:26-47).
// In this implementation the device-side equivalents live in micropp_b200/csrc/fe_math.cuh; the
// host versions below exist because reference tests call them directly
// (test/test_get_elem_nodes.cpp:63-85).
#pragma once

#define CUDA_HOSTDEV

#define DIM 3
#define NPE 8
#define NVOI 6

// Connectivity of element (ex,ey,ez) on an nx*ny*nz node grid (src/common.cpp:30-41).
void get_elem_nodes(int n[8], const int nx, const int ny, const int ex, const int ey, const int ez = 0);

// Gathers the 24 dofs of an element from an interleaved (node*3+d) displacement vector (src/common.cpp:45-54).
void get_elem_displ(const double *u, double elem_disp[NPE * DIM], int nx, int ny, int ex, int ey, int ez);

// eps = B[gp] * u_e (src/common.cpp:58-72).
void get_strain(const double *u, int gp, double *strain_gp, const double bmat[NPE][NVOI][NPE * DIM], int nx, int ny,
                int ex, int ey, int ez);
