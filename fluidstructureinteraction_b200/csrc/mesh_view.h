// mesh_view.h -- internal hand-over between the assembly side (assembly.cu) and the solver side (gpupcg_vec.cuh) of the
// library: device pointers of the D equation the last gpupcg_assemble_D left on the device.  Not part of the C ABI.
#pragma once
#include "../../include/gpupcg.h"

struct gpupcg_mesh_view
{
    int device, nCells, nIF, nBF;
    const double *upper, *diag, *source;        // [nIF], [nCells], [nCells*3]  (before addBoundaryDiag / addBoundarySource)
    const double *ic, *bc;                      // internalCoeffs / boundaryCoeffs [nBF*3]
    const int *cellBfStart, *cellBf;            // cell -> boundary faces (face index - nIF), ascending
};

extern "C" int gpupcgi_mesh_view(gpupcg_mesh m, gpupcg_mesh_view* v);

// ---- the stages of one outer iteration of evolve() on the mesh side (assembly.cu), driven by gpupcg_evolve_D ----------
extern "C" int gpupcgi_mesh_bind_stream(gpupcg_mesh m, void* stream, void** previous);
extern "C" int gpupcgi_mesh_outer_begin(gpupcg_mesh m, const gpupcg_eqn_controls* c, const double* g);
extern "C" int gpupcgi_mesh_outer_end(gpupcg_mesh m, const gpupcg_eqn_controls* c, double alpha, int doRelax, int detBounds,
                                      double* hostScal6, int incremental);
extern "C" int gpupcgi_mesh_snapshot_gradDfB(gpupcg_mesh m);
extern "C" int gpupcgi_mesh_D(gpupcg_mesh m, double** D);
extern "C" void gpupcgi_mesh_set_error(gpupcg_mesh m, const char* msg);
