/*
 * gpupcg.h -- C ABI of libgpupcg_cuda.so, the B200 (sm_100a) implementation of
 * the segregated finite-volume displacement-equation solve that stressFoam,
 * thermalStressFoam and the solid side of fsiFoam share.
 *
 * This header is the drop-in boundary.  Nothing in it is a foam-extend or a
 * torch type: plain pointers, sizes and PODs, callable from C, C++, ctypes.
 * Every entry point replaces one foam-extend-3.1 interface that the reference
 * (chnrdu/fluidStructureInteraction) reaches through DEqn.solve():
 *
 *   reference call site                                  replaced interface [EXT = foam-extend-3.1]
 *   ---------------------------------------------------  --------------------------------------------
 *   unsTotalLagrangianStress.C:908  DEqn.solve()          fvMatrix<vector>::solve -> lduMatrix::solver::New
 *   unsTotalLagrangianStressSolve.C:378,560               (same)
 *   thermalStressFoam/TEqn.H:15     TEqn.solve()          fvScalarMatrix::solve   -> lduMatrix::solver::New
 *   consistentIcoFlow.C:224         pEqn.solve()          (same)
 *   run/.../system/fvSolution  "solver PCG; preconditioner DIC;"   PCG::solve + DICPreconditioner
 *
 * (paths relative to /root/reference/src/fluidStructureInteraction/stressModels/
 *  unsTotalLagrangianStress/ unless they start with a directory name.)
 *
 * Threading: one handle per (process, GPU, lduAddressing).  Calls on a handle
 * are synchronous with respect to the caller unless stated; a handle is not
 * re-entrant (foam-extend calls a solver from one thread per MPI rank).
 * Errors: every function returns 0 (GPUPCG_OK) or a negative code; the text is
 * available from gpupcg_last_error().  No C++ exception crosses this boundary.
 * There is no CPU fallback: without a usable sm_100 device gpupcg_create fails.
 */
#ifndef GPUPCG_H
#define GPUPCG_H

#ifdef __cplusplus
extern "C" {
#endif

#define GPUPCG_OK              0
#define GPUPCG_ERR_ARG        -1   /* NULL / negative size / inconsistent arguments          */
#define GPUPCG_ERR_ORDERING   -2   /* lowerAddr/upperAddr not in upper-triangular order      */
#define GPUPCG_ERR_CUDA       -3   /* CUDA runtime error (text in gpupcg_last_error)         */
#define GPUPCG_ERR_NODEVICE   -4   /* no CUDA device / not an sm_100 class device            */
#define GPUPCG_ERR_STATE      -5   /* call sequence error (e.g. solve before set_matrix)     */
#define GPUPCG_ERR_NCCL       -6   /* communicator error                                     */
#define GPUPCG_ERR_UNSUPPORTED -7  /* asymmetric matrix, unknown preconditioner, ...         */

typedef struct gpupcg_handle_s* gpupcg_handle;

/* Mirror of lduMatrix::solverPerformance [EXT lduMatrix.H] plus timing. */
typedef struct gpupcg_perf
{
    double initialResidual;
    double finalResidual;
    double normFactor;
    int    nIterations;
    int    converged;
    int    singular;
    int    reserved;
    double solve_ms;        /* device time of the solve (CUDA events on the handle's stream) */
    double h2d_ms;          /* host->device staging inside the call (0 for *_device entry points) */
    double d2h_ms;
} gpupcg_perf;

/* Static description of what create() built: the wavefront schedule. */
typedef struct gpupcg_plan_info
{
    int       nCells;
    int       nFaces;
    int       nRows;            /* rows after per-level padding to the warp width  */
    int       nSlices;          /* nRows / 32                                      */
    int       nLevels;          /* length of the DIC dependency chain              */
    int       maxLevelRows;
    long long entriesUpper;     /* SELL-32 storage incl. padding, owner side       */
    long long entriesLower;     /* SELL-32 storage incl. padding, neighbour side   */
    int       nInterfaceFaces;  /* coupled patch faces                             */
    int       nInterfaces;
    int       sweepGrid;        /* CTAs of the DIC sweep kernels (= nTiles)        */
    int       streamGrid;       /* CTAs of the streaming kernels                   */
    int       numSMs;
    int       nTiles;           /* row tiles of the DIC sweeps (one CTA each)      */
    int       tileRowsMax;
    int       maxRounds;        /* longest tile-local dependency chain             */
    int       sweepSmemBytes;   /* dynamic shared memory of a sweep CTA            */
    int       extForward;       /* cross-tile dependencies, forward / backward     */
    int       extBackward;
    int       roundEntries;     /* length of the roundStart table                  */
} gpupcg_plan_info;

/* ---- life cycle ----------------------------------------------------------------------------
 * gpupcg_create: analogue of constructing lduAddressing-derived data once per mesh
 * ([EXT] lduAddressing::{ownerStartAddr,losortAddr,losortStartAddr} + the DIC level schedule).
 *   lowerAddr/upperAddr : lduAddr().lowerAddr()/upperAddr(), int32, upper-triangular order.
 *   nPatches            : number of COUPLED interfaces (processor patches) of this sub-domain.
 *   patchSizes[p]       : faces in interface p;  patchFaceCells[p] : lduAddr().patchAddr(p).
 *   patchNbrRank[p]     : rank owning the other side (processorFvPatch::neighbProcNo()).
 * The host arrays are copied; nothing is retained.                                            */
int gpupcg_create(gpupcg_handle* h, int device,
                  int nCells, int nFaces,
                  const int* lowerAddr, const int* upperAddr,
                  int nPatches, const int* patchSizes,
                  const int* const* patchFaceCells, const int* patchNbrRank);
int gpupcg_destroy(gpupcg_handle h);
const char* gpupcg_last_error(gpupcg_handle h);   /* h may be NULL: last create() error */
int gpupcg_get_plan_info(gpupcg_handle h, gpupcg_plan_info* info);

/* Debug (GPUPCG_TILE_TRACE=1 at create): {CTA start, staging done, sweep done} %globaltimer stamps of every
 * tile of the last forward sweep, 3*nTiles values.                                                       */
int gpupcg_debug_tile_trace(gpupcg_handle h, unsigned long long* out3n);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches evidence). */
int gpupcg_launch_count(gpupcg_handle h, long long* count);

/* Run all later work of this handle on an externally owned cudaStream_t (e.g. torch's
 * current stream, so that the caller's CUDA events bracket the kernels).  NULL restores the
 * handle's own stream.                                                                       */
int gpupcg_set_stream(gpupcg_handle h, void* cudaStream);

/* ---- matrix ---------------------------------------------------------------------------------
 * gpupcg_set_matrix: the lduMatrix the solver object was constructed with:
 *   diag = matrix.diag() (after fvMatrix::addBoundaryDiag for this component),
 *   upper = matrix.upper(); lower must be NULL or == upper (symMatrix table only),
 *   patchBouCoeffs[p] = coupleBouCoeffs[p] (interface coefficients used by Amul).
 * upper == NULL keeps the coefficients of the previous call (the three components of one
 * fvMatrix<vector>::solve share upper; only diag differs).  Host pointers.                    */
int gpupcg_set_matrix(gpupcg_handle h, const double* diag, const double* upper,
                      const double* lower_or_null, const double* const* patchBouCoeffs);
/* Same with device pointers in the ORIGINAL cell / face order (output of gpupcg_assemble_*). */
int gpupcg_set_matrix_device(gpupcg_handle h, const double* d_diag, const double* d_upper,
                             const double* d_patchBouCoeffsConcat);

/* ---- solve ----------------------------------------------------------------------------------
 * gpupcg_solve == PCG::solve(x, b, cmpt) with the DIC preconditioner [EXT PCG.C,
 * DICPreconditioner.C]; controls are lduMatrix::solver::readControls' tolerance, relTol,
 * minIter, maxIter.  x is in/out, b is const, both double[nCells] on the HOST in the caller's
 * cell order.  history (optional, host) receives finalResidual after every iteration, up to
 * historyCap entries.  Multi-rank: collective over the communicator given to comm_init.       */
int gpupcg_solve(gpupcg_handle h, double* x, const double* b,
                 double tolerance, double relTol, int minIter, int maxIter,
                 gpupcg_perf* perf, double* history_or_null, int historyCap);
/* x, b are DEVICE pointers in the original cell order. */
int gpupcg_solve_device(gpupcg_handle h, double* d_x, const double* d_b,
                        double tolerance, double relTol, int minIter, int maxIter,
                        gpupcg_perf* perf, double* history_or_null, int historyCap);

/* ---- kernels in isolation (parity tests and roofline runs) -----------------------------------
 * gpupcg_amul       : lduMatrix::Amul(Ax, x, coupleBouCoeffs, interfaces, cmpt)  [EXT lduMatrixATmul.C]
 *                     xNbr_or_null = concatenated received patch-neighbour values (single-rank tests).
 * gpupcg_dic_factor : DICPreconditioner::calcReciprocalD -> rD                    [EXT DICPreconditioner.C]
 * gpupcg_precondition: DICPreconditioner::precondition(wA, rA)
 * All pointers are HOST arrays in the caller's order.                                          */
int gpupcg_amul(gpupcg_handle h, const double* x, double* Ax, const double* xNbr_or_null);
int gpupcg_dic_factor(gpupcg_handle h, double* rD_out);
int gpupcg_precondition(gpupcg_handle h, const double* rA, double* wA);

/* Time `reps` back-to-back launches of one kernel class on resident data with CUDA events on
 * the handle's stream; flushL2 != 0 writes a 256 MiB scratch buffer before every launch
 * (excluded from the time).  which: 0 Amul, 1 DIC precondition (forward+backward),
 * 2 pA update, 3 x/r update + residual norm, 4 DIC factor.  Returns mean ms per launch.        */
int gpupcg_time_kernel(gpupcg_handle h, int which, int reps, int flushL2, double* ms_per_launch);

/* ---- component-batched solve ------------------------------------------------------------------
 * fvMatrix<vector>::solve [EXT fvMatrixSolve.C], as DEqn.solve() reaches it (unsTotalLagrangianStress.C:908;
 * unsTotalLagrangianStressSolve.C:378,560), loops over the valid components: diag + internalCoeffs.component(cmpt),
 * source.component(cmpt), one lduMatrix::solver::New(...)->solve(psiCmpt, sourceCmpt, cmpt) each.  The nCmpt (2 or 3)
 * systems share upper() and lduAddr(), differ in diag and source, and are independent -- and the DIC sweeps of ONE
 * system are bound by the latency of their dependency chain.  These entry points take the whole loop: every kernel
 * advances all components in the same launch (kernels_vec.cuh), each component runs the reference's PCG recurrence
 * with its own scalars and stops on its own (results per component = those of gpupcg_solve on that component).
 *   diagCmpt            : [nCells*nCmpt], foam's Field<vector> layout (component-minor): matrix.diag() after
 *                         addBoundaryDiag(diag, cmpt) for every cmpt
 *   upper               : matrix.upper() (NULL keeps the previous coefficients)
 *   patchBouCoeffsCmpt  : per coupled patch [size*nCmpt] = coupleBouCoeffs (boundaryCoeffs.component(cmpt))
 *   x, b                : [nCells*nCmpt] = psi.internalField() (in/out) and the total source, same layout
 *   perfCmpt            : [nCmpt];  history: [nCmpt*historyCap] (component-major) or NULL
 * Meshes with more than four faces per row and side (polyhedra) and multi-rank runs without peer-memory mailboxes
 * run the components one after the other on the scalar kernels (gpupcg_vector_mode returns 0 then, 1 when batched).
 * Single-rank structured-block plans of 100 k ... 10 M cells run the components as nCmpt scalar solves IN FLIGHT AT ONCE
 * (one stream, host thread and set of vectors per component inside the library; plan tables and upper() shared; every
 * component is then exactly gpupcg_solve of that system) -- gpupcg_vector_mode returns 2; GPUPCG_VECTOR_CONCURRENT=0/1
 * overrides the choice.
 * Multi-rank: the ranks agree on the path with one all-reduce the first time it is needed, so gpupcg_vector_mode and
 * the first gpupcg_solve_vector after gpupcg_comm_init are collective calls.                                        */
int gpupcg_vector_mode(gpupcg_handle h, int nCmpt);
int gpupcg_set_matrix_vector(gpupcg_handle h, int nCmpt, const double* diagCmpt, const double* upper,
                             const double* const* patchBouCoeffsCmpt);
int gpupcg_set_matrix_vector_device(gpupcg_handle h, int nCmpt, const double* d_diagCmpt, const double* d_upper,
                                    const double* d_patchBouCoeffsCmptConcat);
int gpupcg_solve_vector(gpupcg_handle h, int nCmpt, double* x, const double* b,
                        double tolerance, double relTol, int minIter, int maxIter,
                        gpupcg_perf* perfCmpt, double* history_or_null, int historyCap);
int gpupcg_solve_vector_device(gpupcg_handle h, int nCmpt, double* d_x, const double* d_b,
                               double tolerance, double relTol, int minIter, int maxIter,
                               gpupcg_perf* perfCmpt, double* history_or_null, int historyCap);
/* gpupcg_time_kernel for the batched kernels (same `which` codes; one launch advances nCmpt systems). */
int gpupcg_time_kernel_vector(gpupcg_handle h, int nCmpt, int which, int reps, int flushL2, double* ms_per_launch);

/* ---- multi-GPU -------------------------------------------------------------------------------
 * One handle per rank/GPU (decomposePar sub-domain).  ncclUniqueId is the 128-byte
 * ncclUniqueId obtained on rank 0 (gpupcg_comm_unique_id) and broadcast by the host side
 * (Pstream / torch.distributed / a file).  Halo swaps (processorFvPatchField::
 * initInterfaceMatrixUpdate/updateInterfaceMatrix) and gSumProd/gSumMag/gAverage reductions
 * then run over NCCL on the handle's stream.                                                   */
int gpupcg_comm_unique_id(void* ncclUniqueId_128bytes);
int gpupcg_comm_init(gpupcg_handle h, int rank, int nRanks, const void* ncclUniqueId_128bytes);

/* ---- GPU assembly of the D equation ---------------------------------------------------------------
 * What unsTotalLagrangianStress::evolve() builds per outer iteration before DEqn.solve()
 * (unsTotalLagrangianStress.C:846-859, steadyState d2dt2, linear elasticity):
 *   fvm::laplacian(2*muf+lambdaf, D) with the reference's `skewCorrected` snGrad scheme
 *   (skewCorrectedVectorSnGrad.C:50-294), fvc::div(Sf & (...gradDf...)), rho*g, and the boundary
 *   coefficients of tractionDisplacement (patchType 0) / fixedDisplacement (1) / symmetryDisplacement (2) patches
 *   (fvPatchFields/{tractionDisplacement,fixedDisplacement,symmetryDisplacement}/ *.C).
 * gpupcg_mesh_create uploads what foam-extend's fvMesh provides once per mesh: owner/neighbour, the patch
 * table and mesh.Sf(), magSf(), Cf(), C(), V(), weights(), deltaCoeffs() (boundary faces included).
 * gpupcg_assemble_D: per-cell mu, lambda, rho, D[3], gradD[9]; per-face gradDf[9]; per boundary face
 * traction[3], pressure, fixedValue[3]; g[3]; limitCoeff of the snGrad scheme.  Outputs (host, the caller's
 * numbering): upper[nInternalFaces], diag[nCells] (before addBoundaryDiag), source[nCells*3] (before
 * addBoundarySource), internalCoeffs / boundaryCoeffs[nBoundaryFaces*3].  reps > 1 repeats the three kernels
 * for timing; kernel_ms (optional) = mean device time of one assembly.
 * gpupcg_sigma: epsilon = symm(gradD), sigma = 2 mu eps + lambda tr(eps) I (symmTensor order xx xy xz yy yz zz). */
typedef struct gpupcg_mesh_s* gpupcg_mesh;
int gpupcg_mesh_create(gpupcg_mesh* m, int device, int nCells, int nFaces, int nInternalFaces,
                       const int* owner, const int* neighbour, int nPatches, const int* patchStart,
                       const int* patchSize, const int* patchType, const double* Sf, const double* magSf,
                       const double* Cf, const double* C, const double* V, const double* weights,
                       const double* deltaCoeffs);
int gpupcg_mesh_destroy(gpupcg_mesh m);
const char* gpupcg_mesh_last_error(gpupcg_mesh m);
int gpupcg_assemble_D(gpupcg_mesh m, const double* mu, const double* lambda, const double* rho, const double* g,
                      const double* D, const double* gradD, const double* gradDf, const double* traction,
                      const double* pressure, const double* fixedValue, double limitCoeff, int reps,
                      double* upper, double* diag, double* source, double* internalCoeffs,
                      double* boundaryCoeffs, double* kernel_ms);
int gpupcg_sigma(gpupcg_mesh m, const double* mu, const double* lambda, const double* gradD,
                 double* epsilon, double* sigma);
/* Vertex-based gradients of the reference (numerics/fvc/fvcGradf.C): gradD = fvc::grad(D, pointD) (:485-717, boundary
 * field in gradDb) followed by gradDf = fvc::fGrad(D, pointD) (:47-231) exactly as evolve() calls them
 * (unsTotalLagrangianStress.C:925-926).  gpupcg_mesh_set_points uploads mesh.points() and the face point loops.
 * gradDold = the registered grad(D) at call time; patchGradient[nBoundaryFaces*3] = fixedGradient values of the
 * tractionDisplacement patches (what their last updateCoeffs set).                                              */
/* gradient() of the tractionDisplacement patches as the last gpupcg_assemble_D left it on the device
 * (tractionDisplacementFvPatchVectorField::updateCoeffs, :148-257); gpupcg_gradients uses it when its patchGradient
 * argument is NULL -- the order of unsTotalLagrangianStress::evolve(): assemble, solve, then the gradients.         */
int gpupcg_patch_gradient(gpupcg_mesh m, double* patchGradient);
/* D.relax(alpha) (unsTotalLagrangianStress.C:910; skipped when relax == 0: no relaxationFactor for D in fvSolution)
 * followed by unsTotalLagrangianStress::residual() (unsTotalLagrangianStressSolve.C:647-672):
 *   maxDD = gMax(mag(D - D.oldTime()));  residual = gMax(mag(D - D.prevIter())/(maxDD + SMALL)).
 * D[nCells*3] in/out (host), Dprev = D.prevIter(), Dold = D.oldTime().                                              */
int gpupcg_relax_residual(gpupcg_mesh m, double alpha, int relax, double* D, const double* Dprev, const double* Dold,
                          double* residual);
int gpupcg_mesh_set_points(gpupcg_mesh m, int nPoints, const double* points, const int* faceStart, const int* facePoints);
/* Residency across the outer iterations of evolve(): grad(D) and the face gradient (72 B per cell / per face) stay on the
 * device between gpupcg_gradients (:925-926) and the next gpupcg_assemble_D (:846-859) instead of crossing PCIe twice.
 *   gpupcg_assemble_D : gradD == NULL / gradDf == NULL -> what the device holds (from gpupcg_gradients, or the last
 *                       upload);
 *   gpupcg_gradients  : gradDold == NULL -> the device's current grad(D) is copied aside as the registered gradient;
 *                       gradD / gradDb / gradDf == NULL -> not copied back;
 *   gpupcg_sigma      : gradD == NULL -> the device's grad(D);
 *   gpupcg_mesh_fetch : copies a device-resident field out on demand: which = 0 D [nCells*3], 1 gradD [nCells*9],
 *                       2 gradDf [nFaces*9], 3 gradDb [nBoundaryFaces*9].
 * GPUPCG_ERR_STATE if the device holds nothing yet.                                                                    */
int gpupcg_mesh_fetch(gpupcg_mesh m, int which, double* out);
/* DEqn.solve() (unsTotalLagrangianStress.C:908) on the D equation the last gpupcg_assemble_D left on the device, without
 * the matrix crossing PCIe: fvMatrix::addBoundaryDiag / addBoundarySource of the (non-coupled) patches [EXT fvMatrix.C]
 * are applied on the device in fvMatrix's order, then the three components are solved by the batched PCG -- the same
 * bits as gpupcg_assemble_D -> host -> gpupcg_set_matrix_vector + gpupcg_solve_vector.  h: a solver handle created
 * with this mesh's owner/neighbour (internal faces) addressing on the same device; D [nCells*3] host, in/out;
 * perfCmpt [3].  gpupcg_assemble_D may be called with NULL output pointers then (nothing is copied back).            */
int gpupcg_solve_assembled_D(gpupcg_handle h, gpupcg_mesh m, double* D, double tolerance, double relTol, int minIter,
                             int maxIter, gpupcg_perf* perfCmpt);
int gpupcg_gradients(gpupcg_mesh m, const double* D, const double* pointD, const double* gradDold,
                     const double* fixedValue, const double* patchGradient, double limitCoeff, int reps,
                     double* gradD, double* gradDb, double* gradDf, double* kernel_ms);

/* ---- the whole outer loop of evolve() on resident fields (round 2) ---------------------------------------------------
 * unsTotalLagrangianStress::evolve() (unsTotalLagrangianStress.C:769-1009) with D, pointD, grad(D), gradDf, D.oldTime(),
 * D.boundaryField() ... resident on the device across outer iterations AND time steps: nothing but six scalars per
 * outer iteration crosses PCIe.  Patch types of gpupcg_mesh_create: 0 tractionDisplacement, 1 fixedDisplacement,
 * 2 symmetryDisplacement (fvPatchFields/ of the reference), 3 symmetryPlane [EXT symmetryFvPatchField], 4 empty (2-D).
 *
 * gpupcg_eqn_controls: what evolve() reads from fvSchemes / stressProperties for the equation itself
 *   d2dt2Scheme / ddtScheme  0 steadyState, 1 Euler, 2 backward  (d2dt2(D), ddt(D) of fvSchemes; deltaT0 = runTime.deltaT0())
 *   eD2 ... eU               `backward` only (numerics/backwardD2dt2Scheme/backwardD2dt2Scheme.C:230-280 over [EXT]
 *                            backwardDdtScheme): the EFFECTIVE old time step of every use -- the real deltaT0, or GREAT (1e15)
 *                            while the field has too few distinct old-time levels (backwardD2dt2Scheme.C:58-76, [EXT]
 *                            backwardDdtScheme::deltaT0_: vf.nOldTimes() < 2).  The caller owns D's time levels
 *                            (GPUPCG_F_DOLD ... GPUPCG_F_DOOOO, gpupcg_mesh_copy_field) and knows their state:
 *                            eD2 fvmD2dt2(D); eDdtD the fvmDdt(D) inside it; eDdtD0 / eDdtD00 the fvcDdt(D.oldTime()) /
 *                            fvcDdt(D.oldTime().oldTime()) inside it; eDamp the damping term's fvm::ddt(D); eU U = fvc::ddt(D)
 *   K                        damping coefficient (stressProperties "K", :800-804), active when > SMALL (:862)
 *   nonLinear                stressProperties "nonLinear" (:796); thermal: stressModel::thermalStress() (:815)
 *   limitCoeff               snGrad scheme "skewCorrected <limitCoeff>" of fvSchemes                                      */
typedef struct gpupcg_eqn_controls
{
    int    d2dt2Scheme, ddtScheme;
    double deltaT, deltaT0;
    double K;
    int    nonLinear, thermal;
    double limitCoeff;
    double eD2, eDdtD, eDdtD0, eDdtD00, eDamp, eU;
} gpupcg_eqn_controls;

/* Field ids of gpupcg_mesh_set_field / gpupcg_mesh_get_field (host <-> resident device field, caller's numbering).
 * Cell fields [nCells*k], face fields [nFaces*k], boundary fields [nBoundaryFaces*k] (face index - nInternalFaces).   */
#define GPUPCG_F_MU              0   /* rheology mu()              [nCells]                    */
#define GPUPCG_F_LAMBDA          1
#define GPUPCG_F_RHO             2
#define GPUPCG_F_D               3   /* D internal field           [nCells*3]                  */
#define GPUPCG_F_DOLD            4   /* D.oldTime()                                            */
#define GPUPCG_F_DOLDOLD         5   /* D.oldTime().oldTime()                                  */
#define GPUPCG_F_GRADD           6   /* grad(D)                    [nCells*9]                  */
#define GPUPCG_F_GRADDF          7   /* gradDf                     [nFaces*9]                  */
#define GPUPCG_F_TRACTION        8   /* tractionDisplacement::traction()  [nBoundaryFaces*3]   */
#define GPUPCG_F_PRESSURE        9   /*                       pressure()  [nBoundaryFaces]     */
#define GPUPCG_F_FIXEDVALUE     10   /* fixedDisplacement values          [nBoundaryFaces*3]   */
#define GPUPCG_F_THREEK         11   /* threeK() = rheology.threeK()*rho  [nCells]  (:1313)    */
#define GPUPCG_F_ALPHA          12   /* thermal.alpha()                   [nCells]             */
#define GPUPCG_F_DT             13   /* DT = T - T0 (TEqn.H:34)           [nCells]             */
#define GPUPCG_F_DTB            14   /* boundary values of DT             [nBoundaryFaces]     */
#define GPUPCG_F_POINTD         15   /* pointD                            [nPoints*3]          */
#define GPUPCG_F_DB             16   /* D.boundaryField()                 [nBoundaryFaces*3]   */
#define GPUPCG_F_GRADDB         17   /* boundary field of grad(D)         [nBoundaryFaces*9]   */
#define GPUPCG_F_U              18   /* U = fvc::ddt(D)                   [nCells*3]           */
#define GPUPCG_F_EPSILON        19   /* symmTensor xx xy xz yy yz zz      [nCells*6]           */
#define GPUPCG_F_SIGMA          20
#define GPUPCG_F_UPPER          21   /* the assembled equation: upper [nInternalFaces], diag [nCells], source [nCells*3], */
#define GPUPCG_F_DIAG           22   /* internalCoeffs / boundaryCoeffs [nBoundaryFaces*3]                                */
#define GPUPCG_F_SOURCE         23
#define GPUPCG_F_INTERNALCOEFFS 24
#define GPUPCG_F_BOUNDARYCOEFFS 25
#define GPUPCG_F_PATCHGRADIENT  26   /* gradient() of the tractionDisplacement patches [nBoundaryFaces*3] */
#define GPUPCG_F_DPREV          27   /* D.prevIter()                                           */
#define GPUPCG_F_DBPREV         28
#define GPUPCG_F_DOOO           29   /* D.oldTime().oldTime().oldTime()   (`backward` d2dt2)   */
#define GPUPCG_F_DOOOO          30   /* ... and its oldTime()                                  */
#define GPUPCG_F_GRADDFB_START  31   /* boundary rows of gradDf at the START of the last outer iteration [nBoundaryFaces*9]:
                                        what DSigmaf_ of unsIncrTotalLagrangianStress::evolve() (:1013-1018) was built from
                                        (incremental loop only)                                */
int gpupcg_mesh_field_size(gpupcg_mesh m, int fieldId, long long* nDoubles);
int gpupcg_mesh_set_field(gpupcg_mesh m, int fieldId, const double* host);
int gpupcg_mesh_get_field(gpupcg_mesh m, int fieldId, double* host);

/* leastSquaresVolPointInterpolation (numerics/leastSquaresVolPointInterpolation/): the data the reference's volToPoint_
 * object holds, flattened.  Point p owns rows pointStart[p] .. pointStart[p+1]: its cells (mesh.pointCells()), then its
 * boundary faces (pointBndFaces(), leastSquaresVolPointInterpolation.C:54-205) -- entry >= 0: cell, < 0: boundary face
 * -1 - entry --; points on a mirror plane (empty patch, :1829-1940, mirror[p] != 0) carry the rows twice.  invLs[row][3]
 * is column `row` of invLsMatrices()[p] (:1425-1826), origin = origins()[p] (:1138-1421); weights() are all 1 (:1116-1133).
 * gpupcg_vol_to_point = interpolate(D, pointD) (leastSquaresVolPointInterpolate.C:43-304) on the resident D and
 * D.boundaryField(); kernel_ms (optional) = mean device time over reps launches.                                          */
int gpupcg_mesh_set_lsq(gpupcg_mesh m, const int* pointStart, const int* entry, const double* invLs, const double* origin,
                        const unsigned char* mirror);
int gpupcg_vol_to_point(gpupcg_mesh m, int reps, double* kernel_ms);
/* The point patch fields of pointD that act in pf.correctBoundaryConditions() at the end of interpolate()
 * (leastSquaresVolPointInterpolate.C:301; declared in 0/pointD, e.g. run/stressFoam/plateHole/plateHole/0/pointD:19-50,
 * run/fsiFoam/beamInCrossFlow/solid/0/pointD:19-37).  Constrained point i = point[i] (distinct, ascending) owns rows
 * start[i] .. start[i+1], applied in boundary-patch order: kind 1 symmetryPlane [EXT SymmetryPointPatchField::evaluate]
 * pf = transform(I - nHat*nHat, pf) with vec[row] = the patch's pointNormals(); kind 2 fixedValue [EXT
 * ValuePointPatchField::evaluate] pf = vec[row].  `calculated` and `empty` point patches need no entry.  nConstrained = 0
 * clears the table.  gpupcg_vol_to_point and the resident outer loop apply it after every interpolation.                  */
int gpupcg_mesh_set_point_constraints(gpupcg_mesh m, int nConstrained, const int* point, const int* start, const int* kind,
                                      const double* vec);
/* D.correctBoundaryConditions(): evaluate() of every patch field on the resident D, grad(D), gradient() -> D.boundaryField()
 * (tractionDisplacementFvPatchVectorField.C:259-300, symmetryDisplacementFvPatchVectorField.C:169-207).                   */
int gpupcg_patch_evaluate(gpupcg_mesh m);
/* dst = src on the device (two field ids of the same size): how the caller creates and shifts D's old-time levels the way
 * [EXT] GeometricField::oldTime() / storeOldTimes() do when a `backward` scheme needs more than two of them.            */
int gpupcg_mesh_copy_field(gpupcg_mesh m, int dstId, int srcId);
/* gpupcg_assemble_D / gpupcg_gradients / gpupcg_sigma on the resident fields, every branch of :846-890, :973-990
 * (gpupcg_sigma_resident: controls->nonLinear is the dictionary switch of :979, controls->ddtScheme gives U = fvc::ddt(D)). */
int gpupcg_assemble_D_resident(gpupcg_mesh m, const gpupcg_eqn_controls* controls, const double* g);
int gpupcg_gradients_resident(gpupcg_mesh m, double limitCoeff);
int gpupcg_sigma_resident(gpupcg_mesh m, const gpupcg_eqn_controls* controls);
/* ---- the other symmetric scalar systems of the boundary, assembled on the device --------------
 *   TEqn  src/solvers/thermalStressFoam/TEqn.H:9-15:   rhoC*fvm::ddt(T) == fvm::laplacian(k, T, "laplacian(k,T)")
 *         -> sign +1, ddtScheme 1, rate = rhoC [nCells], psiOld = T.oldTime(), gammaCells = k (interpolated linearly)
 *   pEqn  flowModels/consistentIcoFlow/consistentIcoFlow.C:211-224: fvm::laplacian(1/interpolate(AU), p) == fvc::div(phi);
 *         pEqn.setReference(pRefCell, pRefValue)  -> sign -1, gammaFaces, phi [nFaces], refCell / refValue
 * [EXT] gaussLaplacianScheme::fvmLaplacianUncorrected, EulerDdtScheme::fvmDdt, surfaceIntegrate, setReference, in the reference's
 * loop order (bit-exact against oracle/fvm_oracle.c: orc_assemble_scalar).  Boundary faces: bfKind 0 zeroGradient, 1 fixedValue
 * (bfValue); faces of `empty` patches of the mesh handle take no part.  The explicit correction of the scalar `skewCorrected`
 * scheme (skewCorrectedSnGrad.C:57-294, least-squares gradient) is added when psi and the least-squares vectors are given
 * (3-D meshes, non-coupled patches).  Outputs in LDU numbering,
 * ready for gpupcg_set_matrix / gpupcg_solve after fvMatrix's addBoundaryDiag / addBoundarySource (any may be NULL).            */
typedef struct gpupcg_scalar_eqn
{
    int           sign;                 /* +1: rate*ddt(psi) - laplacian(gamma, psi);  -1: laplacian(gamma, psi) - div(phi)     */
    const double* gammaCells;           /* [nCells] or NULL                                                                     */
    const double* gammaFaces;           /* [nFaces] or NULL (exactly one of the two)                                            */
    const int*    bfKind;               /* [nBoundaryFaces]                                                                     */
    const double* bfValue;              /* [nBoundaryFaces]                                                                     */
    int           ddtScheme;            /* 0 steadyState, 1 Euler                                                               */
    double        deltaT;
    const double* rate;                 /* [nCells] coefficient of the ddt term (rhoC)                                          */
    const double* psiOld;               /* [nCells] psi.oldTime()                                                               */
    const double* phi;                  /* [nFaces] face flux of the explicit divergence, or NULL                               */
    int           refCell;              /* setReference: -1 none                                                                */
    double        refValue;
    /* the explicit correction of `Gauss linear skewCorrected <limitCoeff>` (numerics/skewCorrectedSnGrad/skewCorrectedSnGrad.C:
     * 57-294 with the gradient scheme snGradCorr(psi) leastSquares; [EXT] gaussLaplacianScheme: source -= V*div(gammaMagSf*corr)):
     * psi != NULL switches it on                                                                                                  */
    const double* psi;                  /* [nCells] the current field                                                           */
    const double* lsP;                  /* [nInternalFaces*3] leastSquaresVectors::pVectors()                                   */
    const double* lsN;                  /* [nInternalFaces*3] leastSquaresVectors::nVectors()                                   */
    const double* lsB;                  /* [nBoundaryFaces*3] the patch fields of pVectors()                                    */
    double        limitCoeff;
} gpupcg_scalar_eqn;
int gpupcg_assemble_scalar(gpupcg_mesh m, const gpupcg_scalar_eqn* eqn, double* upper, double* diag, double* source,
                           double* internalCoeffs, double* boundaryCoeffs, double* kernel_ms);

/* runTime++: D.oldTime().oldTime() = D.oldTime(); D.oldTime() = D                                                          */
int gpupcg_mesh_new_time_step(gpupcg_mesh m);

typedef struct gpupcg_evolve_controls
{
    int    nCorrectors;                                     /* stressProperties (:773-788)                        */
    double convergenceTolerance, relConvergenceTolerance;
    double tolerance, relTol;                               /* fvSolution solvers { D { ... } }                   */
    int    minIter, maxIter;
    double relaxD;                                          /* fvSolution relaxationFactors D                     */
    int    doRelax;                                         /* 0: no relaxation factor for D in fvSolution        */
    int    validCmpt[3];                                    /* mesh.solutionD() != -1 per direction (2-D: one 0)  */
    double g[3];                                            /* stressProperties "g" (:806-810)                    */
    gpupcg_eqn_controls eqn;
    int    incremental;                                     /* 1: the loop of unsIncrTotalLagrangianStress::evolve()
                                                               (unsIncrTotalLagrangianStress.C:980-1159): the resident D is
                                                               the increment DD; residual() = max|DD - DD.prevIter()| /
                                                               (max|DD| + SMALL) (...Solve.C:695-716) and the loop runs
                                                               while res > convergenceTolerance (:1155-1159)       */
} gpupcg_evolve_controls;

typedef struct gpupcg_evolve_result
{
    int    nOuterIterations;            /* iCorr at exit (:967-971)                                       */
    int    totalPcgIterations;
    int    enforceLinear;               /* set when det(I + gradDf) left [0.01, 100] (:939-945)           */
    int    converged;                   /* evolve()'s return value (:1002-1008)                           */
    double initialResidual;             /* solverPerf.initialResidual() of the first outer iteration      */
    double lastSolverInitialResidual;   /* ... of the last one ("Final residual" of the summary, :994)    */
    double maxRes, res;
    double assemble_ms, solve_ms, update_ms;    /* device time: assembly | DEqn.solve() | evaluate+relax+interpolate+gradients+residual */
} gpupcg_evolve_result;

/* h: solver handle created with this mesh's internal-face addressing on the same device (no coupled patches yet).
 * residualHistory [historyCap] (optional): res of every outer iteration; pcgIterations [historyCap*3] (optional):
 * PCG iterations per valid component and outer iteration.  gpupcg_sigma_resident afterwards gives U, epsilon, sigma.   */
int gpupcg_evolve_D(gpupcg_handle h, gpupcg_mesh m, const gpupcg_evolve_controls* controls, gpupcg_evolve_result* result,
                    double* residualHistory, int* pcgIterations, int historyCap);

/* ---- host memory helpers ----------------------------------------------------------------------
 * Page-lock caller-owned arrays (scalarField storage) so that the staging copies are DMA.       */
int gpupcg_host_register(void* ptr, unsigned long long bytes);
int gpupcg_host_unregister(void* ptr);

/* Host-only (no GPU needed): build the tiling / SELL-32 schedule create() would build and copy it out, so
 * that CPU tests can check the schedule against the reference face order.  Call once with arrays == NULL to
 * obtain the sizes in *info, then with an array of GPUPCG_PLAN_NARRAYS int pointers (any may be NULL):
 *  0 rowOld[nRows]  1 sliceInfo[4*nSlices]  2 Ucol  3 UfaceOld [entriesUpper]  4 Lidx  5 Lcol [entriesLower]
 *  6 tileStart  7 tileRoundPtr [nTiles+1]  8 roundStart[roundEntries]  9 LcolTile[entriesLower]
 * 10 UcolTile[entriesUpper]  11 extFPtr[nTiles+1]  12 extFCol[extForward]  13 extBPtr[nTiles+1]
 * 14 extBCol[extBackward]                                                                              */
#define GPUPCG_PLAN_NARRAYS 15
int gpupcg_plan_export(int nCells, int nFaces, const int* lowerAddr, const int* upperAddr, int tileRows,
                       int tileMode /* 0 natural-order intervals, 1 greedy compact blobs (create() default) */,
                       gpupcg_plan_info* info, int** arrays);

/* Number of visible CUDA devices (0 if none / no driver). */
int gpupcg_device_count(void);

/* Library / device report: fills a short text (device name, SM count, cc) into buf. */
int gpupcg_device_report(int device, char* buf, int bufLen);

#ifdef __cplusplus
}
#endif
#endif /* GPUPCG_H */
