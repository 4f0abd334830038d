/*
 * oracle/ldu_oracle.c  --  CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement, loop order and operand order included, of the linear
 * solver half of the displacement-equation hot path:
 *   lduMatrix::Amul / sumA, lduMatrix::solver::normFactor / stop,
 *   solverPerformance::checkConvergence / checkSingularity,
 *   DICPreconditioner::{calcReciprocalD, precondition}, PCG::solve,
 *   gSumProd / gSumMag / gAverage / gSum, processor-interface update.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libgpupcg_cuda.so)
 * never links or calls it.
 *
 * PARITY AGAINST foam-extend's BITS UNPINNED.  (What pins the restated path as an algorithm, from outside: the reference-held
 * Kirsch solution on config 1, Lame's cylinder on the shipped tube mesh, free thermal expansion -- tests/test_plate_hole.py,
 * tests/test_shipped_cases.py, tests/test_thermal_stress.py, oracle/ASSUMPTIONS.md.)  None of this arithmetic is under /root/reference: it lives
 * in foam-extend-3.1 (libfoam), which is neither vendored nor installed
 * (SURVEY.md section 8c).  The reference only *calls* it:
 *   src/fluidStructureInteraction/stressModels/unsTotalLagrangianStress/
 *       unsTotalLagrangianStress.C:908                 DEqn.solve()
 *   .../unsTotalLagrangianStressSolve.C:132,378,560    DEqn.solve()/residual()
 *   src/solvers/thermalStressFoam/TEqn.H:15            TEqn.solve()
 *   src/fluidStructureInteraction/flowModels/consistentIcoFlow/
 *       consistentIcoFlow.C:224                        pEqn.solve()
 * and selects it through run/.../system/fvSolution ("solver PCG;
 * preconditioner DIC;", e.g. run/stressFoam/plateHole/plateHole/system/
 * fvSolution:18-27).  The reference ships no golden residual logs, so this
 * file follows the published foam-extend-3.1 algorithm (upstream paths quoted
 * at each function) and every assumption that could differ from the real 3.1
 * source is collected in oracle/ASSUMPTIONS.md.
 *
 * Build: gcc -O3 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off because the stock foam-extend linux64GccDPOpt build is
 * plain x86-64 -O3: no FMA contraction.  No -ffast-math: sums are sequential.
 *
 * Multi-rank: a decomposed case is passed as an array of sub-domains that are
 * advanced in lock step inside one process; the MPI reduce of foam-extend's
 * Pstream (linear, rank-ordered for < 16 ranks) becomes v0 + v1 + ... in rank
 * order, and the processor-patch swap becomes an in-memory gather from the
 * neighbour sub-domain.  With -fopenmp the rank loops run one sub-domain per
 * thread (the "mpirun -np k" stand-in of BASELINE.md section 3); results do not
 * depend on the thread count because every rank-local loop stays sequential.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_SMALL   1.0e-15   /* Foam::SMALL  (doubleScalar) */
#define ORC_VSMALL  1.0e-300  /* Foam::VSMALL */
#define ORC_MSMALL  1.0e-20   /* lduMatrix::small_ */
#define ORC_MGREAT  1.0e+20   /* lduMatrix::great_ */

typedef struct orc_domain
{
    int nCells;
    int nFaces;
    const int* lowerAddr;           /* l[f] (owner), l < u, faces sorted by (l,u) */
    const int* upperAddr;           /* u[f] (neighbour) */
    const double* diag;             /* [nCells] */
    const double* upper;            /* [nFaces]; symmetric matrix: lower aliases upper */
    int nPatches;                   /* coupled (processor) interfaces only */
    const int* patchStart;          /* [nPatches+1] offsets into the two arrays below */
    const int* patchFaceCells;      /* lduAddr().patchAddr(p) concatenated */
    const double* patchBouCoeffs;   /* coupleBouCoeffs[p] concatenated */
    const int* patchNbrDomain;      /* [nPatches] neighbour sub-domain (rank) */
    const int* patchNbrPatch;       /* [nPatches] index of the matching patch there */
    double* x;                      /* [nCells] in/out */
    const double* b;                /* [nCells] */
} orc_domain;

typedef struct orc_perf
{
    double initialResidual;
    double finalResidual;
    double normFactor;
    int nIterations;
    int converged;
    int singular;
} orc_perf;

/* ------------------------------------------------------------------------- */
/* lduMatrix::Amul, rank-local part                                          */
/* [EXT] src/foam/matrices/lduMatrix/lduMatrix/lduMatrixATmul.C              */
/* ------------------------------------------------------------------------- */
void orc_amul_local
(
    int nCells, int nFaces, const int* l, const int* u,
    const double* diag, const double* upper, const double* lower,
    const double* x, double* Ax
)
{
    for (int c = 0; c < nCells; c++)
    {
        Ax[c] = diag[c]*x[c];
    }
    for (int f = 0; f < nFaces; f++)
    {
        Ax[u[f]] += lower[f]*x[l[f]];
        Ax[l[f]] += upper[f]*x[u[f]];
    }
}

/* processorFvPatchField::updateInterfaceMatrix applied for every coupled patch
 * in patch order: result[faceCells[i]] -= coeffs[i]*xNbr[i]
 * [EXT] lduMatrixUpdateMatrixInterfaces.C, processorFvPatchField.C           */
static void orc_update_interfaces
(
    const orc_domain* doms, int d, double* const* xs, double* Ax
)
{
    const orc_domain* D = &doms[d];
    for (int p = 0; p < D->nPatches; p++)
    {
        const orc_domain* N = &doms[D->patchNbrDomain[p]];
        const int np = D->patchNbrPatch[p];
        const int* fc = D->patchFaceCells + D->patchStart[p];
        const double* bc = D->patchBouCoeffs + D->patchStart[p];
        const int* fcN = N->patchFaceCells + N->patchStart[np];
        const double* xN = xs[D->patchNbrDomain[p]];
        const int n = D->patchStart[p + 1] - D->patchStart[p];
        for (int i = 0; i < n; i++)
        {
            Ax[fc[i]] -= bc[i]*xN[fcN[i]];
        }
    }
}

/* Amul over all sub-domains, xs[d] -> Axs[d]. */
static void orc_amul_all
(
    int nD, const orc_domain* doms, double* const* xs, double* const* Axs
)
{
    #pragma omp parallel for schedule(static, 1)
    for (int d = 0; d < nD; d++)
    {
        const orc_domain* D = &doms[d];
        orc_amul_local
        (
            D->nCells, D->nFaces, D->lowerAddr, D->upperAddr,
            D->diag, D->upper, D->upper, xs[d], Axs[d]
        );
    }
    /* the swap: every rank has posted x[faceCells] before anyone updates */
    #pragma omp parallel for schedule(static, 1)
    for (int d = 0; d < nD; d++)
    {
        orc_update_interfaces(doms, d, xs, Axs[d]);
    }
}

/* Single-domain Amul with explicit neighbour values (test entry point):
 * xNbr is the concatenated received patch field.                            */
void orc_amul
(
    int nCells, int nFaces, const int* l, const int* u,
    const double* diag, const double* upper,
    const double* x, double* Ax,
    int nPatches, const int* patchStart, const int* patchFaceCells,
    const double* patchBouCoeffs, const double* xNbr
)
{
    orc_amul_local(nCells, nFaces, l, u, diag, upper, upper, x, Ax);
    for (int p = 0; p < nPatches; p++)
    {
        for (int i = patchStart[p]; i < patchStart[p + 1]; i++)
        {
            Ax[patchFaceCells[i]] -= patchBouCoeffs[i]*xNbr[i];
        }
    }
}

/* lduMatrix::sumA  [EXT] lduMatrixOperations.C */
void orc_sumA
(
    int nCells, int nFaces, const int* l, const int* u,
    const double* diag, const double* upper, const double* lower,
    int nPatches, const int* patchStart, const int* patchFaceCells,
    const double* patchBouCoeffs, double* sumA
)
{
    for (int c = 0; c < nCells; c++)
    {
        sumA[c] = diag[c];
    }
    for (int f = 0; f < nFaces; f++)
    {
        sumA[u[f]] += lower[f];
        sumA[l[f]] += upper[f];
    }
    for (int p = 0; p < nPatches; p++)
    {
        for (int i = patchStart[p]; i < patchStart[p + 1]; i++)
        {
            sumA[patchFaceCells[i]] -= patchBouCoeffs[i];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* DICPreconditioner  [EXT] src/foam/matrices/lduMatrix/preconditioners/      */
/*                          DICPreconditioner/DICPreconditioner.C             */
/* Interfaces are ignored: block-local on every sub-domain.                   */
/* ------------------------------------------------------------------------- */
void orc_dic_calc_reciprocal_d
(
    int nCells, int nFaces, const int* l, const int* u,
    const double* diag, const double* upper, double* rD
)
{
    for (int c = 0; c < nCells; c++)
    {
        rD[c] = diag[c];
    }
    for (int f = 0; f < nFaces; f++)
    {
        rD[u[f]] -= upper[f]*upper[f]/rD[l[f]];
    }
    for (int c = 0; c < nCells; c++)
    {
        rD[c] = 1.0/rD[c];
    }
}

void orc_dic_precondition
(
    int nCells, int nFaces, const int* l, const int* u,
    const double* upper, const double* rD, const double* rA, double* wA
)
{
    for (int c = 0; c < nCells; c++)
    {
        wA[c] = rD[c]*rA[c];
    }
    for (int f = 0; f < nFaces; f++)
    {
        wA[u[f]] -= rD[u[f]]*upper[f]*wA[l[f]];
    }
    for (int f = nFaces - 1; f >= 0; f--)
    {
        wA[l[f]] -= rD[l[f]]*upper[f]*wA[u[f]];
    }
}

/* ------------------------------------------------------------------------- */
/* Field reductions  [EXT] src/foam/fields/Fields/Field/FieldFunctions.C      */
/* local sequential sum, then reduce(sumOp) = rank-ordered linear sum         */
/* ------------------------------------------------------------------------- */
/* Summation order.  0 (default) = the reference's sequential loops.  1 = pairwise (tree) summation of the same
 * terms: NOT the reference order -- used by the tests only to measure how far the residual history of a given case
 * moves when nothing but the association of the reductions changes (the GPU necessarily re-associates them).    */
static int orc_sum_mode = 0;
void orc_set_sum_mode(int m) { orc_sum_mode = m; }

static double orc_pairwise(int n, const double* a, const double* b, int kind)
{
    if (n <= 64)
    {
        double s = 0.0;
        if (kind == 0) for (int i = 0; i < n; i++) s += a[i]*b[i];
        else if (kind == 1) for (int i = 0; i < n; i++) s += fabs(a[i]);
        else for (int i = 0; i < n; i++) s += a[i];
        return s;
    }
    const int h = n/2;
    return orc_pairwise(h, a, b, kind) + orc_pairwise(n - h, a + h, b ? b + h : b, kind);
}

static double orc_sumProd(int n, const double* a, const double* b)
{
    if (orc_sum_mode) return orc_pairwise(n, a, b, 0);
    double s = 0.0;
    for (int i = 0; i < n; i++) s += a[i]*b[i];
    return s;
}

static double orc_sumMag(int n, const double* a)
{
    if (orc_sum_mode) return orc_pairwise(n, a, 0, 1);
    double s = 0.0;
    for (int i = 0; i < n; i++) s += fabs(a[i]);
    return s;
}

static double orc_sum(int n, const double* a)
{
    if (orc_sum_mode) return orc_pairwise(n, a, 0, 2);
    double s = 0.0;
    for (int i = 0; i < n; i++) s += a[i];
    return s;
}

static double orc_reduce(int nD, const double* local)
{
    double s = local[0];
    for (int d = 1; d < nD; d++) s += local[d];
    return s;
}

static double orc_gSumProd
(
    int nD, const orc_domain* doms, double* const* a, double* const* b,
    double* scratch
)
{
    #pragma omp parallel for schedule(static, 1)
    for (int d = 0; d < nD; d++)
    {
        scratch[d] = orc_sumProd(doms[d].nCells, a[d], b[d]);
    }
    return orc_reduce(nD, scratch);
}

static double orc_gSumMag
(
    int nD, const orc_domain* doms, double* const* a, double* scratch
)
{
    #pragma omp parallel for schedule(static, 1)
    for (int d = 0; d < nD; d++)
    {
        scratch[d] = orc_sumMag(doms[d].nCells, a[d]);
    }
    return orc_reduce(nD, scratch);
}

/* solverPerformance::checkConvergence  [EXT] lduMatrixSolverPerformance.C
 * (ASSUMPTIONS.md #2: "<" on Tolerance, "<=" on the relative test, RelTol
 * compared against SMALL).                                                   */
static int orc_checkConvergence
(
    orc_perf* perf, double tol, double relTol
)
{
    if
    (
        perf->finalResidual < tol
     || (relTol > ORC_SMALL && perf->finalResidual <= relTol*perf->initialResidual)
    )
    {
        perf->converged = 1;
    }
    else
    {
        perf->converged = 0;
    }
    return perf->converged;
}

/* lduMatrix::solver::stop  [EXT] lduMatrixSolver.C / lduMatrix.H */
static int orc_stop
(
    orc_perf* perf, double tol, double relTol, int minIter, int maxIter
)
{
    if (perf->nIterations < minIter)
    {
        return 0;
    }
    if (perf->nIterations >= maxIter || orc_checkConvergence(perf, tol, relTol))
    {
        return 1;
    }
    return 0;
}

/* solverPerformance::checkSingularity */
static int orc_checkSingularity(orc_perf* perf, double residual)
{
    perf->singular = (residual > ORC_VSMALL) ? 0 : 1;
    return perf->singular;
}

/* ------------------------------------------------------------------------- */
/* PCG::solve  [EXT] src/foam/matrices/lduMatrix/solvers/PCG/PCG.C            */
/* with lduMatrix::solver::normFactor [EXT] lduMatrixSolver.C                 */
/* history[k] = finalResidual after iteration k+1 (history[-1] would be the   */
/* initial residual, which is returned in perf).                              */
/* ------------------------------------------------------------------------- */
int orc_pcg_solve
(
    int nD, orc_domain* doms,
    double tol, double relTol, int minIter, int maxIter,
    orc_perf* perf, double* history, int historyCap
)
{
    if (nD < 1) return -1;

    double** xs = (double**)malloc(sizeof(double*)*nD);
    double** pA = (double**)malloc(sizeof(double*)*nD);
    double** wA = (double**)malloc(sizeof(double*)*nD);
    double** rA = (double**)malloc(sizeof(double*)*nD);
    double** rD = (double**)malloc(sizeof(double*)*nD);
    double* scratch = (double*)malloc(sizeof(double)*nD);
    long nGlobal = 0;
    for (int d = 0; d < nD; d++)
    {
        const int n = doms[d].nCells;
        nGlobal += n;
        xs[d] = doms[d].x;
        pA[d] = (double*)malloc(sizeof(double)*(n > 0 ? n : 1));
        wA[d] = (double*)malloc(sizeof(double)*(n > 0 ? n : 1));
        rA[d] = (double*)malloc(sizeof(double)*(n > 0 ? n : 1));
        rD[d] = (double*)malloc(sizeof(double)*(n > 0 ? n : 1));
    }

    memset(perf, 0, sizeof(*perf));

    double wArA = ORC_MGREAT;
    double wArAold = wArA;

    /* --- Calculate A.x */
    orc_amul_all(nD, doms, xs, wA);

    /* --- Calculate initial residual field */
    #pragma omp parallel for schedule(static, 1)
    for (int d = 0; d < nD; d++)
    {
        const int n = doms[d].nCells;
        const double* b = doms[d].b;
        for (int i = 0; i < n; i++) rA[d][i] = b[i] - wA[d][i];
    }

    /* --- normFactor(x, b, wA, pA): foam-extend form (ASSUMPTIONS.md #1):
     *     xRef = gAverage(x); tmp = A*(xRef field) by a full Amul;
     *     gSum(mag(Ax - tmp) + mag(b - tmp)) + small_                       */
    double normFactor;
    {
        for (int d = 0; d < nD; d++)
        {
            scratch[d] = orc_sum(doms[d].nCells, xs[d]);
        }
        const double xRef = orc_reduce(nD, scratch)/(double)nGlobal;

        /* rD is free at this point: use it for the constant xRef field */
        for (int d = 0; d < nD; d++)
        {
            for (int i = 0; i < doms[d].nCells; i++) rD[d][i] = xRef;
        }
        orc_amul_all(nD, doms, rD, pA);

        #pragma omp parallel for schedule(static, 1)
        for (int d = 0; d < nD; d++)
        {
            const int n = doms[d].nCells;
            const double* b = doms[d].b;
            double s = 0.0;
            for (int i = 0; i < n; i++)
            {
                s += fabs(wA[d][i] - pA[d][i]) + fabs(b[i] - pA[d][i]);
            }
            scratch[d] = s;
        }
        normFactor = orc_reduce(nD, scratch) + ORC_MSMALL;
    }
    perf->normFactor = normFactor;

    /* --- Calculate normalised residual norm */
    perf->initialResidual = orc_gSumMag(nD, doms, rA, scratch)/normFactor;
    perf->finalResidual = perf->initialResidual;

    /* --- Check convergence, solve if not converged */
    if (!orc_stop(perf, tol, relTol, minIter, maxIter))
    {
        /* preconditioner::New -> DICPreconditioner ctor: calcReciprocalD */
        #pragma omp parallel for schedule(static, 1)
        for (int d = 0; d < nD; d++)
        {
            const orc_domain* D = &doms[d];
            orc_dic_calc_reciprocal_d
            (
                D->nCells, D->nFaces, D->lowerAddr, D->upperAddr,
                D->diag, D->upper, rD[d]
            );
        }

        do
        {
            wArAold = wArA;

            /* --- Precondition residual */
            #pragma omp parallel for schedule(static, 1)
            for (int d = 0; d < nD; d++)
            {
                const orc_domain* D = &doms[d];
                orc_dic_precondition
                (
                    D->nCells, D->nFaces, D->lowerAddr, D->upperAddr,
                    D->upper, rD[d], rA[d], wA[d]
                );
            }

            /* --- Update search directions */
            wArA = orc_gSumProd(nD, doms, wA, rA, scratch);

            if (perf->nIterations == 0)
            {
                #pragma omp parallel for schedule(static, 1)
                for (int d = 0; d < nD; d++)
                {
                    for (int i = 0; i < doms[d].nCells; i++) pA[d][i] = wA[d][i];
                }
            }
            else
            {
                const double beta = wArA/wArAold;
                #pragma omp parallel for schedule(static, 1)
                for (int d = 0; d < nD; d++)
                {
                    for (int i = 0; i < doms[d].nCells; i++)
                    {
                        pA[d][i] = wA[d][i] + beta*pA[d][i];
                    }
                }
            }

            /* --- Update preconditioned residual */
            orc_amul_all(nD, doms, pA, wA);

            const double wApA = orc_gSumProd(nD, doms, wA, pA, scratch);

            /* --- Test for singularity */
            if (orc_checkSingularity(perf, fabs(wApA)/normFactor)) break;

            /* --- Update solution and residual */
            const double alpha = wArA/wApA;
            #pragma omp parallel for schedule(static, 1)
            for (int d = 0; d < nD; d++)
            {
                double* x = xs[d];
                for (int i = 0; i < doms[d].nCells; i++)
                {
                    x[i] += alpha*pA[d][i];
                    rA[d][i] -= alpha*wA[d][i];
                }
            }

            perf->finalResidual = orc_gSumMag(nD, doms, rA, scratch)/normFactor;
            if (history && perf->nIterations < historyCap)
            {
                history[perf->nIterations] = perf->finalResidual;
            }
            perf->nIterations++;
        }
        while (!orc_stop(perf, tol, relTol, minIter, maxIter));
    }

    for (int d = 0; d < nD; d++)
    {
        free(pA[d]); free(wA[d]); free(rA[d]); free(rD[d]);
    }
    free(xs); free(pA); free(wA); free(rA); free(rD); free(scratch);
    return 0;
}

/* One PCG iteration's worth of kernels on fixed data, used by bench.py's
 * cpu_baseline to time Amul / DIC in isolation.  Returns a checksum so the
 * work cannot be optimised away.                                             */
double orc_bench_amul
(
    int reps, int nCells, int nFaces, const int* l, const int* u,
    const double* diag, const double* upper, const double* x, double* Ax
)
{
    double s = 0.0;
    for (int r = 0; r < reps; r++)
    {
        orc_amul_local(nCells, nFaces, l, u, diag, upper, upper, x, Ax);
        s += Ax[(r*7919) % nCells];
    }
    return s;
}

double orc_bench_dic
(
    int reps, int nCells, int nFaces, const int* l, const int* u,
    const double* upper, const double* rD, const double* rA, double* wA
)
{
    double s = 0.0;
    for (int r = 0; r < reps; r++)
    {
        orc_dic_precondition(nCells, nFaces, l, u, upper, rD, rA, wA);
        s += wA[(r*7919) % nCells];
    }
    return s;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
