// assembly_scalar.cuh -- the OTHER symmetric scalar systems that reach the lduMatrix::solver boundary,
// assembled on the device with the reference's face / cell loops:
//   TEqn  solvers/thermalStressFoam/TEqn.H:9-15              rhoC*fvm::ddt(T) == fvm::laplacian(k, T, "laplacian(k,T)")
//   pEqn  flowModels/consistentIcoFlow/consistentIcoFlow.C:211-224   fvm::laplacian(1/interpolate(AU), p) == fvc::div(phi); setReference
// [EXT] gaussLaplacianScheme::fvmLaplacianUncorrected (upper = deltaCoeffs*gammaMagSf, negSumDiag, patch coefficients from the
// patch fields' gradientInternalCoeffs / gradientBoundaryCoeffs), EulerDdtScheme::fvmDdt, surfaceIntegrate, fvMatrix ==, setReference.
// The explicit correction of `skewCorrected` for a scalar (numerics/skewCorrectedSnGrad/skewCorrectedSnGrad.C:57-294) with the
// least-squares gradient [EXT leastSquaresGrad] is added when the caller gives psi and the least-squares vectors
// (k_sc_lsgrad, k_sc_corr_faces; the source picks it up in k_sc_cells).
// Operand order of every sum = the reference's loops (per cell: internal faces ascending, then boundary faces), no FMA.
#pragma once

struct ScalarEqnDev
{
    int sign;                   // +1: (rate*ddt(psi)) - laplacian   (TEqn form);  -1: laplacian - div(phi)   (pEqn form)
    int ddt;                    // 1: Euler
    double rDeltaT;
    int gammaOnFaces;           // gamma given per face (pEqn) or per cell, interpolated linearly (TEqn)
};

// faces: gammaMagSf (kept for the patch kernel), upper of the laplacian L (before the sign)
__global__ void __launch_bounds__(ATPB) k_sc_faces(MeshDev M, ScalarEqnDev E, const double* __restrict__ gamma,
                                                   double* __restrict__ gms, double* __restrict__ Lup, double* __restrict__ upper)
{
    for (int f = blockIdx.x*ATPB + threadIdx.x; f < M.nFaces; f += gridDim.x*ATPB)
    {
        const bool internal = f < M.nIF;
        if (!internal && M.bfPatchType[f - M.nIF] == 4) { gms[f] = 0.0; continue; }
        double gf;
        if (E.gammaOnFaces) gf = gamma[f];
        else if (internal)
        {
            const int P = M.owner[f], N = M.neighbour[f];
            gf = add(mul(M.weights[f], sub(gamma[P], gamma[N])), gamma[N]);         // [EXT linear]: w*(P - N) + N
        }
        else gf = gamma[M.owner[f]];                                               // zeroGradient / calculated patch of gamma
        const double g = mul(gf, M.magSf[f]);
        gms[f] = g;
        if (internal)
        {
            const double up = mul(M.deltaCoeffs[f], g);
            Lup[f] = up;
            upper[f] = (E.sign > 0) ? sub(0.0, up) : up;
        }
    }
}

// [EXT] leastSquaresGrad of psi: per cell, its faces in the order of the reference's loops (internal faces ascending -- as owner
// += pVectors*(psi[nei] - psi[own]), as neighbour -= nVectors*(...) -- then its boundary faces: += patch pVectors*(psi_b - psi))
__global__ void __launch_bounds__(ATPB) k_sc_lsgrad(MeshDev M, const double* __restrict__ psi, const int* __restrict__ bfKind,
                                                    const double* __restrict__ bfValue, const double* __restrict__ lsP,
                                                    const double* __restrict__ lsN, const double* __restrict__ lsB,
                                                    double* __restrict__ grad)
{
    for (int c = blockIdx.x*ATPB + threadIdx.x; c < M.nCells; c += gridDim.x*ATPB)
    {
        double g[3] = {0.0, 0.0, 0.0};
        for (int k = M.cellFaceStart[c]; k < M.cellFaceStart[c + 1]; k++)
        {
            const int code = M.cellFace[k], f = code & 0x7fffffff;
            if (f < M.nIF)
            {
                const double d = sub(psi[M.neighbour[f]], psi[M.owner[f]]);
                if (code < 0) { for (int q = 0; q < 3; q++) g[q] = sub(g[q], mul(lsN[3*(size_t)f + q], d)); }
                else { for (int q = 0; q < 3; q++) g[q] = add(g[q], mul(lsP[3*(size_t)f + q], d)); }
            }
            else
            {
                const int bf = f - M.nIF;
                if (M.bfPatchType[bf] == 4) continue;
                const double pb = bfKind[bf] == 1 ? bfValue[bf] : psi[c];
                const double d = sub(pb, psi[c]);
                for (int q = 0; q < 3; q++) g[q] = add(g[q], mul(lsB[3*(size_t)bf + q], d));
            }
        }
        for (int q = 0; q < 3; q++) grad[3*(size_t)c + q] = g[q];
    }
}

// skewCorrectedSnGrad<scalar>::correction per internal face, limited, times gammaMagSf   (skewCorrectedSnGrad.C:137-141, 207-222, 255-288)
__global__ void __launch_bounds__(ATPB) k_sc_corr_faces(MeshDev M, const double* __restrict__ gms, const double* __restrict__ psi,
                                                        const double* __restrict__ grad, double limitCoeff, double* __restrict__ flux)
{
    for (int f = blockIdx.x*ATPB + threadIdx.x; f < M.nIF; f += gridDim.x*ATPB)
    {
        const int P = M.owner[f], N = M.neighbour[f];
        const double Sf[3] = {M.Sf[3*(size_t)f], M.Sf[3*(size_t)f + 1], M.Sf[3*(size_t)f + 2]};
        const double m2 = mul(M.magSf[f], M.magSf[f]);
        double kP[3], kN[3];
        for (int q = 0; q < 3; q++) { kP[q] = sub(M.Cf[3*(size_t)f + q], M.C[3*(size_t)P + q]); kN[q] = sub(M.Cf[3*(size_t)f + q], M.C[3*(size_t)N + q]); }
        const double sP = dot3(Sf, kP);
        for (int q = 0; q < 3; q++) kP[q] = sub(kP[q], dvd(mul(Sf[q], sP), m2));
        const double sN = dot3(Sf, kN);
        for (int q = 0; q < 3; q++) kN[q] = sub(kN[q], dvd(mul(Sf[q], sN), m2));
        const double gN[3] = {grad[3*(size_t)N], grad[3*(size_t)N + 1], grad[3*(size_t)N + 2]};
        const double gP[3] = {grad[3*(size_t)P], grad[3*(size_t)P + 1], grad[3*(size_t)P + 2]};
        const double dc = M.deltaCoeffs[f];
        double corr = mul(sub(dot3(kN, gN), dot3(kP, gP)), dc);
        const double orth = mul(dc, sub(psi[N], psi[P]));
        double lim = dvd(mul(limitCoeff, fabs(add(orth, corr))), add(mul(sub(1.0, limitCoeff), fabs(corr)), A_SMALL));
        if (lim > 1.0) lim = 1.0;
        corr = mul(corr, lim);
        flux[f] = mul(gms[f], corr);
    }
}

// cells: diag (negSumDiag in face order, +- the ddt diagonal), source (ddt term, V*div(phi), the laplacian's explicit correction)
__global__ void __launch_bounds__(ATPB) k_sc_cells(MeshDev M, ScalarEqnDev E, const double* __restrict__ Lup,
                                                   const double* __restrict__ rate, const double* __restrict__ psiOld,
                                                   const double* __restrict__ phi, double* __restrict__ diag, double* __restrict__ source,
                                                   const double* __restrict__ corrFlux = nullptr)
{
    for (int c = blockIdx.x*ATPB + threadIdx.x; c < M.nCells; c += gridDim.x*ATPB)
    {
        double Ld = 0.0, div = 0.0, divc = 0.0;
        for (int k = M.cellFaceStart[c]; k < M.cellFaceStart[c + 1]; k++)
        {
            const int code = M.cellFace[k], f = code & 0x7fffffff;
            const bool nbr = code < 0;
            if (f < M.nIF) Ld = sub(Ld, Lup[f]);
            if (corrFlux && f < M.nIF) divc = nbr ? sub(divc, corrFlux[f]) : add(divc, corrFlux[f]);
            if (phi && !(f >= M.nIF && M.bfPatchType[f - M.nIF] == 4)) div = nbr ? sub(div, phi[f]) : add(div, phi[f]);
        }
        const double V = M.V[c];
        double dA = 0.0, sA = 0.0;
        if (E.ddt)
        {
            dA = mul(mul(E.rDeltaT, V), rate[c]);                       // (rDeltaT*V) then *= rhoC
            sA = mul(mul(mul(E.rDeltaT, psiOld[c]), V), rate[c]);       // (rDeltaT*psi.oldTime()*V) then *= rhoC
        }
        // source of the laplacian matrix itself: 0, or [EXT gaussLaplacianScheme] -= V*div(gammaMagSf*correction)
        const double Lsrc = corrFlux ? sub(0.0, mul(V, dvd(divc, V))) : 0.0;
        if (E.sign > 0)
        {
            diag[c] = sub(dA, Ld);                                      // A - L
            source[c] = sub(sA, Lsrc);
        }
        else
        {
            diag[c] = Ld;
            source[c] = phi ? add(Lsrc, mul(V, dvd(div, V))) : Lsrc;    // L == div(phi): source += V*div(phi)
        }
    }
}

// patches: fixedValue (kind 1): gradientInternalCoeffs = -deltaCoeffs, gradientBoundaryCoeffs = deltaCoeffs*value;
// zeroGradient (0): none; empty (bfPatchType 4): none.  L: internalCoeffs = gms*gIC, boundaryCoeffs = -gms*gBC.
__global__ void __launch_bounds__(ATPB) k_sc_patches(MeshDev M, ScalarEqnDev E, const int* __restrict__ bfKind,
                                                     const double* __restrict__ value, const double* __restrict__ gms,
                                                     double* __restrict__ ic, double* __restrict__ bc)
{
    const int nBF = M.nFaces - M.nIF;
    for (int bf = blockIdx.x*ATPB + threadIdx.x; bf < nBF; bf += gridDim.x*ATPB)
    {
        const int f = M.nIF + bf;
        double Lic = 0.0, Lbc = 0.0;
        if (bfKind[bf] == 1 && M.bfPatchType[bf] != 4)
        {
            const double dcf = M.deltaCoeffs[f], g = gms[f];
            Lic = mul(g, -dcf);
            Lbc = mul(-g, mul(dcf, value[bf]));
        }
        ic[bf] = (E.sign > 0) ? sub(0.0, Lic) : Lic;
        bc[bf] = (E.sign > 0) ? sub(0.0, Lbc) : Lbc;
    }
}

// fvMatrix::setReference(celli, value): source[celli] += diag[celli]*value; diag[celli] += diag[celli]
__global__ void k_sc_reference(int cell, double value, double* diag, double* source)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
    {
        source[cell] = add(source[cell], mul(diag[cell], value));
        diag[cell] = add(diag[cell], diag[cell]);
    }
}
