// assembly_ext.cuh -- included by assembly.cu inside its anonymous namespace, after the kernels of the linear steady
// D equation.  The rest of unsTotalLagrangianStress::evolve() (unsTotalLagrangianStress.C:769-1009) on the device:
//   * Euler d2dt2(D) [EXT EulerD2dt2Scheme, same weights as the reference's backwardD2dt2Scheme.C:306-331], the damping
//     term K*rho*fvm::ddt(D) (:861-865), the non-linear fluxes (:867-881) and the thermal flux (:883-890);
//   * tractionDisplacement's non-linear / thermal gradient (tractionDisplacementFvPatchVectorField.C:199-251);
//   * D.correctBoundaryConditions(): evaluate() of the patch fields (tractionDisplacement...C:259-300,
//     symmetryDisplacement...C:169-207, [EXT] symmetryFvPatchField) -> the boundary values the point interpolation reads;
//   * leastSquaresVolPointInterpolation::interpolate (leastSquaresVolPointInterpolate.C:89-299);
//   * det(I + gradDf) bounds (:929-946), U = fvc::ddt(D) (:973), epsilon / sigma with the non-linear and thermal terms
//     (:975-990).
// Same rule as everywhere in this library: every product / sum is a __d*_rn intrinsic in the operand order of the
// reference's field expressions (one elementwise operator after the other, left to right), no FMA.
#pragma once

struct AsmCtl
{
    int d2dt2;          // 0 steadyState, 1 Euler
    int ddt;            // 0 steadyState, 1 Euler
    int damping;        // K > SMALL
    int nonLinear;      // nonLinear && !enforceLinear
    int thermal;
    double coefft, coefft00, rDeltaT2, rDeltaT, K;
    double cbD, cb0D, cb00D;    // backward fvm::ddt(D) of the damping term: coefft, coefft0, coefft00 (effective deltaT0 = eDamp)
};

// weights of one use of a backward scheme [EXT backwardDdtScheme / backwardD2dt2Scheme.C:251-253]
struct BwW { double c, c0, c00; };
struct BackwardCtl { double rDt, sc; BwW d2, ddtD, ddtD0, ddtD00; };

// rho*fvm::d2dt2(D), backward (backwardD2dt2Scheme.C:230-280 + [EXT] backwardDdtScheme::fvmDdt / fvcDdt): diag and source
// of the term per cell -> timeA[4*c] = {diag, source x, y, z}; k_asm_cells picks them up.  Operand order as oracle/fvm_oracle.c.
__global__ void __launch_bounds__(ATPB) k_time_backward(int nCells, BackwardCtl W, const double* __restrict__ V,
                                                        const double* __restrict__ rho, const double* __restrict__ D0,
                                                        const double* __restrict__ D00, const double* __restrict__ D000,
                                                        const double* __restrict__ D0000, double* __restrict__ timeA)
{
    for (int c = blockIdx.x*ATPB + threadIdx.x; c < nCells; c += gridDim.x*ATPB)
    {
        const double v = V[c], rh = rho[c];
        timeA[4*(size_t)c] = mul(mul(mul(mul(W.ddtD.c, W.rDt), v), W.sc), rh);
        const double rV = mul(W.rDt, v);
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            const double a0 = D0[3*(size_t)c + d], a00 = D00[3*(size_t)c + d], a000 = D000[3*(size_t)c + d], a0000 = D0000[3*(size_t)c + d];
            const double s1 = mul(mul(rV, sub(mul(W.ddtD.c0, a0), mul(W.ddtD.c00, a00))), W.sc);
            const double X = mul(W.rDt, add(sub(mul(W.ddtD0.c, a0), mul(W.ddtD0.c0, a00)), mul(W.ddtD0.c00, a000)));
            const double Y = mul(W.rDt, add(sub(mul(W.ddtD00.c, a00), mul(W.ddtD00.c0, a000)), mul(W.ddtD00.c00, a0000)));
            const double s2 = add(s1, mul(rV, sub(mul(W.d2.c0, X), mul(W.d2.c00, Y))));
            timeA[4*(size_t)c + 1 + d] = mul(s2, rh);
        }
    }
}

// U = fvc::ddt(D), backward [EXT backwardDdtScheme::fvcDdt]
__global__ void __launch_bounds__(ATPB) k_ddt_backward(long long n3, double rDt, BwW w, const double* __restrict__ D,
                                                       const double* __restrict__ D0, const double* __restrict__ D00,
                                                       double* __restrict__ U)
{
    for (long long i = (long long)blockIdx.x*ATPB + threadIdx.x; i < n3; i += (long long)gridDim.x*ATPB)
        U[i] = mul(rDt, add(sub(mul(w.c, D[i]), mul(w.c0, D0[i])), mul(w.c00, D00[i])));
}

// (A & B)_ij = A_ik B_kj accumulated k = x, y, z
__device__ __forceinline__ void TdotT(const double* A, const double* B, double* R)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            R[3*i + j] = add(add(mul(A[3*i], B[j]), mul(A[3*i + 1], B[3 + j])), mul(A[3*i + 2], B[6 + j]));
}
__device__ __forceinline__ void transposeT(const double* A, double* R)
{
    R[0] = A[0]; R[1] = A[3]; R[2] = A[6]; R[3] = A[1]; R[4] = A[4]; R[5] = A[7]; R[6] = A[2]; R[7] = A[5]; R[8] = A[8];
}
__device__ __forceinline__ void symmT(const double* T, double* S)
{
    S[0] = T[0]; S[1] = mul(0.5, add(T[1], T[3])); S[2] = mul(0.5, add(T[2], T[6]));
    S[3] = T[4]; S[4] = mul(0.5, add(T[5], T[7])); S[5] = T[8];
}
__device__ __forceinline__ void Tdotv(const double* T, const double* v, double* r)
{
    r[0] = add(add(mul(T[0], v[0]), mul(T[1], v[1])), mul(T[2], v[2]));
    r[1] = add(add(mul(T[3], v[0]), mul(T[4], v[1])), mul(T[5], v[2]));
    r[2] = add(add(mul(T[6], v[0]), mul(T[7], v[1])), mul(T[8], v[2]));
}
__device__ __forceinline__ double detT(const double* t)
{
    const double a = mul(mul(t[0], t[4]), t[8]), b = mul(mul(t[1], t[5]), t[6]), c = mul(mul(t[2], t[3]), t[7]);
    const double d = mul(mul(t[0], t[5]), t[7]), e = mul(mul(t[1], t[3]), t[8]), f = mul(mul(t[2], t[4]), t[6]);
    return sub(sub(sub(add(add(a, b), c), d), e), f);
}
__device__ __forceinline__ void invT(const double* t, double* r)
{
    const double d = detT(t);
    r[0] = __ddiv_rn(sub(mul(t[4], t[8]), mul(t[7], t[5])), d); r[1] = __ddiv_rn(sub(mul(t[2], t[7]), mul(t[1], t[8])), d);
    r[2] = __ddiv_rn(sub(mul(t[1], t[5]), mul(t[2], t[4])), d); r[3] = __ddiv_rn(sub(mul(t[6], t[5]), mul(t[3], t[8])), d);
    r[4] = __ddiv_rn(sub(mul(t[0], t[8]), mul(t[2], t[6])), d); r[5] = __ddiv_rn(sub(mul(t[3], t[2]), mul(t[0], t[5])), d);
    r[6] = __ddiv_rn(sub(mul(t[3], t[7]), mul(t[4], t[6])), d); r[7] = __ddiv_rn(sub(mul(t[1], t[6]), mul(t[0], t[7])), d);
    r[8] = __ddiv_rn(sub(mul(t[0], t[4]), mul(t[3], t[1])), d);
}

// the two extra explicit fluxes of the non-linear branch (unsTotalLagrangianStress.C:869-880)
__device__ __forceinline__ void nonlinear_fluxes(const double* Sf, double muf, double laf, const double* G, double* X, double* Y)
{
    double GT[9], GG[9], v[3], sG[6], sGG[6], Ef[6], sg[6], SG[9];
    transposeT(G, GT);
    TdotT(G, GT, GG);
    vdotT(Sf, GG, v);
    const double trGG = add(add(GG[0], GG[4]), GG[8]);
    const double c = mul(mul(0.5, laf), trGG);
#pragma unroll
    for (int d = 0; d < 3; d++) X[d] = add(mul(muf, v[d]), mul(c, Sf[d]));
    symmT(G, sG);
    symmT(GG, sGG);
#pragma unroll
    for (int k = 0; k < 6; k++) Ef[k] = add(sG[k], mul(0.5, sGG[k]));
    const double trE = add(add(Ef[0], Ef[3]), Ef[5]);
    const double tm = mul(2.0, muf), sph = mul(laf, trE);
#pragma unroll
    for (int k = 0; k < 6; k++) sg[k] = mul(tm, Ef[k]);
    sg[0] = add(sg[0], sph); sg[3] = add(sg[3], sph); sg[5] = add(sg[5], sph);
    const double F[9] = {sg[0], sg[1], sg[2], sg[1], sg[3], sg[4], sg[2], sg[4], sg[5]};
    TdotT(F, G, SG);
    vdotT(Sf, SG, Y);
}

// [EXT] symmetryFvPatchField<vector>::snGrad()  (patch type 3, symmetryPlane)
__device__ void symmetryPlane_snGrad(const MeshDev& M, int f, const double* __restrict__ D, double* sn)
{
    const int c = M.owner[f];
    const double magSf = M.magSf[f];
    double n[3], UP[3], S[6], T[6], tv[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { n[d] = dvd(M.Sf[3*(size_t)f + d], magSf); UP[d] = D[3*(size_t)c + d]; }
    S[0] = mul(n[0], n[0]); S[1] = mul(n[0], n[1]); S[2] = mul(n[0], n[2]); S[3] = mul(n[1], n[1]); S[4] = mul(n[1], n[2]); S[5] = mul(n[2], n[2]);
    T[0] = sub(1.0, mul(2.0, S[0])); T[1] = -mul(2.0, S[1]); T[2] = -mul(2.0, S[2]);
    T[3] = sub(1.0, mul(2.0, S[3])); T[4] = -mul(2.0, S[4]); T[5] = sub(1.0, mul(2.0, S[5]));
    tv[0] = add(add(mul(T[0], UP[0]), mul(T[1], UP[1])), mul(T[2], UP[2]));
    tv[1] = add(add(mul(T[1], UP[0]), mul(T[3], UP[1])), mul(T[4], UP[2]));
    tv[2] = add(add(mul(T[2], UP[0]), mul(T[4], UP[1])), mul(T[5], UP[2]));
    const double h = dvd(M.deltaCoeffs[f], 2.0);
#pragma unroll
    for (int d = 0; d < 3; d++) sn[d] = mul(sub(tv[d], UP[d]), h);
}

// tractionDisplacementFvPatchVectorField::updateCoeffs, all branches (:148-257): gradient() of one boundary face
__device__ void traction_gradient(const MeshDev& M, const AsmCtl& T, int f, double m, double la, const double* G,
                                  const double* tr3, double pressure, double threeK, double alpha, double DTb, double* grad)
{
    const double magSf = M.magSf[f];
    double n[3], t[3], Tm[9], nT[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { n[d] = dvd(M.Sf[3*(size_t)f + d], magSf); t[d] = tr3[d]; }
    if (T.nonLinear)
    {
        double F[9], invF[9], nC[3], Jn[3], tv[3];
#pragma unroll
        for (int k = 0; k < 9; k++) F[k] = G[k];
        F[0] = add(1.0, G[0]); F[4] = add(1.0, G[4]); F[8] = add(1.0, G[8]);
        const double J = detT(F);
        invT(F, invF);
        Tdotv(invF, n, nC);
#pragma unroll
        for (int d = 0; d < 3; d++) Jn[d] = mul(J, nC[d]);
        const double SoS0 = mag3(Jn);
        const double mc = mag3(nC);
#pragma unroll
        for (int d = 0; d < 3; d++) nC[d] = __ddiv_rn(nC[d], mc);
#pragma unroll
        for (int d = 0; d < 3; d++) t[d] = sub(t[d], mul(pressure, nC[d]));
        vdotT(t, invF, tv);
#pragma unroll
        for (int d = 0; d < 3; d++) t[d] = mul(tv[d], SoS0);
    }
    else
    {
#pragma unroll
        for (int d = 0; d < 3; d++) t[d] = sub(t[d], mul(pressure, n[d]));
    }
    const double ml = add(m, la);
    Tm[0] = sub(mul(m, G[0]), mul(ml, G[0])); Tm[1] = sub(mul(m, G[3]), mul(ml, G[1])); Tm[2] = sub(mul(m, G[6]), mul(ml, G[2]));
    Tm[3] = sub(mul(m, G[1]), mul(ml, G[3])); Tm[4] = sub(mul(m, G[4]), mul(ml, G[4])); Tm[5] = sub(mul(m, G[7]), mul(ml, G[5]));
    Tm[6] = sub(mul(m, G[2]), mul(ml, G[6])); Tm[7] = sub(mul(m, G[5]), mul(ml, G[7])); Tm[8] = sub(mul(m, G[8]), mul(ml, G[8]));
    vdotT(n, Tm, nT);
    const double tr = add(add(G[0], G[4]), G[8]);
#pragma unroll
    for (int d = 0; d < 3; d++) grad[d] = sub(sub(t[d], nT[d]), mul(mul(n[d], la), tr));
    if (T.nonLinear)
    {
        double GT[9], GG[9], mGG[9], a[3];
        transposeT(G, GT);
        TdotT(G, GT, GG);
#pragma unroll
        for (int k = 0; k < 9; k++) mGG[k] = mul(m, GG[k]);
        vdotT(n, mGG, a);
        const double trGG = add(add(GG[0], GG[4]), GG[8]);
#pragma unroll
        for (int d = 0; d < 3; d++) grad[d] = sub(grad[d], add(a[d], mul(mul(mul(0.5, n[d]), la), trGG)));
    }
    if (T.thermal)
    {
#pragma unroll
        for (int d = 0; d < 3; d++) grad[d] = add(grad[d], mul(mul(mul(n[d], threeK), alpha), DTb));
    }
    const double den = add(mul(2.0, m), la);
#pragma unroll
    for (int d = 0; d < 3; d++) grad[d] = dvd(grad[d], den);
}

// ---- extra explicit fluxes: one thread per face -> fluxX, fluxY (non-linear), fluxT (thermal), [nFaces*3] each -------
__global__ void __launch_bounds__(ATPB) k_asm_faces_extra(MeshDev M, AsmCtl T, const double* __restrict__ mu, const double* __restrict__ lambda,
                                                          const double* __restrict__ gradDf, const double* __restrict__ threeK,
                                                          const double* __restrict__ alpha, const double* __restrict__ DT,
                                                          const double* __restrict__ DTb, double* __restrict__ fluxX,
                                                          double* __restrict__ fluxY, double* __restrict__ fluxT)
{
    for (int f = blockIdx.x*ATPB + threadIdx.x; f < M.nFaces; f += gridDim.x*ATPB)
    {
        const int P = M.owner[f];
        const bool internal = f < M.nIF;
        const int N = internal ? M.neighbour[f] : P;
        const bool empty = !internal && M.bfPatchType[f - M.nIF] == 4;
        const double w = internal ? M.weights[f] : 1.0;
        const double Sf[3] = {M.Sf[3*(size_t)f], M.Sf[3*(size_t)f + 1], M.Sf[3*(size_t)f + 2]};
        if (T.nonLinear)
        {
            double X[3] = {0.0, 0.0, 0.0}, Y[3] = {0.0, 0.0, 0.0};
            if (!empty)
            {
                double muf, laf, G[9];
                if (internal)
                {
                    muf = add(mul(w, sub(mu[P], mu[N])), mu[N]);
                    laf = add(mul(w, sub(lambda[P], lambda[N])), lambda[N]);
                }
                else { muf = mu[P]; laf = lambda[P]; }
#pragma unroll
                for (int k = 0; k < 9; k++) G[k] = gradDf[9*(size_t)f + k];
                nonlinear_fluxes(Sf, muf, laf, G, X, Y);
            }
#pragma unroll
            for (int d = 0; d < 3; d++) { fluxX[3*(size_t)f + d] = X[d]; fluxY[3*(size_t)f + d] = Y[d]; }
        }
        if (T.thermal)
        {
            double tk = 0.0, al = 0.0, dt = 0.0;
            if (internal)
            {
                tk = add(mul(w, sub(threeK[P], threeK[N])), threeK[N]);
                al = add(mul(w, sub(alpha[P], alpha[N])), alpha[N]);
                dt = add(mul(w, sub(DT[P], DT[N])), DT[N]);
            }
            else if (!empty) { tk = threeK[P]; al = alpha[P]; dt = DTb[f - M.nIF]; }
#pragma unroll
            for (int d = 0; d < 3; d++) fluxT[3*(size_t)f + d] = empty ? 0.0 : mul(mul(mul(Sf[d], tk), al), dt);
        }
    }
}

// ... and their surfaceIntegrate into the source, after k_asm_cells, in the order evolve() applies them:
//   DEqn -= fvc::div(X) + fvc::div(Y)  ->  source += V*(divX + divY);   DEqn += fvc::div(T)  ->  source -= V*divT
__global__ void __launch_bounds__(ATPB) k_asm_cells_extra(MeshDev M, AsmCtl T, const double* __restrict__ fluxX,
                                                          const double* __restrict__ fluxY, const double* __restrict__ fluxT,
                                                          double* __restrict__ source)
{
    for (int c = blockIdx.x*ATPB + threadIdx.x; c < M.nCells; c += gridDim.x*ATPB)
    {
        double sX[3] = {0.0, 0.0, 0.0}, sY[3] = {0.0, 0.0, 0.0}, sT[3] = {0.0, 0.0, 0.0};
        for (int k = M.cellFaceStart[c]; k < M.cellFaceStart[c + 1]; k++)
        {
            const int code = M.cellFace[k];
            const size_t f = (size_t)(code & 0x7fffffff);
            const bool nbr = code < 0;
#pragma unroll
            for (int d = 0; d < 3; d++)
            {
                if (T.nonLinear)
                {
                    const double x = fluxX[3*f + d], y = fluxY[3*f + d];
                    sX[d] = nbr ? sub(sX[d], x) : add(sX[d], x);
                    sY[d] = nbr ? sub(sY[d], y) : add(sY[d], y);
                }
                if (T.thermal)
                {
                    const double t = fluxT[3*f + d];
                    sT[d] = nbr ? sub(sT[d], t) : add(sT[d], t);
                }
            }
        }
        const double V = M.V[c];
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            double s = source[3*(size_t)c + d];
            if (T.nonLinear) s = add(s, mul(V, add(__ddiv_rn(sX[d], V), __ddiv_rn(sY[d], V))));
            if (T.thermal) s = sub(s, mul(V, __ddiv_rn(sT[d], V)));
            source[3*(size_t)c + d] = s;
        }
    }
}

// ---- D.correctBoundaryConditions(): boundary values of D, one thread per boundary face ----------------------------
__global__ void __launch_bounds__(ATPB) k_patch_evaluate(MeshDev M, const double* __restrict__ D, const double* __restrict__ gradD,
                                                         const double* __restrict__ fixedValue, const double* __restrict__ patchGradient,
                                                         double* __restrict__ Db)
{
    const int nBF = M.nFaces - M.nIF;
    for (int bf = blockIdx.x*ATPB + threadIdx.x; bf < nBF; bf += gridDim.x*ATPB)
    {
        const int pt = M.bfPatchType[bf];
        if (pt == 4) continue;
        if (pt == 1 || pt == 5)
        {
#pragma unroll
            for (int d = 0; d < 3; d++) Db[3*(size_t)bf + d] = fixedValue[3*(size_t)bf + d];
            continue;
        }
        const int f = M.nIF + bf, c = M.owner[f];
        const double magSf = M.magSf[f];
        double n[3], delta[3], k[3], kg[3], G[9];
#pragma unroll
        for (int d = 0; d < 3; d++) { n[d] = dvd(M.Sf[3*(size_t)f + d], magSf); delta[d] = sub(M.Cf[3*(size_t)f + d], M.C[3*(size_t)c + d]); }
        const double nd = dot3(n, delta);
#pragma unroll
        for (int d = 0; d < 3; d++) k[d] = sub(delta[d], mul(n[d], nd));
#pragma unroll
        for (int q = 0; q < 9; q++) G[q] = gradD[9*(size_t)c + q];
        if (pt == 0)
        {
            vdotT(k, G, kg);
            const double dc = M.deltaCoeffs[f];
#pragma unroll
            for (int d = 0; d < 3; d++)
                Db[3*(size_t)bf + d] = add(add(D[3*(size_t)c + d], kg[d]), __ddiv_rn(patchGradient[3*(size_t)bf + d], dc));
            continue;
        }
        double UP[3], S[6], T[6], tv[3];
        if (pt == 2)
        {
            vdotT(k, G, kg);
#pragma unroll
            for (int d = 0; d < 3; d++) UP[d] = add(D[3*(size_t)c + d], kg[d]);
        }
        else
        {
#pragma unroll
            for (int d = 0; d < 3; d++) UP[d] = D[3*(size_t)c + d];
        }
        S[0] = mul(n[0], n[0]); S[1] = mul(n[0], n[1]); S[2] = mul(n[0], n[2]); S[3] = mul(n[1], n[1]); S[4] = mul(n[1], n[2]); S[5] = mul(n[2], n[2]);
        T[0] = sub(1.0, mul(2.0, S[0])); T[1] = -mul(2.0, S[1]); T[2] = -mul(2.0, S[2]);
        T[3] = sub(1.0, mul(2.0, S[3])); T[4] = -mul(2.0, S[4]); T[5] = sub(1.0, mul(2.0, S[5]));
        tv[0] = add(add(mul(T[0], UP[0]), mul(T[1], UP[1])), mul(T[2], UP[2]));
        tv[1] = add(add(mul(T[1], UP[0]), mul(T[3], UP[1])), mul(T[4], UP[2]));
        tv[2] = add(add(mul(T[2], UP[0]), mul(T[4], UP[1])), mul(T[5], UP[2]));
#pragma unroll
        for (int d = 0; d < 3; d++) Db[3*(size_t)bf + d] = mul(add(UP[d], tv[d]), 0.5);      // x/2.0 == x*0.5 exactly
    }
}

// GeometricField::relax on the boundary field
__global__ void __launch_bounds__(ATPB) k_relax_boundary(long long n3, double alpha, double* __restrict__ Db, const double* __restrict__ DbPrev)
{
    for (long long i = (long long)blockIdx.x*ATPB + threadIdx.x; i < n3; i += (long long)gridDim.x*ATPB)
    {
        const double pv = DbPrev[i];
        Db[i] = add(pv, mul(alpha, sub(Db[i], pv)));
    }
}

// ---- leastSquaresVolPointInterpolation::interpolate (leastSquaresVolPointInterpolate.C:89-299) ----------------------
// One thread per point.  lsqStart[p]..: the point's stencil rows (cells ascending, then boundary faces in the
// reference's hash-set order; points on an empty patch: the same rows again, mirrored); lsqEntry: cell id >= 0 or
// -1 - (boundary face index); lsqInv[row][3] = column `row` of the point's inverse least-squares matrix; W == 1.
// Two passes over the stencil: avg (all sources summed in stencil order), then coeffs += inv*(source - avg); the second
// pass re-reads the sources out of L1/L2.  Mirrored rows: source = transform(I, source), evaluated literally.
struct LsqDev
{
    int nPoints;
    const int *start, *entry;
    const double *inv, *origin;
    const unsigned char* mirror;
};

// point patch fields of pointD that act in pf.correctBoundaryConditions() (leastSquaresVolPointInterpolate.C:301)
struct PointBcDev
{
    int n;                          // constrained points
    const int *point, *start, *kind;
    const double* vec;              // [rows][3]: nHat (kind 1, symmetryPlane) or the value (kind 2, fixedValue)
};

__device__ __forceinline__ void identity_transform(const double* v, double* r)
{
    r[0] = add(add(mul(1.0, v[0]), mul(0.0, v[1])), mul(0.0, v[2]));
    r[1] = add(add(mul(0.0, v[0]), mul(1.0, v[1])), mul(0.0, v[2]));
    r[2] = add(add(mul(0.0, v[0]), mul(0.0, v[1])), mul(1.0, v[2]));
}

__global__ void __launch_bounds__(ATPB) k_lsq_interpolate(LsqDev L, const double* __restrict__ points, const double* __restrict__ D,
                                                          const double* __restrict__ Db, double* __restrict__ pointD)
{
    for (int p = blockIdx.x*ATPB + threadIdx.x; p < L.nPoints; p += gridDim.x*ATPB)
    {
        const int r0 = L.start[p], m = L.start[p + 1] - r0;
        const bool mir = L.mirror[p] != 0;
        const int n = mir ? m/2 : m;
        double avg[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < n; i++)
        {
            const int e = L.entry[r0 + i];
            const double* v = e >= 0 ? D + 3*(size_t)e : Db + 3*(size_t)(-1 - e);
#pragma unroll
            for (int d = 0; d < 3; d++) avg[d] = add(avg[d], v[d]);          // sqr(W)*v == v
        }
        if (mir)
        {
            double ta[3];
            identity_transform(avg, ta);
#pragma unroll
            for (int d = 0; d < 3; d++) avg[d] = add(avg[d], ta[d]);
        }
        const double sw = (double)m;
#pragma unroll
        for (int d = 0; d < 3; d++) avg[d] = __ddiv_rn(avg[d], sw);
        double co[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
        for (int j = 0; j < m; j++)
        {
            const int e = L.entry[r0 + j];
            const double* v = e >= 0 ? D + 3*(size_t)e : Db + 3*(size_t)(-1 - e);
            double s[3] = {v[0], v[1], v[2]};
            if (j >= n) { double t[3]; identity_transform(s, t); s[0] = t[0]; s[1] = t[1]; s[2] = t[2]; }
#pragma unroll
            for (int d = 0; d < 3; d++) s[d] = sub(s[d], avg[d]);
#pragma unroll
            for (int a = 0; a < 3; a++)
            {
                const double w = L.inv[3*(size_t)(r0 + j) + a];
#pragma unroll
                for (int d = 0; d < 3; d++) co[a][d] = add(co[a][d], mul(w, s[d]));
            }
        }
        double dr[3];
#pragma unroll
        for (int d = 0; d < 3; d++) dr[d] = sub(points[3*(size_t)p + d], L.origin[3*(size_t)p + d]);
#pragma unroll
        for (int d = 0; d < 3; d++)
            pointD[3*(size_t)p + d] = add(add(add(avg[d], mul(co[0][d], dr[0])), mul(co[1][d], dr[1])), mul(co[2][d], dr[2]));
    }
}

// ---- pf.correctBoundaryConditions(): one thread per constrained point, its patches in boundary order -----------------
//   symmetryPlane [EXT SymmetryPointPatchField::evaluate]  pf = transform(I - nHat*nHat, pf)
//   fixedValue    [EXT ValuePointPatchField::evaluate]     pf = value
__global__ void __launch_bounds__(ATPB) k_point_constraints(PointBcDev B, double* __restrict__ pointD)
{
    for (int i = blockIdx.x*ATPB + threadIdx.x; i < B.n; i += gridDim.x*ATPB)
    {
        double* v = pointD + 3*(size_t)B.point[i];
        double x = v[0], y = v[1], z = v[2];
        for (int r = B.start[i]; r < B.start[i + 1]; r++)
        {
            const double a0 = B.vec[3*(size_t)r], a1 = B.vec[3*(size_t)r + 1], a2 = B.vec[3*(size_t)r + 2];
            if (B.kind[r] == 2) { x = a0; y = a1; z = a2; continue; }
            const double T0 = sub(1.0, mul(a0, a0)), T1 = -mul(a0, a1), T2 = -mul(a0, a2);
            const double T3 = -mul(a1, a0), T4 = sub(1.0, mul(a1, a1)), T5 = -mul(a1, a2);
            const double T6 = -mul(a2, a0), T7 = -mul(a2, a1), T8 = sub(1.0, mul(a2, a2));
            const double nx = add(add(mul(T0, x), mul(T1, y)), mul(T2, z));
            const double ny = add(add(mul(T3, x), mul(T4, y)), mul(T5, z));
            const double nz = add(add(mul(T6, x), mul(T7, y)), mul(T8, z));
            x = nx; y = ny; z = nz;
        }
        v[0] = x; v[1] = y; v[2] = z;
    }
}

// ---- epsilon / sigma with the non-linear and thermal terms (unsTotalLagrangianStress.C:975-990) --------------------
__global__ void __launch_bounds__(ATPB) k_sigma_ex(int nCells, int nonLinear, int thermal, const double* __restrict__ mu,
                                                   const double* __restrict__ lambda, const double* __restrict__ gradD,
                                                   const double* __restrict__ threeK, const double* __restrict__ alpha,
                                                   const double* __restrict__ DT, double* __restrict__ epsilon, double* __restrict__ sigma)
{
    for (int c = blockIdx.x*ATPB + threadIdx.x; c < nCells; c += gridDim.x*ATPB)
    {
        double G[9], e[6], sg[6];
#pragma unroll
        for (int q = 0; q < 9; q++) G[q] = gradD[9*(size_t)c + q];
        symmT(G, e);
        if (nonLinear)
        {
            double GT[9], GG[9], s[6];
            transposeT(G, GT); TdotT(G, GT, GG); symmT(GG, s);
#pragma unroll
            for (int k = 0; k < 6; k++) e[k] = add(e[k], mul(0.5, s[k]));
        }
        const double tr = add(add(e[0], e[3]), e[5]);
        const double sph = mul(lambda[c], tr);
        const double tm = mul(2.0, mu[c]);
#pragma unroll
        for (int k = 0; k < 6; k++) sg[k] = mul(tm, e[k]);
        sg[0] = add(sg[0], sph); sg[3] = add(sg[3], sph); sg[5] = add(sg[5], sph);
        if (thermal)
        {
            const double one[6] = {1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
#pragma unroll
            for (int k = 0; k < 6; k++) sg[k] = sub(sg[k], mul(mul(mul(one[k], threeK[c]), alpha[c]), DT[c]));
        }
#pragma unroll
        for (int k = 0; k < 6; k++) { epsilon[6*(size_t)c + k] = e[k]; sigma[6*(size_t)c + k] = sg[k]; }
    }
}

// min / max of det(I + gradDf) over the faces that carry a face gradient   (:929-946); per-block results (exact)
__global__ void __launch_bounds__(ATPB) k_det_minmax(MeshDev M, const double* __restrict__ gradDf, double* __restrict__ blockMin,
                                                     double* __restrict__ blockMax)
{
    __shared__ double shn[ATPB], shx[ATPB];
    double mn = 1.0e300, mx = -1.0e300;
    for (int f = blockIdx.x*ATPB + threadIdx.x; f < M.nFaces; f += gridDim.x*ATPB)
    {
        if (f >= M.nIF && M.bfPatchType[f - M.nIF] == 4) continue;
        double F[9];
#pragma unroll
        for (int k = 0; k < 9; k++) F[k] = gradDf[9*(size_t)f + k];
        F[0] = add(1.0, F[0]); F[4] = add(1.0, F[4]); F[8] = add(1.0, F[8]);
        const double d = detT(F);
        mn = fmin(mn, d); mx = fmax(mx, d);
    }
    shn[threadIdx.x] = mn; shx[threadIdx.x] = mx;
    __syncthreads();
    for (int o = ATPB/2; o > 0; o >>= 1)
    {
        if (threadIdx.x < o) { shn[threadIdx.x] = fmin(shn[threadIdx.x], shn[threadIdx.x + o]); shx[threadIdx.x] = fmax(shx[threadIdx.x], shx[threadIdx.x + o]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { blockMin[blockIdx.x] = shn[0]; blockMax[blockIdx.x] = shx[0]; }
}

// U = fvc::ddt(D) [EXT EulerDdtScheme::fvcDdt]   (:973)
__global__ void __launch_bounds__(ATPB) k_ddt(long long n3, int scheme, double rDeltaT, const double* __restrict__ D,
                                              const double* __restrict__ Dold, double* __restrict__ U)
{
    for (long long i = (long long)blockIdx.x*ATPB + threadIdx.x; i < n3; i += (long long)gridDim.x*ATPB)
        U[i] = scheme == 1 ? mul(rDeltaT, sub(D[i], Dold[i])) : 0.0;
}

// final reduction of per-block partial maxima / minima: out[0] = max of a[0..n), out[1] = min of b[0..n), one block
__global__ void __launch_bounds__(ATPB) k_final_minmax(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out)
{
    __shared__ double shx[ATPB], shn[ATPB];
    double mx = -1.0e300, mn = 1.0e300;
    for (int i = threadIdx.x; i < n; i += ATPB) { mx = fmax(mx, a[i]); if (b) mn = fmin(mn, b[i]); }
    shx[threadIdx.x] = mx; shn[threadIdx.x] = mn;
    __syncthreads();
    for (int o = ATPB/2; o > 0; o >>= 1)
    {
        if (threadIdx.x < o) { shx[threadIdx.x] = fmax(shx[threadIdx.x], shx[threadIdx.x + o]); shn[threadIdx.x] = fmin(shn[threadIdx.x], shn[threadIdx.x + o]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = shx[0]; out[1] = shn[0]; }
}
