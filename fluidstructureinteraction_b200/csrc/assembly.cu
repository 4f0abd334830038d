// assembly.cu -- GPU assembly of the D equation (K1-K5, K14): the face and cell loops that
//   fvm::laplacian(2*muf+lambdaf, D) (gaussLaplacianScheme [EXT] + skewCorrectedSnGrad<vector>::correction,
//   skewCorrectedVectorSnGrad.C:50-294), fvc::div(Sf & (...gradDf...)) (unsTotalLagrangianStress.C:850-857,
//   fvc::surfaceIntegrate [EXT]) and the displacement boundary conditions (tractionDisplacementFvPatchVectorField.C:
//   148-257, fixedDisplacementFvPatchVectorField.C:129-157) run on the CPU, as three streaming kernels:
//     k_asm_faces   one thread per face : muf/lambdaf interpolation, gammaMagSf, upper, skew-correction flux, stress flux
//     k_asm_cells   one thread per cell : negSumDiag + surfaceIntegrate of both fluxes in the reference face order
//                                         (internal faces ascending, then boundary faces) -> diag, source
//     k_asm_patches one thread per boundary face : internalCoeffs / boundaryCoeffs
//   and k_sigma (epsilon = symm(gradD), sigma = 2 mu eps + lambda tr(eps) I, unsTotalLagrangianStress.C:975-990).
// Every product / sum is a __d*_rn intrinsic in the operand order of the field expressions, so results are
// bit-identical to the CPU evaluation (no FMA).  HBM-bound streaming kernels; no tensor cores.
#include "../../include/gpupcg.h"

#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include "mesh_view.h"
#include <cstdlib>
#include <string>
#include <vector>

namespace
{

constexpr int ATPB = 256;
#ifndef GRAD_CELLS_MINB
#define GRAD_CELLS_MINB 5
#endif
#ifndef FGRAD_MINB
#define FGRAD_MINB 3
#endif
#ifndef ASM_CELLS_MINB
#define ASM_CELLS_MINB 3
#endif
#ifndef ASM_FACES_MINB
#define ASM_FACES_MINB 5
#endif
constexpr double A_SMALL = 1.0e-15;

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
// a/b.  A zero numerator sends __ddiv_rn through its ~65-instruction special-operand routine, and axis-aligned meshes
// are full of them (two of the three components of every face vector).  0/b for a normal b is the signed zero 0*b.
__device__ __forceinline__ double dvd(double a, double b)
{
    const unsigned eb = (unsigned)__double2hiint(b) & 0x7ff00000u;
    if (a == 0.0 && eb - 0x00100000u < 0x7fe00000u) return __dmul_rn(a, b);
    return __ddiv_rn(a, b);
}

__device__ __forceinline__ double dot3(const double* a, const double* b)
{
    return add(add(mul(a[0], b[0]), mul(a[1], b[1])), mul(a[2], b[2]));
}
__device__ __forceinline__ double mag3(const double* a)
{
    return __dsqrt_rn(add(add(mul(a[0], a[0]), mul(a[1], a[1])), mul(a[2], a[2])));
}
// vector & tensor: r_j = v.x*T.xj + v.y*T.yj + v.z*T.zj
__device__ __forceinline__ void vdotT(const double* v, const double* T, double* r)
{
    r[0] = add(add(mul(v[0], T[0]), mul(v[1], T[3])), mul(v[2], T[6]));
    r[1] = add(add(mul(v[0], T[1]), mul(v[1], T[4])), mul(v[2], T[7]));
    r[2] = add(add(mul(v[0], T[2]), mul(v[1], T[5])), mul(v[2], T[8]));
}

struct MeshDev
{
    int nCells, nFaces, nIF, nPatches;
    const int *owner, *neighbour, *bfPatchType;      // bfPatchType[nFaces - nIF]
    const int *cellFaceStart, *cellFace;              // CSR; face id, bit 31 set = the cell is the face's neighbour
    const double *Sf, *magSf, *Cf, *C, *V, *weights, *deltaCoeffs;
    int nPoints;
    const int *faceStart, *facePoints;                // point loops of the faces (vertex-based gradients)
    const double* points;
    // geometry of the faces' triangle fans, pre-computed by gpupcg_mesh_set_points (k_fan_geometry): per (face, fan
    // triangle) the area vector St = ((a - centre) x (b - centre))/2 and St & Ct; per cell V = (sum +-St & Ct)/3
    const double *fanS, *fanV, *gradV;
};


// symmetryDisplacementFvPatchVectorField::snGrad()  (symmetryDisplacementFvPatchVectorField.C:134-165)
__device__ void symmetry_snGrad(const MeshDev& M, int f, const double* __restrict__ D, const double* __restrict__ gradD, double* sn)
{
    const int c = M.owner[f];
    const double magSf = M.magSf[f];
    double n[3], delta[3], k[3], kg[3], UP[3], G[9], S[6], T[6], tv[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { n[d] = dvd(M.Sf[3*(size_t)f + d], magSf); delta[d] = sub(M.Cf[3*(size_t)f + d], M.C[3*(size_t)c + d]); }
    const double nd = dot3(n, delta);
#pragma unroll
    for (int d = 0; d < 3; d++) k[d] = sub(delta[d], mul(n[d], nd));
#pragma unroll
    for (int q = 0; q < 9; q++) G[q] = gradD[9*(size_t)c + q];
    vdotT(k, G, kg);
#pragma unroll
    for (int d = 0; d < 3; d++) UP[d] = add(D[3*(size_t)c + d], kg[d]);
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

#include "assembly_ext.cuh"
#include "assembly_scalar.cuh"

// One chunk of the assembly (see AsmChunk): slot s of the chunk's flux ring is internal face f0 + s for s < nOwn (the
// faces the chunk's cells own), then the chunk's extra faces: internal faces owned by an EARLIER chunk whose neighbour
// cell is in this chunk (recomputed here: 1-2 % of the faces), then the boundary faces of the chunk's cells.
struct AsmChunk { int c0, c1, f0, nOwn, extraStart, nExtra; };

__global__ void __launch_bounds__(ATPB, ASM_FACES_MINB) k_asm_faces(MeshDev M, AsmChunk K, const int* __restrict__ extraFace,
                                                    const double* __restrict__ mu, const double* __restrict__ lambda,
                                                    const double* __restrict__ D, const double* __restrict__ gradD,
                                                    const double* __restrict__ gradDf, double limitCoeff,
                                                    double* __restrict__ upper, double* __restrict__ Lup, double* __restrict__ gms,
                                                    double* __restrict__ fluxE, double* __restrict__ fluxC)
{
    for (int sl = blockIdx.x*ATPB + threadIdx.x; sl < K.nOwn + K.nExtra; sl += gridDim.x*ATPB)
    {
        const int f = (sl < K.nOwn) ? K.f0 + sl : extraFace[K.extraStart + sl - K.nOwn];
        const int P = M.owner[f];
        const bool internal = f < M.nIF;
        if (!internal && M.bfPatchType[f - M.nIF] == 4)
        {
            // empty patch (2-D cases): its fvsPatchFields have no faces -- no flux, no coefficients
            gms[f] = 0.0;
            fluxE[3*(size_t)sl] = 0.0; fluxE[3*(size_t)sl + 1] = 0.0; fluxE[3*(size_t)sl + 2] = 0.0;
            continue;
        }
        const int N = internal ? M.neighbour[f] : P;
        double muf, laf;
        if (internal)
        {
            const double w = M.weights[f];
            muf = add(mul(w, sub(mu[P], mu[N])), mu[N]);
            laf = add(mul(w, sub(lambda[P], lambda[N])), lambda[N]);
        }
        else { muf = mu[P]; laf = lambda[P]; }
        const double g = mul(add(mul(2.0, muf), laf), M.magSf[f]);
        if (!internal) gms[f] = g;                  // k_asm_patches reads it
        double Sf[3] = {M.Sf[3*f], M.Sf[3*f + 1], M.Sf[3*f + 2]};
        // explicit stress flux: Sf & ( -(muf+lambdaf)*G + muf*G.T() + lambdaf*(I*tr(G)) )
        {
            double G[9];
#pragma unroll
            for (int k = 0; k < 9; k++) G[k] = gradDf[9*(size_t)f + k];
            const double a = -add(muf, laf);
            const double tr = add(add(G[0], G[4]), G[8]);
            const double sph = mul(laf, tr);
            double T[9];
            T[0] = add(add(mul(a, G[0]), mul(muf, G[0])), sph); T[1] = add(mul(a, G[1]), mul(muf, G[3])); T[2] = add(mul(a, G[2]), mul(muf, G[6]));
            T[3] = add(mul(a, G[3]), mul(muf, G[1])); T[4] = add(add(mul(a, G[4]), mul(muf, G[4])), sph); T[5] = add(mul(a, G[5]), mul(muf, G[7]));
            T[6] = add(mul(a, G[6]), mul(muf, G[2])); T[7] = add(mul(a, G[7]), mul(muf, G[5])); T[8] = add(add(mul(a, G[8]), mul(muf, G[8])), sph);
            double fl[3];
            vdotT(Sf, T, fl);
            fluxE[3*(size_t)sl] = fl[0]; fluxE[3*(size_t)sl + 1] = fl[1]; fluxE[3*(size_t)sl + 2] = fl[2];
        }
        if (internal)
        {
            const double dc = M.deltaCoeffs[f];
            const double up = mul(dc, g);
            if (Lup) Lup[sl] = up;                      // one chunk: slot == face, the cell kernel reads `upper` instead
            if (sl < K.nOwn) upper[f] = sub(0.0, up);
            // skewCorrectedSnGrad<vector>::correction
            const double Cf[3] = {M.Cf[3*f], M.Cf[3*f + 1], M.Cf[3*f + 2]};
            const double magSqrSf = add(add(mul(Sf[0], Sf[0]), mul(Sf[1], Sf[1])), mul(Sf[2], Sf[2]));
            double kP[3], kN[3], a3[3], b3[3], ssf[3], t[3], GP[9], GN[9];
#pragma unroll
            for (int d = 0; d < 3; d++) kP[d] = sub(Cf[d], M.C[3*(size_t)P + d]);
            {
                const double s = dot3(Sf, kP);
#pragma unroll
                for (int d = 0; d < 3; d++) kP[d] = sub(kP[d], dvd(mul(Sf[d], s), magSqrSf));
            }
#pragma unroll
            for (int d = 0; d < 3; d++) kN[d] = sub(Cf[d], M.C[3*(size_t)N + d]);
            {
                const double s = dot3(Sf, kN);
#pragma unroll
                for (int d = 0; d < 3; d++) kN[d] = sub(kN[d], dvd(mul(Sf[d], s), magSqrSf));
            }
#pragma unroll
            for (int k = 0; k < 9; k++) { GP[k] = gradD[9*(size_t)P + k]; GN[k] = gradD[9*(size_t)N + k]; }
            vdotT(kN, GN, a3);
            vdotT(kP, GP, b3);
#pragma unroll
            for (int d = 0; d < 3; d++) ssf[d] = mul(sub(a3[d], b3[d]), dc);
#pragma unroll
            for (int d = 0; d < 3; d++) t[d] = add(mul(dc, sub(D[3*(size_t)N + d], D[3*(size_t)P + d])), ssf[d]);
            double lim = dvd(mul(limitCoeff, mag3(t)), add(mul(sub(1.0, limitCoeff), mag3(ssf)), A_SMALL));
            if (lim > 1.0) lim = 1.0;
#pragma unroll
            for (int d = 0; d < 3; d++) fluxC[3*(size_t)sl + d] = mul(g, mul(ssf[d], lim));
        }
    }
}

// cellSlot[k] (parallel to the cell -> face CSR): ring slot of the face | bit 31: the cell is the face's neighbour |
// bit 30: boundary face
__global__ void __launch_bounds__(ATPB, ASM_CELLS_MINB) k_asm_cells(MeshDev M, AsmChunk K, AsmCtl T, const double* __restrict__ Dold,
                                                    const double* __restrict__ Doldold, const int* __restrict__ cellSlot,
                                                    const double* __restrict__ rho, double g0, double g1, double g2,
                                                    const double* __restrict__ Lup, const double* __restrict__ upperSingle,
                                                    const double* __restrict__ fluxE,
                                                    const double* __restrict__ fluxC, double* __restrict__ diag,
                                                    double* __restrict__ source, const double* __restrict__ timeA = nullptr)
{
    for (int c = K.c0 + blockIdx.x*ATPB + threadIdx.x; c < K.c1; c += gridDim.x*ATPB)
    {
        double Ld = 0.0, sE[3] = {0.0, 0.0, 0.0}, sC[3] = {0.0, 0.0, 0.0};
        // three faces per pass: their slots first, then all 21 flux loads in flight together, then the sums in face order
        const int kEnd = M.cellFaceStart[c + 1];
        for (int k0 = M.cellFaceStart[c]; k0 < kEnd; k0 += 3)
        {
            int code[3];
            double fe[3][3], fc[3][3], lu[3];
#pragma unroll
            for (int j = 0; j < 3; j++) code[j] = (k0 + j < kEnd) ? cellSlot[k0 + j] : 0x40000000;     // pad: boundary-type slot 0
#pragma unroll
            for (int j = 0; j < 3; j++)
            {
                const bool have = k0 + j < kEnd;
                const bool internal = !(code[j] & 0x40000000);
                const size_t sl = (size_t)(code[j] & 0x3fffffff);
#pragma unroll
                for (int d = 0; d < 3; d++)
                {
                    fe[j][d] = have ? fluxE[3*sl + d] : 0.0;
                    fc[j][d] = (have && internal) ? fluxC[3*sl + d] : 0.0;
                }
                // Ld - up == Ld + upper exactly (upper = 0 - up, and Ld is never -0)
                lu[j] = (have && internal) ? (Lup ? Lup[sl] : -upperSingle[sl]) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 3; j++)
            {
                if (k0 + j >= kEnd) break;
                const bool nbr = code[j] < 0;
                if (!(code[j] & 0x40000000))
                {
                    Ld = sub(Ld, lu[j]);
#pragma unroll
                    for (int d = 0; d < 3; d++)
                    {
                        sC[d] = nbr ? sub(sC[d], fc[j][d]) : add(sC[d], fc[j][d]);
                        sE[d] = nbr ? sub(sE[d], fe[j][d]) : add(sE[d], fe[j][d]);
                    }
                }
                else
                {
#pragma unroll
                    for (int d = 0; d < 3; d++) sE[d] = add(sE[d], fe[j][d]);
                }
            }
        }
        const double V = M.V[c];
        const double rh = rho[c];
        // A = rho*fvm::d2dt2(D): steadyState contributes nothing; Euler [EXT EulerD2dt2Scheme, weights as in the
        // reference's backwardD2dt2Scheme.C:306-331]: diag = (coefft*rDeltaT2)*V, source = rDeltaT2*V*((coefft+coefft00)*
        // D.oldTime() - coefft00*D.oldTime().oldTime()), both then scaled by rho
        double dA = 0.0, sA[3] = {0.0, 0.0, 0.0};
        if (T.d2dt2 == 1)
        {
            dA = mul(mul(mul(T.coefft, T.rDeltaT2), V), rh);
            const double t1 = mul(T.rDeltaT2, V), cc = add(T.coefft, T.coefft00);
#pragma unroll
            for (int d = 0; d < 3; d++)
                sA[d] = mul(mul(t1, sub(mul(cc, Dold[3*(size_t)c + d]), mul(T.coefft00, Doldold[3*(size_t)c + d]))), rh);
        }
        else if (T.d2dt2 == 2)      // backward: evaluated per cell by k_time_backward (assembly_ext.cuh)
        {
            dA = timeA[4*(size_t)c];
#pragma unroll
            for (int d = 0; d < 3; d++) sA[d] = timeA[4*(size_t)c + 1 + d];
        }
        double dg = sub(dA, Ld);
        const double gv[3] = {g0, g1, g2};
        double src[3];
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            const double dcv = dvd(sC[d], V), de = dvd(sE[d], V);
            const double rg = mul(rh, gv[d]);
            const double R = sub(sub(sub(0.0, mul(V, dcv)), mul(V, de)), mul(V, rg));
            src[d] = sub(sA[d], R);
        }
        if (T.damping)
        {
            // DEqn += K*rho_*fvm::ddt(D_)   (unsTotalLagrangianStress.C:861-865; [EXT] EulerDdtScheme::fvmDdt)
            const double sf = mul(T.K, rh);
            if (T.ddt == 1)
            {
                dg = add(dg, mul(mul(T.rDeltaT, V), sf));
#pragma unroll
                for (int d = 0; d < 3; d++) src[d] = add(src[d], mul(mul(mul(T.rDeltaT, Dold[3*(size_t)c + d]), V), sf));
            }
            else if (T.ddt == 2)    // [EXT] backwardDdtScheme::fvmDdt, then scaled by K*rho
            {
                dg = add(dg, mul(mul(mul(T.cbD, T.rDeltaT), V), sf));
                const double rV = mul(T.rDeltaT, V);
#pragma unroll
                for (int d = 0; d < 3; d++)
                    src[d] = add(src[d], mul(mul(rV, sub(mul(T.cb0D, Dold[3*(size_t)c + d]), mul(T.cb00D, Doldold[3*(size_t)c + d]))), sf));
            }
            else
            {
                dg = add(dg, 0.0);
#pragma unroll
                for (int d = 0; d < 3; d++) src[d] = add(src[d], 0.0);
            }
        }
        diag[c] = dg;
#pragma unroll
        for (int d = 0; d < 3; d++) source[3*(size_t)c + d] = src[d];
    }
}

__global__ void __launch_bounds__(ATPB) k_asm_patches(MeshDev M, AsmCtl T, const double* __restrict__ threeK, const double* __restrict__ alpha,
                                                      const double* __restrict__ DTb,
                                                      const double* __restrict__ mu, const double* __restrict__ lambda,
                                                      const double* __restrict__ D,
                                                      const double* __restrict__ gradD, const double* __restrict__ gradDf,
                                                      const double* __restrict__ gms, const double* __restrict__ traction,
                                                      const double* __restrict__ pressure, const double* __restrict__ fixedValue,
                                                      double* __restrict__ internalCoeffs, double* __restrict__ boundaryCoeffs,
                                                      double* __restrict__ patchGrad)
{
    const int nBF = M.nFaces - M.nIF;
    for (int bf = blockIdx.x*ATPB + threadIdx.x; bf < nBF; bf += gridDim.x*ATPB)
    {
        const int f = M.nIF + bf, c = M.owner[f];
        const double Sf[3] = {M.Sf[3*f], M.Sf[3*f + 1], M.Sf[3*f + 2]};
        const double magSf = M.magSf[f], dcf = M.deltaCoeffs[f];
        double n[3], gIC[3], gBC[3];
        const int ptype = M.bfPatchType[bf];
        if (ptype == 4)
        {
#pragma unroll
            for (int d = 0; d < 3; d++) { internalCoeffs[3*(size_t)bf + d] = 0.0; boundaryCoeffs[3*(size_t)bf + d] = 0.0; }
            continue;
        }
#pragma unroll
        for (int d = 0; d < 3; d++) n[d] = dvd(Sf[d], magSf);
        if (ptype == 1 || ptype == 5)
        {
            double delta[3], k[3], kg[3], G[9];
#pragma unroll
            for (int d = 0; d < 3; d++) delta[d] = sub(M.Cf[3*(size_t)f + d], M.C[3*(size_t)c + d]);
            const double nd = dot3(n, delta);
#pragma unroll
            for (int d = 0; d < 3; d++) k[d] = sub(delta[d], mul(n[d], nd));
#pragma unroll
            for (int q = 0; q < 9; q++) G[q] = gradD[9*(size_t)c + q];
            vdotT(k, G, kg);
            if (ptype == 5) kg[0] = kg[1] = kg[2] = 0.0;        // [EXT] fixedValueFvPatchField: no non-orthogonal term
#pragma unroll
            for (int d = 0; d < 3; d++)
            {
                gIC[d] = -dcf;
                gBC[d] = mul(dcf, sub(fixedValue[3*(size_t)bf + d], kg[d]));
            }
        }
        else if (ptype == 2 || ptype == 3)
        {
            double sn[3];
            if (ptype == 2) symmetry_snGrad(M, f, D, gradD, sn); else symmetryPlane_snGrad(M, f, D, sn);
#pragma unroll
            for (int d = 0; d < 3; d++)
            {
                gIC[d] = mul(-dcf, fabs(n[d]));
                gBC[d] = sub(sn[d], mul(gIC[d], D[3*(size_t)c + d]));
            }
        }
        else
        {
            const double m = mu[c], la = lambda[c];
            double G[9];
#pragma unroll
            for (int q = 0; q < 9; q++) G[q] = gradDf[9*(size_t)f + q];
            const double tr3[3] = {traction[3*(size_t)bf], traction[3*(size_t)bf + 1], traction[3*(size_t)bf + 2]};
            traction_gradient(M, T, f, m, la, G, tr3, pressure[bf], T.thermal ? threeK[c] : 0.0, T.thermal ? alpha[c] : 0.0,
                              T.thermal ? DTb[bf] : 0.0, gBC);
#pragma unroll
            for (int d = 0; d < 3; d++)
            {
                gIC[d] = 0.0;
                patchGrad[3*(size_t)bf + d] = gBC[d];       // the patch's gradient(): fvc::fGrad reads it next
            }
        }
        const double g = gms[f];
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            internalCoeffs[3*(size_t)bf + d] = sub(0.0, mul(g, gIC[d]));
            boundaryCoeffs[3*(size_t)bf + d] = sub(0.0, mul(-g, gBC[d]));
        }
    }
}

__global__ void __launch_bounds__(ATPB) k_sigma(int nCells, const double* __restrict__ mu, const double* __restrict__ lambda,
                                                const double* __restrict__ gradD, double* __restrict__ epsilon, double* __restrict__ sigma)
{
    for (int c = blockIdx.x*ATPB + threadIdx.x; c < nCells; c += gridDim.x*ATPB)
    {
        double G[9], e[6];
#pragma unroll
        for (int q = 0; q < 9; q++) G[q] = gradD[9*(size_t)c + q];
        e[0] = G[0]; e[1] = mul(0.5, add(G[1], G[3])); e[2] = mul(0.5, add(G[2], G[6]));
        e[3] = G[4]; e[4] = mul(0.5, add(G[5], G[7])); e[5] = G[8];
        const double tr = add(add(e[0], e[3]), e[5]);
        const double sph = mul(lambda[c], tr);
        const double tm = mul(2.0, mu[c]);
#pragma unroll
        for (int q = 0; q < 6; q++) epsilon[6*(size_t)c + q] = e[q];
        sigma[6*(size_t)c + 0] = add(mul(tm, e[0]), sph); sigma[6*(size_t)c + 1] = mul(tm, e[1]); sigma[6*(size_t)c + 2] = mul(tm, e[2]);
        sigma[6*(size_t)c + 3] = add(mul(tm, e[3]), sph); sigma[6*(size_t)c + 4] = mul(tm, e[4]); sigma[6*(size_t)c + 5] = add(mul(tm, e[5]), sph);
    }
}

// ---- vertex-based gradients of the reference (numerics/fvc/fvcGradf.C) ---------------------------------------
__device__ __forceinline__ void cross3(const double* a, const double* b, double* r)
{
    r[0] = sub(mul(a[1], b[2]), mul(a[2], b[1]));
    r[1] = sub(mul(a[2], b[0]), mul(a[0], b[2]));
    r[2] = sub(mul(a[0], b[1]), mul(a[1], b[0]));
}

// x/3.0, correctly rounded, without the IEEE division sequence: q0 = x*RN(1/3), exact residual by FMA, one
// correction (Markstein: with y = RN(1/b) and q0 within an ulp of a/b, q0 + (a - b*q0)*y rounds to RN(a/b); b = 3 is not
// one of the all-ones-significand exceptions).  Only inside 2^-900 <= |x| < 2^901, where neither the residual nor the
// correction can lose bits to underflow; zeros, subnormals, huge values, Inf and NaN take the division.
// tools/div3_check.c compares it with x/3.0 on 4.6e8 operands (0 mismatches).
__device__ __forceinline__ double div3(double x)
{
    const unsigned e = (unsigned)__double2hiint(x) & 0x7ff00000u;
    if (e - 0x07b00000u < 0x70900000u)
    {
        const double r = 0x1.5555555555555p-2;
        const double q0 = __dmul_rn(x, r);
        const double rem = __fma_rn(-3.0, q0, x);
        return __fma_rn(rem, r, q0);
    }
    return __ddiv_rn(x, 3.0);
}
__device__ __forceinline__ void load3(const double* __restrict__ a, int i, double* r)
{
    r[0] = a[3*(size_t)i]; r[1] = a[3*(size_t)i + 1]; r[2] = a[3*(size_t)i + 2];
}

// ---- pre-computed fan geometry --------------------------------------------------------------------------------------
// The geometric half of grad_face (area vectors and volume terms of the fan triangles) depends on the mesh only; it is
// evaluated once per gpupcg_mesh_set_points with the same operations in the same order, so k_grad_cells below gives
// the same bits as the all-in-one evaluation (and the oracle) while doing a third of the arithmetic per call.
__global__ void __launch_bounds__(ATPB) k_fan_geometry(MeshDev M, double* __restrict__ fanS, double* __restrict__ fanV)
{
    const double* X = M.points;
    for (int f = blockIdx.x*ATPB + threadIdx.x; f < M.nFaces; f += gridDim.x*ATPB)
    {
        const int p0 = M.faceStart[f], n = M.faceStart[f + 1] - p0;
        const int* fp = M.facePoints + p0;
        if (n == 3)
        {
            const int i0 = fp[0], i1 = fp[1], i2 = fp[2];
            double ab[3], ac[3], nrm[3], ctr[3];
#pragma unroll
            for (int d = 0; d < 3; d++) { ab[d] = sub(X[3*i1 + d], X[3*i0 + d]); ac[d] = sub(X[3*i2 + d], X[3*i0 + d]); }
            cross3(ab, ac, nrm);
#pragma unroll
            for (int d = 0; d < 3; d++)
            {
                nrm[d] = mul(0.5, nrm[d]);
                ctr[d] = mul(1.0/3.0, add(add(X[3*i0 + d], X[3*i1 + d]), X[3*i2 + d]));
                fanS[3*(size_t)p0 + d] = nrm[d];
            }
            fanV[p0] = dot3(nrm, ctr);
            for (int p = 1; p < 3; p++) { fanV[p0 + p] = 0.0; fanS[3*(size_t)(p0 + p)] = fanS[3*(size_t)(p0 + p) + 1] = fanS[3*(size_t)(p0 + p) + 2] = 0.0; }
            continue;
        }
        double centre[3] = {0.0, 0.0, 0.0};
        for (int p = 0; p < n; p++)
        {
            const int ip = fp[p];
#pragma unroll
            for (int d = 0; d < 3; d++) centre[d] = add(centre[d], X[3*ip + d]);
        }
        const double dn = (double)n;
#pragma unroll
        for (int d = 0; d < 3; d++) centre[d] = (n == 4) ? mul(centre[d], 0.25) : dvd(centre[d], dn);
        double Xa[3], Xb[3];
        load3(X, fp[0], Xa);
        for (int p = 0; p < n; p++)
        {
            load3(X, fp[(p + 1 == n) ? 0 : p + 1], Xb);
            double a[3], b[3], St[3], Ct[3];
#pragma unroll
            for (int d = 0; d < 3; d++) { a[d] = sub(Xa[d], centre[d]); b[d] = sub(Xb[d], centre[d]); }
            cross3(a, b, St);
#pragma unroll
            for (int d = 0; d < 3; d++)
            {
                St[d] = mul(St[d], 0.5);
                Ct[d] = div3(add(add(centre[d], Xa[d]), Xb[d]));
                fanS[3*(size_t)(p0 + p) + d] = St[d];
            }
            fanV[p0 + p] = dot3(St, Ct);
#pragma unroll
            for (int d = 0; d < 3; d++) Xa[d] = Xb[d];
        }
    }
}

// V_c = (sum over the cell's faces, fan triangle by fan triangle, of +-St & Ct)/3   (fvcGradf.C:537,554-555,607-608,688-690)
__global__ void __launch_bounds__(ATPB) k_cell_fan_volume(MeshDev M, const double* __restrict__ fanV, double* __restrict__ gradV)
{
    for (int c = blockIdx.x*ATPB + threadIdx.x; c < M.nCells; c += gridDim.x*ATPB)
    {
        double V = 0.0;
        for (int k = M.cellFaceStart[c]; k < M.cellFaceStart[c + 1]; k++)
        {
            const int code = M.cellFace[k], f = code & 0x7fffffff;
            const bool minus = code < 0;
            const int p0 = M.faceStart[f], n = M.faceStart[f + 1] - p0;
            const int nt = (n == 3) ? 1 : n;
            for (int p = 0; p < nt; p++) V = minus ? sub(V, fanV[p0 + p]) : add(V, fanV[p0 + p]);
        }
        gradV[c] = div3(V);
    }
}

// one face's contribution sign*(St*ttcf) with the pre-computed area vectors
__device__ __forceinline__ void grad_face_pre(const MeshDev& M, int f, const double* __restrict__ pf, bool minus, double* G)
{
    const int p0 = M.faceStart[f], n = M.faceStart[f + 1] - p0;
    const int* fp = M.facePoints + p0;
    const double* S = M.fanS + 3*(size_t)p0;
    if (n == 3)
    {
        const int i0 = fp[0], i1 = fp[1], i2 = fp[2];
        double avg[3], nrm[3];
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            avg[d] = mul(1.0/3.0, add(add(pf[3*i0 + d], pf[3*i1 + d]), pf[3*i2 + d]));
            nrm[d] = minus ? -S[d] : S[d];
        }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) G[3*i + j] = add(G[3*i + j], mul(nrm[i], avg[j]));
        return;
    }
    if (n == 4)
    {
        double Pq[4][3], cf[3];
#pragma unroll
        for (int p = 0; p < 4; p++) load3(pf, fp[p], Pq[p]);
#pragma unroll
        for (int d = 0; d < 3; d++) cf[d] = mul(add(add(add(add(0.0, Pq[0][d]), Pq[1][d]), Pq[2][d]), Pq[3][d]), 0.25);
#pragma unroll
        for (int p = 0; p < 4; p++)
        {
            double St[3], ttcf[3];
#pragma unroll
            for (int d = 0; d < 3; d++)
            {
                St[d] = minus ? -S[3*p + d] : S[3*p + d];
                ttcf[d] = div3(add(add(Pq[p][d], Pq[(p + 1) & 3][d]), cf[d]));
            }
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) G[3*i + j] = add(G[3*i + j], mul(St[i], ttcf[j]));
        }
        return;
    }
    double cf[3] = {0.0, 0.0, 0.0};
    for (int p = 0; p < n; p++)
    {
        const int ip = fp[p];
#pragma unroll
        for (int d = 0; d < 3; d++) cf[d] = add(cf[d], pf[3*ip + d]);
    }
    const double dn = (double)n;
#pragma unroll
    for (int d = 0; d < 3; d++) cf[d] = dvd(cf[d], dn);
    double Pa[3], Pb[3];
    load3(pf, fp[0], Pa);
    for (int p = 0; p < n; p++)
    {
        load3(pf, fp[(p + 1 == n) ? 0 : p + 1], Pb);
        double St[3], ttcf[3];
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            St[d] = minus ? -S[3*p + d] : S[3*p + d];
            ttcf[d] = div3(add(add(Pa[d], Pb[d]), cf[d]));
            Pa[d] = Pb[d];
        }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) G[3*i + j] = add(G[3*i + j], mul(St[i], ttcf[j]));
    }
}

// one edge (s -> e) of a face: Le*fe into G   (fvcGradf.C:97-142)
__device__ __forceinline__ void tangential_edge(const double* Xs, const double* Xe, const double* Ps, const double* Pe,
                                                const double* nf, double* G)
{
    double e[3], Le[3], fe[3];
#pragma unroll
    for (int d = 0; d < 3; d++) e[d] = sub(Xe[d], Xs[d]);
    const double ne = dot3(nf, e);
#pragma unroll
    for (int d = 0; d < 3; d++) e[d] = sub(e[d], mul(nf[d], ne));
    cross3(e, nf, Le);
#pragma unroll
    for (int d = 0; d < 3; d++) fe[d] = mul(0.5, add(Ps[d], Pe[d]));
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) G[3*i + j] = add(G[3*i + j], mul(Le[i], fe[j]));
}

// tangential face gradient from point values: sum_edges Le*fe / mag   (fvcGradf.C:97-142, 144-198, 410-482)
__device__ void face_tangential_grad(const MeshDev& M, int f, const double* __restrict__ pf, double* G)
{
    const int p0 = M.faceStart[f], n = M.faceStart[f + 1] - p0;
    const int* fp = M.facePoints + p0;
    const double* X = M.points;
    const double magSf = M.magSf[f];
    double nf[3];
#pragma unroll
    for (int d = 0; d < 3; d++) nf[d] = dvd(M.Sf[3*(size_t)f + d], magSf);
#pragma unroll
    for (int k = 0; k < 9; k++) G[k] = 0.0;
    if (n == 4)
    {
        double Xq[4][3], Pq[4][3];
#pragma unroll
        for (int p = 0; p < 4; p++) { const int ip = fp[p]; load3(X, ip, Xq[p]); load3(pf, ip, Pq[p]); }
#pragma unroll
        for (int p = 0; p < 4; p++) tangential_edge(Xq[p], Xq[(p + 1) & 3], Pq[p], Pq[(p + 1) & 3], nf, G);
    }
    else if (n == 3)
    {
        double Xq[3][3], Pq[3][3];
#pragma unroll
        for (int p = 0; p < 3; p++) { const int ip = fp[p]; load3(X, ip, Xq[p]); load3(pf, ip, Pq[p]); }
#pragma unroll
        for (int p = 0; p < 3; p++) tangential_edge(Xq[p], Xq[(p + 1) % 3], Pq[p], Pq[(p + 1) % 3], nf, G);
    }
    else
    {
        double Xs[3], Ps[3], Xe[3], Pe[3];
        load3(X, fp[0], Xs); load3(pf, fp[0], Ps);
        for (int p = 0; p < n; p++)
        {
            const int ie = fp[(p + 1 == n) ? 0 : p + 1];
            load3(X, ie, Xe); load3(pf, ie, Pe);
            tangential_edge(Xs, Xe, Ps, Pe, nf, G);
#pragma unroll
            for (int d = 0; d < 3; d++) { Xs[d] = Xe[d]; Ps[d] = Pe[d]; }
        }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) G[k] = dvd(G[k], magSf);
}

// boundary snGrad of D: fixedDisplacement::snGrad() (fixedDisplacementFvPatchVectorField.C:96-127) or fixedGradient's gradient()
__device__ void patch_snGrad(const MeshDev& M, int f, const double* __restrict__ D, const double* __restrict__ gradD,
                             const double* __restrict__ fixedValue, const double* __restrict__ patchGradient, double* sn)
{
    const int bf = f - M.nIF, c = M.owner[f];
    if (M.bfPatchType[bf] == 1 || M.bfPatchType[bf] == 5)
    {
        double n[3], delta[3], k[3], kg[3], G[9];
        const double magSf = M.magSf[f];
#pragma unroll
        for (int d = 0; d < 3; d++) { n[d] = dvd(M.Sf[3*(size_t)f + d], magSf); delta[d] = sub(M.Cf[3*(size_t)f + d], M.C[3*(size_t)c + d]); }
        const double nd = dot3(n, delta);
#pragma unroll
        for (int d = 0; d < 3; d++) k[d] = sub(delta[d], mul(n[d], nd));
#pragma unroll
        for (int q = 0; q < 9; q++) G[q] = gradD[9*(size_t)c + q];
        vdotT(k, G, kg);
        if (M.bfPatchType[bf] == 5) kg[0] = kg[1] = kg[2] = 0.0;   // [EXT] fixedValueFvPatchField::snGrad
#pragma unroll
        for (int d = 0; d < 3; d++) sn[d] = mul(sub(fixedValue[3*(size_t)bf + d], add(D[3*(size_t)c + d], kg[d])), M.deltaCoeffs[f]);
    }
    else if (M.bfPatchType[bf] == 2) symmetry_snGrad(M, f, D, gradD, sn);
    else if (M.bfPatchType[bf] == 3) symmetryPlane_snGrad(M, f, D, sn);
    else
    {
#pragma unroll
        for (int d = 0; d < 3; d++) sn[d] = patchGradient[3*(size_t)bf + d];
    }
}

// K6: fvc::grad(D, pointD) -- one thread per cell gathers its faces in the order the reference's face loop visits it
// gather-latency bound (cell -> faces -> points -> values): small blocks, as many as fit
constexpr int GTPB = 128;
__global__ void __launch_bounds__(GTPB, GRAD_CELLS_MINB) k_grad_cells(MeshDev M, const double* __restrict__ pointD, double* __restrict__ gradD)
{
    for (int c = blockIdx.x*GTPB + threadIdx.x; c < M.nCells; c += gridDim.x*GTPB)
    {
        double G[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const int k1 = M.cellFaceStart[c + 1];
        for (int k = M.cellFaceStart[c]; k < k1; k++)
        {
            const int code = M.cellFace[k];
            grad_face_pre(M, code & 0x7fffffff, pointD, code < 0, G);
        }
        const double V = M.gradV[c];
#pragma unroll
        for (int q = 0; q < 9; q++) gradD[9*(size_t)c + q] = dvd(G[q], V);
    }
}

// boundary field of grad(D): patch fGrad + gaussGrad::correctBoundaryConditions (fvcGradf.C:696-714)
__global__ void __launch_bounds__(ATPB) k_grad_patches(MeshDev M, const double* __restrict__ D, const double* __restrict__ pointD,
                                                       const double* __restrict__ gradDold, const double* __restrict__ fixedValue,
                                                       const double* __restrict__ patchGradient, double* __restrict__ gradDb)
{
    const int nBF = M.nFaces - M.nIF;
    for (int bf = blockIdx.x*ATPB + threadIdx.x; bf < nBF; bf += gridDim.x*ATPB)
    {
        const int f = M.nIF + bf;
        if (M.bfPatchType[bf] == 4) continue;           // empty: the fvPatch has no faces
        double G[9], n[3], sn[3], nG[3];
        face_tangential_grad(M, f, pointD, G);
        const double magSf = M.magSf[f];
#pragma unroll
        for (int d = 0; d < 3; d++) n[d] = dvd(M.Sf[3*(size_t)f + d], magSf);
        patch_snGrad(M, f, D, gradDold, fixedValue, patchGradient, sn);
        vdotT(n, G, nG);
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) gradDb[9*(size_t)bf + 3*a + b] = add(G[3*a + b], mul(n[a], sub(sn[b], nG[b])));
    }
}

// K7: fvc::fGrad(D, pointD) -- one thread per face: tangential part + n*snGrad(D) (skewCorrected scheme / patch snGrad)
__global__ void __launch_bounds__(ATPB, FGRAD_MINB) k_fgrad_faces(MeshDev M, const double* __restrict__ D, const double* __restrict__ pointD,
                                                      const double* __restrict__ gradD, const double* __restrict__ fixedValue,
                                                      const double* __restrict__ patchGradient, double limitCoeff,
                                                      double* __restrict__ gradDf)
{
    for (int f = blockIdx.x*ATPB + threadIdx.x; f < M.nFaces; f += gridDim.x*ATPB)
    {
        if (f >= M.nIF && M.bfPatchType[f - M.nIF] == 4) continue;
        double G[9], n[3], sn[3];
        face_tangential_grad(M, f, pointD, G);
        const double magSf = M.magSf[f];
        const double Sf[3] = {M.Sf[3*(size_t)f], M.Sf[3*(size_t)f + 1], M.Sf[3*(size_t)f + 2]};
#pragma unroll
        for (int d = 0; d < 3; d++) n[d] = dvd(Sf[d], magSf);
        if (f < M.nIF)
        {
            const int P = M.owner[f], N = M.neighbour[f];
            const double dc = M.deltaCoeffs[f];
            const double Cf[3] = {M.Cf[3*(size_t)f], M.Cf[3*(size_t)f + 1], M.Cf[3*(size_t)f + 2]};
            const double magSqrSf = add(add(mul(Sf[0], Sf[0]), mul(Sf[1], Sf[1])), mul(Sf[2], Sf[2]));
            double kP[3], kN[3], a3[3], b3[3], ssf[3], t[3], GP[9], GN[9];
#pragma unroll
            for (int d = 0; d < 3; d++) kP[d] = sub(Cf[d], M.C[3*(size_t)P + d]);
            {
                const double s = dot3(Sf, kP);
#pragma unroll
                for (int d = 0; d < 3; d++) kP[d] = sub(kP[d], dvd(mul(Sf[d], s), magSqrSf));
            }
#pragma unroll
            for (int d = 0; d < 3; d++) kN[d] = sub(Cf[d], M.C[3*(size_t)N + d]);
            {
                const double s = dot3(Sf, kN);
#pragma unroll
                for (int d = 0; d < 3; d++) kN[d] = sub(kN[d], dvd(mul(Sf[d], s), magSqrSf));
            }
#pragma unroll
            for (int q = 0; q < 9; q++) { GP[q] = gradD[9*(size_t)P + q]; GN[q] = gradD[9*(size_t)N + q]; }
            vdotT(kN, GN, a3);
            vdotT(kP, GP, b3);
#pragma unroll
            for (int d = 0; d < 3; d++) ssf[d] = mul(sub(a3[d], b3[d]), dc);
#pragma unroll
            for (int d = 0; d < 3; d++) t[d] = add(mul(dc, sub(D[3*(size_t)N + d], D[3*(size_t)P + d])), ssf[d]);
            double lim = dvd(mul(limitCoeff, mag3(t)), add(mul(sub(1.0, limitCoeff), mag3(ssf)), A_SMALL));
            if (lim > 1.0) lim = 1.0;
#pragma unroll
            for (int d = 0; d < 3; d++) sn[d] = add(mul(dc, sub(D[3*(size_t)N + d], D[3*(size_t)P + d])), mul(ssf[d], lim));
        }
        else patch_snGrad(M, f, D, gradD, fixedValue, patchGradient, sn);
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) gradDf[9*(size_t)f + 3*a + b] = add(G[3*a + b], mul(n[a], sn[b]));
    }
}

// ---- outer-iteration algebra of evolve(): D.relax() and residual() --------------------------------------------------
// D = prevIter + alpha*(D - prevIter)  [EXT GeometricField::relax]; per-block max of mag(D - Dold)   (max is exact)
__global__ void __launch_bounds__(ATPB) k_relax_maxdd(int nCells, double alpha, int doRelax, double* __restrict__ D,
                                                      const double* __restrict__ Dprev, const double* __restrict__ Dold,
                                                      double* __restrict__ blockMax)
{
    __shared__ double sh[ATPB];
    double mx = -1.0e300;
    for (int c = blockIdx.x*ATPB + threadIdx.x; c < nCells; c += gridDim.x*ATPB)
    {
        double d[3];
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            double v = D[3*(size_t)c + k];
            if (doRelax)
            {
                const double pv = Dprev[3*(size_t)c + k];
                v = add(pv, mul(alpha, sub(v, pv)));
                D[3*(size_t)c + k] = v;
            }
            d[k] = Dold ? sub(v, Dold[3*(size_t)c + k]) : v;      // Dold == nullptr: max mag(D) (the incremental model's residual)
        }
        mx = fmax(mx, mag3(d));
    }
    sh[threadIdx.x] = mx;
    __syncthreads();
    for (int o = ATPB/2; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]); __syncthreads(); }
    if (threadIdx.x == 0) blockMax[blockIdx.x] = sh[0];
}

// res = max_c mag(D - Dprev)/(maxDD + SMALL)   (unsTotalLagrangianStressSolve.C:647-672)
__global__ void __launch_bounds__(ATPB) k_residual(int nCells, int nBlocksPrev, const double* __restrict__ D,
                                                   const double* __restrict__ Dprev, const double* __restrict__ blockMaxDD,
                                                   double* __restrict__ blockRes)
{
    __shared__ double sh[ATPB];
    double maxDD = -1.0e300;
    for (int i = 0; i < nBlocksPrev; i++) maxDD = fmax(maxDD, blockMaxDD[i]);
    const double den = add(maxDD, A_SMALL);
    double mx = -1.0e300;
    for (int c = blockIdx.x*ATPB + threadIdx.x; c < nCells; c += gridDim.x*ATPB)
    {
        double d[3];
#pragma unroll
        for (int k = 0; k < 3; k++) d[k] = sub(D[3*(size_t)c + k], Dprev[3*(size_t)c + k]);
        mx = fmax(mx, dvd(mag3(d), den));
    }
    sh[threadIdx.x] = mx;
    __syncthreads();
    for (int o = ATPB/2; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]); __syncthreads(); }
    if (threadIdx.x == 0) blockRes[blockIdx.x] = sh[0];
}

} // namespace

// ------------------------------------------------------------------------------------------------------------------
struct gpupcg_mesh_s
{
    int device = 0, numSMs = 0;
    MeshDev M{};
    std::vector<void*> owned;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    // work arrays
    double *dMu = nullptr, *dLambda = nullptr, *dRho = nullptr, *dD = nullptr, *dGradD = nullptr, *dGradDf = nullptr;
    double *dTraction = nullptr, *dPressure = nullptr, *dFixed = nullptr;
    double *dUpper = nullptr, *dLup = nullptr, *dGms = nullptr, *dFluxE = nullptr, *dFluxC = nullptr;
    double *dDiag = nullptr, *dSource = nullptr, *dIC = nullptr, *dBC = nullptr, *dEps = nullptr, *dSig = nullptr;
    double *dPointD = nullptr, *dGradOld = nullptr, *dPatchGrad = nullptr, *dGradDb = nullptr;
    double *dDprev = nullptr, *dDold = nullptr, *dRed = nullptr;
    // round 2: the remaining state of evolve() (allocated on first use)
    double *dDoldold = nullptr, *dThreeK = nullptr, *dAlpha = nullptr, *dDT = nullptr, *dDTb = nullptr;
    double *dGradDfBStart = nullptr;        // incremental loop: boundary rows of gradDf at the start of the last outer iteration
    double *dDooo = nullptr, *dDoooo = nullptr, *dTimeA = nullptr;      // `backward` d2dt2: 3rd / 4th old level, per-cell time terms
    double *dDb = nullptr, *dDbPrev = nullptr, *dU = nullptr;
    double *dFluxX = nullptr, *dFluxY = nullptr, *dFluxT = nullptr;
    bool haveDold = false, haveDoldold = false, haveThermal = false, haveDb = false;
    LsqDev lsq{};                                  // leastSquaresVolPointInterpolation data (gpupcg_mesh_set_lsq)
    PointBcDev pointBc{};                          // point patch fields of pointD (gpupcg_mesh_set_point_constraints)
    double* hScal = nullptr;                       // pinned: {residual, maxDD, minDet, maxDet}
    // chunked assembly: the per-face fluxes of one chunk live in a ring sized to stay in L2 between the face and the cell kernel
    std::vector<AsmChunk> chunks;
    const int *dExtraFace = nullptr, *dCellSlot = nullptr;
    int ringSlots = 0;
    // two chunks in flight: the (latency-bound) cell kernel of chunk i overlaps the (bandwidth-bound) face kernel of
    // chunk i+1 on a second stream, each parity with its own flux ring
    const int *dCellBfStart = nullptr, *dCellBf = nullptr;   // cell -> its boundary faces (index - nIF), ascending = patch order
    bool assembled = false;                                   // the device holds an assembled D equation
    bool haveGradD = false, haveGradDf = false;     // the device holds grad(D) / the face gradient (residency across calls)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr, evFaces[2] = {nullptr, nullptr};
    double* ring2 = nullptr;
};

static thread_local std::string g_meshErr;

#define MCK(call)                                                                                   \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char b_[400];                            \
         snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
         m->err = b_; g_meshErr = b_; return GPUPCG_ERR_CUDA; } } while (0)

template <class T>
static cudaError_t up(gpupcg_mesh m, const T*& d, const T* h, size_t n)
{
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1)*sizeof(T));
    if (e != cudaSuccess) return e;
    m->owned.push_back(p);
    d = (const T*)p;
    if (n) e = cudaMemcpyAsync(p, h, n*sizeof(T), cudaMemcpyHostToDevice, m->stream);
    return e;
}

template <class T>
static cudaError_t alloc(gpupcg_mesh m, T*& d, size_t n)
{
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1)*sizeof(T));
    if (e == cudaSuccess) { m->owned.push_back(p); d = (T*)p; }
    return e;
}

extern "C" const char* gpupcg_mesh_last_error(gpupcg_mesh m) { return m ? m->err.c_str() : g_meshErr.c_str(); }

extern "C" int gpupcg_mesh_destroy(gpupcg_mesh m)
{
    if (!m) return GPUPCG_ERR_ARG;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    for (void* p : m->owned) cudaFree(p);
    if (m->evFork) cudaEventDestroy(m->evFork);
    if (m->evJoin) cudaEventDestroy(m->evJoin);
    for (auto& e : m->evFaces) if (e) cudaEventDestroy(e);
    if (m->stream2) cudaStreamDestroy(m->stream2);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return GPUPCG_OK;
}

extern "C" int gpupcg_mesh_create(gpupcg_mesh* out, int device, int nCells, int nFaces, int nInternalFaces,
                                  const int* owner, const int* neighbour, int nPatches, const int* patchStart,
                                  const int* patchSize, const int* patchType, const double* Sf, const double* magSf,
                                  const double* Cf, const double* C, const double* V, const double* weights,
                                  const double* deltaCoeffs)
{
    if (!out || nCells < 0 || nFaces < nInternalFaces || !owner || (nInternalFaces > 0 && !neighbour)) return GPUPCG_ERR_ARG;
    *out = nullptr;
    if (!Sf || !magSf || !Cf || !C || !V || !deltaCoeffs || (nInternalFaces > 0 && !weights) || nPatches < 0 ||
        (nPatches > 0 && (!patchStart || !patchSize || !patchType)))
    {
        g_meshErr = "gpupcg_mesh_create: NULL geometry / patch table";
        return GPUPCG_ERR_ARG;
    }
    for (int p = 0; p < nPatches; p++)
    {
        if (patchSize[p] < 0 || patchStart[p] < nInternalFaces || (long long)patchStart[p] + patchSize[p] > nFaces)
        {
            g_meshErr = "gpupcg_mesh_create: patch " + std::to_string(p) + " lies outside the boundary faces [nInternalFaces, nFaces)";
            return GPUPCG_ERR_ARG;
        }
        if (patchType[p] < 0 || patchType[p] > 5)
        {
            // anything else (processor, cyclic, solidTraction, ...) would silently get tractionDisplacement coefficients
            g_meshErr = "gpupcg_mesh_create: patch " + std::to_string(p) + " has patchType " + std::to_string(patchType[p]) +
                        " (supported: 0 tractionDisplacement, 1 fixedDisplacement, 2 symmetryDisplacement, 3 symmetryPlane, 4 empty, 5 fixedValue)";
            return GPUPCG_ERR_UNSUPPORTED;
        }
    }
    for (int f = 0; f < nFaces; f++)
        if (owner[f] < 0 || owner[f] >= nCells || (f < nInternalFaces && (neighbour[f] < 0 || neighbour[f] >= nCells)))
        {
            g_meshErr = "gpupcg_mesh_create: owner/neighbour label out of range at face " + std::to_string(f);
            return GPUPCG_ERR_ARG;
        }
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0 || device >= nDev)
    {
        g_meshErr = "gpupcg_mesh_create: no CUDA device visible (this library has no CPU fallback)";
        return GPUPCG_ERR_NODEVICE;
    }
    gpupcg_mesh m = new gpupcg_mesh_s();
    m->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
    {
        g_meshErr = "gpupcg_mesh_create: not a Blackwell (sm_100) class GPU";
        delete m;
        return GPUPCG_ERR_NODEVICE;
    }
    m->numSMs = prop.multiProcessorCount;
    MCK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    MCK(cudaEventCreate(&m->ev0)); MCK(cudaEventCreate(&m->ev1));
    const int nIF = nInternalFaces, nBF = nFaces - nIF;
    // cell -> faces in the order surfaceIntegrate / negSumDiag touch a cell: internal faces ascending, then boundary
    std::vector<int> start(nCells + 1, 0);
    for (int f = 0; f < nFaces; f++) start[owner[f] + 1]++;
    for (int f = 0; f < nIF; f++) start[neighbour[f] + 1]++;
    for (int c = 0; c < nCells; c++) start[c + 1] += start[c];
    std::vector<int> cf(start[nCells]), cur(start.begin(), start.end() - 1);
    for (int f = 0; f < nIF; f++)
    {
        cf[cur[owner[f]]++] = f;
        cf[cur[neighbour[f]]++] = f | 0x80000000;
    }
    for (int f = nIF; f < nFaces; f++) cf[cur[owner[f]]++] = f;
    // chunks of consecutive cells; ring slots of a chunk = owned internal faces, halo faces, boundary faces.
    // Default: ONE chunk.  Chunks small enough for the flux ring to stay in L2 between the face and the cell kernel
    // were measured (GPUPCG_ASM_CHUNK_CELLS, with and without a persisting-L2 access window on the ring) and lose to
    // the extra launches and tails at 1 M and 8 M cells: profiles/r1c_assembly_gradients.md.
    std::vector<int> extraFace, cellSlot(cf.size(), 0);
    {
        const char* ev = getenv("GPUPCG_ASM_CHUNK_CELLS");
        const int chunkCells = std::max(32, ev && *ev ? atoi(ev) : 0x3fffffff);
        std::vector<int> ownStart(nCells + 1, 0);                  // first internal face owned by cell c
        for (int f = 0; f < nIF; f++) ownStart[owner[f] + 1]++;
        for (int c = 0; c < nCells; c++) ownStart[c + 1] += ownStart[c];
        const int nChunks = (nCells + chunkCells - 1)/chunkCells;
        std::vector<std::vector<int> > halo(nChunks), bnd(nChunks);
        for (int f = 0; f < nIF; f++)
        {
            const int ko = owner[f]/chunkCells, kn = neighbour[f]/chunkCells;
            if (kn != ko) halo[kn].push_back(f);
        }
        for (int f = nIF; f < nFaces; f++) bnd[owner[f]/chunkCells].push_back(f);
        std::vector<int> slotOf(nFaces, -1);                        // scratch: slot of face f inside the chunk being numbered
        for (int k = 0; k < nChunks; k++)
        {
            AsmChunk K;
            K.c0 = k*chunkCells; K.c1 = std::min(nCells, K.c0 + chunkCells);
            K.f0 = ownStart[K.c0]; K.nOwn = ownStart[K.c1] - K.f0;
            K.extraStart = (int)extraFace.size();
            for (int f : halo[k]) { slotOf[f] = K.nOwn + (int)extraFace.size() - K.extraStart; extraFace.push_back(f); }
            for (int f : bnd[k]) { slotOf[f] = K.nOwn + (int)extraFace.size() - K.extraStart; extraFace.push_back(f); }
            K.nExtra = (int)extraFace.size() - K.extraStart;
            m->ringSlots = std::max(m->ringSlots, K.nOwn + K.nExtra);
            for (int c = K.c0; c < K.c1; c++)
                for (int e = start[c]; e < start[c + 1]; e++)
                {
                    const int f = cf[e] & 0x7fffffff;
                    const int sl = (f < nIF && f >= K.f0 && f < K.f0 + K.nOwn) ? f - K.f0 : slotOf[f];
                    cellSlot[e] = sl | (cf[e] & 0x80000000) | (f >= nIF ? 0x40000000 : 0);
                }
            m->chunks.push_back(K);
        }
        if (m->ringSlots >= 0x40000000) { g_meshErr = "gpupcg_mesh_create: assembly chunk too large"; delete m; return GPUPCG_ERR_ARG; }
    }
    std::vector<int> bft(nBF > 0 ? nBF : 1, 0);
    for (int p = 0; p < nPatches; p++)
        for (int i = 0; i < patchSize[p]; i++) bft[patchStart[p] + i - nIF] = patchType[p];
    MeshDev& M = m->M;
    M.nCells = nCells; M.nFaces = nFaces; M.nIF = nIF; M.nPatches = nPatches;
    MCK(up(m, M.owner, owner, nFaces)); MCK(up(m, M.neighbour, neighbour, nIF)); MCK(up(m, M.bfPatchType, bft.data(), nBF));
    MCK(up(m, M.cellFaceStart, start.data(), nCells + 1)); MCK(up(m, M.cellFace, cf.data(), cf.size()));
    MCK(up(m, m->dExtraFace, extraFace.data(), extraFace.size())); MCK(up(m, m->dCellSlot, cellSlot.data(), cellSlot.size()));
    {
        // cell -> boundary faces, in the order fvMatrix::addBoundaryDiag / addBoundarySource reach a cell (patches in
        // index order, faces in patch order = ascending face index)
        std::vector<int> bs(nCells + 1, 0), bfs(std::max(1, nBF));
        for (int f = nIF; f < nFaces; f++) bs[owner[f] + 1]++;
        for (int c = 0; c < nCells; c++) bs[c + 1] += bs[c];
        std::vector<int> bc(bs.begin(), bs.end() - 1);
        for (int f = nIF; f < nFaces; f++) bfs[bc[owner[f]]++] = f - nIF;
        MCK(up(m, m->dCellBfStart, bs.data(), bs.size())); MCK(up(m, m->dCellBf, bfs.data(), bfs.size()));
    }
    MCK(up(m, M.Sf, Sf, 3*(size_t)nFaces)); MCK(up(m, M.magSf, magSf, nFaces)); MCK(up(m, M.Cf, Cf, 3*(size_t)nFaces));
    MCK(up(m, M.C, C, 3*(size_t)nCells)); MCK(up(m, M.V, V, nCells)); MCK(up(m, M.weights, weights, nIF));
    MCK(up(m, M.deltaCoeffs, deltaCoeffs, nFaces));
    MCK(alloc(m, m->dMu, nCells)); MCK(alloc(m, m->dLambda, nCells)); MCK(alloc(m, m->dRho, nCells));
    MCK(alloc(m, m->dD, 3*(size_t)nCells)); MCK(alloc(m, m->dGradD, 9*(size_t)nCells)); MCK(alloc(m, m->dGradDf, 9*(size_t)nFaces));
    MCK(alloc(m, m->dTraction, 3*(size_t)nBF)); MCK(alloc(m, m->dPressure, nBF)); MCK(alloc(m, m->dFixed, 3*(size_t)nBF));
    MCK(alloc(m, m->dUpper, nIF)); MCK(alloc(m, m->dGms, nFaces));
    {
        // one allocation for the ring {fluxE, fluxC, Lup}: 56 B per slot
        double* ring = nullptr;
        MCK(alloc(m, ring, 7*(size_t)std::max(1, m->ringSlots)));
        m->dFluxE = ring; m->dFluxC = ring + 3*(size_t)m->ringSlots; m->dLup = ring + 6*(size_t)m->ringSlots;
        if (m->chunks.size() > 1)
        {
            MCK(alloc(m, m->ring2, 7*(size_t)std::max(1, m->ringSlots)));
            MCK(cudaStreamCreateWithFlags(&m->stream2, cudaStreamNonBlocking));
            MCK(cudaEventCreateWithFlags(&m->evFork, cudaEventDisableTiming));
            MCK(cudaEventCreateWithFlags(&m->evJoin, cudaEventDisableTiming));
            for (auto& e : m->evFaces) MCK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
    }
    MCK(alloc(m, m->dDiag, nCells)); MCK(alloc(m, m->dSource, 3*(size_t)nCells));
    MCK(alloc(m, m->dIC, 3*(size_t)nBF)); MCK(alloc(m, m->dBC, 3*(size_t)nBF));
    MCK(alloc(m, m->dPatchGrad, 3*(size_t)nBF));
    MCK(cudaMemsetAsync(m->dPatchGrad, 0, sizeof(double)*3*(size_t)std::max(1, nBF), m->stream));
    MCK(alloc(m, m->dDprev, 3*(size_t)nCells)); MCK(alloc(m, m->dDold, 3*(size_t)nCells)); MCK(alloc(m, m->dRed, 2*1024));
    MCK(alloc(m, m->dEps, 6*(size_t)nCells)); MCK(alloc(m, m->dSig, 6*(size_t)nCells));
    MCK(cudaStreamSynchronize(m->stream));
    *out = m;
    return GPUPCG_OK;
}

static int agrid(gpupcg_mesh m, long long n)
{
    const char* v = getenv("GPUPCG_ASM_GRID_MULT");         // CTAs per SM of the grid-stride kernels (0: one element per thread)
    const int mult = (v && *v) ? atoi(v) : 24;      // 24: 1.778 ms per assembly at 8 M cells against 1.796 with 16 (tools/assembly_grid.py)
    long long g = (n + ATPB - 1)/ATPB;
    if (mult > 0) g = std::min<long long>(g, (long long)m->numSMs*mult);
    return (int)std::max<long long>(1, g);
}

extern "C" int gpupcg_assemble_D(gpupcg_mesh m, const double* mu, const double* lambda, const double* rho, const double* g,
                                 const double* D, const double* gradD, const double* gradDf, const double* traction,
                                 const double* pressure, const double* fixedValue, double limitCoeff, int reps,
                                 double* upper, double* diag, double* source, double* internalCoeffs,
                                 double* boundaryCoeffs, double* kernel_ms)
{
    if (!m || !mu || !lambda || !rho || !g || !D) return GPUPCG_ERR_ARG;
    if ((!gradD && !m->haveGradD) || (!gradDf && !m->haveGradDf))
    {
        m->err = "gpupcg_assemble_D: gradD / gradDf == NULL but the device holds no gradients yet (gpupcg_gradients first)";
        return GPUPCG_ERR_STATE;
    }
    const MeshDev& M = m->M;
    const int nC = M.nCells, nF = M.nFaces, nIF = M.nIF, nBF = nF - nIF;
    MCK(cudaSetDevice(m->device));
    if (gradD) m->haveGradD = true;
    if (gradDf) m->haveGradDf = true;
    cudaStream_t s = m->stream;
    auto h2d = [&](double* d, const double* h, size_t n) { return n && h ? cudaMemcpyAsync(d, h, n*sizeof(double), cudaMemcpyHostToDevice, s) : cudaSuccess; };
    MCK(h2d(m->dMu, mu, nC)); MCK(h2d(m->dLambda, lambda, nC)); MCK(h2d(m->dRho, rho, nC));
    MCK(h2d(m->dD, D, 3*(size_t)nC)); MCK(h2d(m->dGradD, gradD, 9*(size_t)nC)); MCK(h2d(m->dGradDf, gradDf, 9*(size_t)nF));
    MCK(h2d(m->dTraction, traction, 3*(size_t)nBF)); MCK(h2d(m->dPressure, pressure, nBF)); MCK(h2d(m->dFixed, fixedValue, 3*(size_t)nBF));
    if (reps < 1) reps = 1;
    MCK(cudaEventRecord(m->ev0, s));
    for (int r = 0; r < reps; r++)
    {
        const char* ov = getenv("GPUPCG_ASM_OVERLAP");
        const bool overlap = m->stream2 && !(ov && *ov == '0');
        if (overlap) { MCK(cudaEventRecord(m->evFork, s)); MCK(cudaStreamWaitEvent(m->stream2, m->evFork, 0)); }
        int ci = 0;
        for (const AsmChunk& K : m->chunks)
        {
            // (reading -upper instead of a separate Lup array in the cell kernel was measured: 3 % slower at 8 M cells)
            const bool odd = overlap && (ci & 1);
            cudaStream_t cs = odd ? m->stream2 : s;
            double* ring = odd ? m->ring2 : m->dFluxE;
            double *fluxE = ring, *fluxC = ring + 3*(size_t)m->ringSlots, *lup = ring + 6*(size_t)m->ringSlots;
            // the face kernels run one after the other (they are bandwidth-bound); what overlaps is the cell kernel of
            // chunk i with the face kernel of chunk i+1
            if (overlap && ci > 0) { MCK(cudaStreamWaitEvent(cs, m->evFaces[(ci - 1) & 1], 0)); }
            k_asm_faces<<<agrid(m, K.nOwn + K.nExtra), ATPB, 0, cs>>>(M, K, m->dExtraFace, m->dMu, m->dLambda, m->dD, m->dGradD, m->dGradDf,
                                                                      limitCoeff, m->dUpper, lup, m->dGms, fluxE, fluxC);
            if (overlap) { MCK(cudaEventRecord(m->evFaces[ci & 1], cs)); }
            k_asm_cells<<<agrid(m, K.c1 - K.c0), ATPB, 0, cs>>>(M, K, AsmCtl{}, nullptr, nullptr, m->dCellSlot, m->dRho, g[0], g[1], g[2],
                                                                lup, m->dUpper, fluxE, fluxC, m->dDiag, m->dSource);
            ci++;
        }
        if (overlap) { MCK(cudaEventRecord(m->evJoin, m->stream2)); MCK(cudaStreamWaitEvent(s, m->evJoin, 0)); }
        if (nBF > 0)
            k_asm_patches<<<agrid(m, nBF), ATPB, 0, s>>>(M, AsmCtl{}, nullptr, nullptr, nullptr, m->dMu, m->dLambda, m->dD, m->dGradD,
                                                         m->dGradDf, m->dGms, m->dTraction, m->dPressure, m->dFixed, m->dIC, m->dBC,
                                                         m->dPatchGrad);
    }
    MCK(cudaEventRecord(m->ev1, s));
    MCK(cudaGetLastError());
    m->assembled = true;
    auto d2h = [&](double* h, const double* d, size_t n) { return n && h ? cudaMemcpyAsync(h, d, n*sizeof(double), cudaMemcpyDeviceToHost, s) : cudaSuccess; };
    MCK(d2h(upper, m->dUpper, nIF)); MCK(d2h(diag, m->dDiag, nC)); MCK(d2h(source, m->dSource, 3*(size_t)nC));
    MCK(d2h(internalCoeffs, m->dIC, 3*(size_t)nBF)); MCK(d2h(boundaryCoeffs, m->dBC, 3*(size_t)nBF));
    MCK(cudaStreamSynchronize(s));
    if (kernel_ms) { float ms = 0.f; cudaEventElapsedTime(&ms, m->ev0, m->ev1); *kernel_ms = ms/reps; }
    return GPUPCG_OK;
}

extern "C" int gpupcg_sigma(gpupcg_mesh m, const double* mu, const double* lambda, const double* gradD,
                            double* epsilon, double* sigma)
{
    if (!m || !mu || !lambda || !epsilon || !sigma) return GPUPCG_ERR_ARG;
    if (!gradD && !m->haveGradD) { m->err = "gpupcg_sigma: gradD == NULL but the device holds no grad(D) yet"; return GPUPCG_ERR_STATE; }
    const int nC = m->M.nCells;
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    MCK(cudaMemcpyAsync(m->dMu, mu, sizeof(double)*nC, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(m->dLambda, lambda, sizeof(double)*nC, cudaMemcpyHostToDevice, s));
    if (gradD) { MCK(cudaMemcpyAsync(m->dGradD, gradD, sizeof(double)*9*(size_t)nC, cudaMemcpyHostToDevice, s)); m->haveGradD = true; }
    k_sigma<<<agrid(m, nC), ATPB, 0, s>>>(nC, m->dMu, m->dLambda, m->dGradD, m->dEps, m->dSig);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(epsilon, m->dEps, sizeof(double)*6*(size_t)nC, cudaMemcpyDeviceToHost, s));
    MCK(cudaMemcpyAsync(sigma, m->dSig, sizeof(double)*6*(size_t)nC, cudaMemcpyDeviceToHost, s));
    MCK(cudaStreamSynchronize(s));
    return GPUPCG_OK;
}


// gradient() of the tractionDisplacement patches as the last gpupcg_assemble_D left it (0 on other patches)
extern "C" int gpupcg_patch_gradient(gpupcg_mesh m, double* patchGradient)
{
    if (!m || !patchGradient) return GPUPCG_ERR_ARG;
    const int nBF = m->M.nFaces - m->M.nIF;
    MCK(cudaSetDevice(m->device));
    MCK(cudaMemcpyAsync(patchGradient, m->dPatchGrad, sizeof(double)*3*(size_t)nBF, cudaMemcpyDeviceToHost, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    return GPUPCG_OK;
}

// D.relax(alpha) (relax != 0) followed by unsTotalLagrangianStress::residual()
extern "C" int gpupcg_relax_residual(gpupcg_mesh m, double alpha, int relax, double* D, const double* Dprev, const double* Dold,
                                     double* residual)
{
    if (!m || !D || !Dprev || !Dold || !residual) return GPUPCG_ERR_ARG;
    const int nC = m->M.nCells;
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    MCK(cudaMemcpyAsync(m->dD, D, sizeof(double)*3*(size_t)nC, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(m->dDprev, Dprev, sizeof(double)*3*(size_t)nC, cudaMemcpyHostToDevice, s));
    MCK(cudaMemcpyAsync(m->dDold, Dold, sizeof(double)*3*(size_t)nC, cudaMemcpyHostToDevice, s));
    const int nb = std::min(1024, agrid(m, nC));
    k_relax_maxdd<<<nb, ATPB, 0, s>>>(nC, alpha, relax, m->dD, m->dDprev, m->dDold, m->dRed);
    k_residual<<<nb, ATPB, 0, s>>>(nC, nb, m->dD, m->dDprev, m->dRed, m->dRed + 1024);
    MCK(cudaGetLastError());
    std::vector<double> part(nb);
    if (relax) MCK(cudaMemcpyAsync(D, m->dD, sizeof(double)*3*(size_t)nC, cudaMemcpyDeviceToHost, s));
    MCK(cudaMemcpyAsync(part.data(), m->dRed + 1024, sizeof(double)*nb, cudaMemcpyDeviceToHost, s));
    MCK(cudaStreamSynchronize(s));
    double r = -1.0e300;
    for (int i = 0; i < nb; i++) r = std::max(r, part[i]);
    *residual = r;
    return GPUPCG_OK;
}

extern "C" int gpupcg_mesh_set_points(gpupcg_mesh m, int nPoints, const double* points, const int* faceStart, const int* facePoints)
{
    if (!m || nPoints < 0 || !points || !faceStart || !facePoints) return GPUPCG_ERR_ARG;
    MCK(cudaSetDevice(m->device));
    MeshDev& M = m->M;
    M.nPoints = nPoints;
    MCK(up(m, M.points, points, 3*(size_t)nPoints));
    MCK(up(m, M.faceStart, faceStart, (size_t)M.nFaces + 1));
    MCK(up(m, M.facePoints, facePoints, (size_t)faceStart[M.nFaces]));
    const int nBF = M.nFaces - M.nIF;
    MCK(alloc(m, m->dPointD, 3*(size_t)nPoints)); MCK(alloc(m, m->dGradOld, 9*(size_t)M.nCells));
    MCK(alloc(m, m->dGradDb, 9*(size_t)nBF));
    {
        double *fanS = nullptr, *fanV = nullptr, *gradV = nullptr;
        const size_t nSlots = (size_t)faceStart[M.nFaces];
        MCK(alloc(m, fanS, 3*nSlots)); MCK(alloc(m, fanV, nSlots)); MCK(alloc(m, gradV, (size_t)M.nCells));
        M.fanS = fanS; M.fanV = fanV; M.gradV = gradV;
        k_fan_geometry<<<agrid(m, M.nFaces), ATPB, 0, m->stream>>>(M, fanS, fanV);
        k_cell_fan_volume<<<agrid(m, M.nCells), ATPB, 0, m->stream>>>(M, fanV, gradV);
        MCK(cudaGetLastError());
    }
    MCK(cudaStreamSynchronize(m->stream));
    return GPUPCG_OK;
}

// fvc::grad(D, pointD) then fvc::fGrad(D, pointD) exactly as evolve() calls them (unsTotalLagrangianStress.C:925-926):
// the face gradient's snGrad correction and fixedDisplacement::snGrad see the NEW cell gradient.
extern "C" int gpupcg_gradients(gpupcg_mesh m, const double* D, const double* pointD, const double* gradDold,
                                const double* fixedValue, const double* patchGradient, double limitCoeff, int reps,
                                double* gradD, double* gradDb, double* gradDf, double* kernel_ms)
{
    if (!m || !D || !pointD) return GPUPCG_ERR_ARG;
    if (!m->M.points) { m->err = "gpupcg_gradients: gpupcg_mesh_set_points has not been called"; return GPUPCG_ERR_STATE; }
    if (!gradDold && !m->haveGradD)
    {
        m->err = "gpupcg_gradients: gradDold == NULL but the device holds no grad(D) yet";
        return GPUPCG_ERR_STATE;
    }
    const MeshDev& M = m->M;
    const int nC = M.nCells, nF = M.nFaces, nBF = nF - M.nIF;
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    auto h2d = [&](double* d, const double* h, size_t n) { return n && h ? cudaMemcpyAsync(d, h, n*sizeof(double), cudaMemcpyHostToDevice, s) : cudaSuccess; };
    MCK(h2d(m->dD, D, 3*(size_t)nC)); MCK(h2d(m->dPointD, pointD, 3*(size_t)M.nPoints));
    if (gradDold) { MCK(h2d(m->dGradOld, gradDold, 9*(size_t)nC)); }
    else { MCK(cudaMemcpyAsync(m->dGradOld, m->dGradD, sizeof(double)*9*(size_t)nC, cudaMemcpyDeviceToDevice, s)); }   // the registered grad(D) at call time
    MCK(h2d(m->dFixed, fixedValue, 3*(size_t)nBF)); MCK(h2d(m->dPatchGrad, patchGradient, 3*(size_t)nBF));
    if (reps < 1) reps = 1;
    MCK(cudaEventRecord(m->ev0, s));
    for (int r = 0; r < reps; r++)
    {
        k_grad_cells<<<std::max(1, std::min((nC + GTPB - 1)/GTPB, m->numSMs*32)), GTPB, 0, s>>>(M, m->dPointD, m->dGradD);
        if (nBF > 0)
            k_grad_patches<<<agrid(m, nBF), ATPB, 0, s>>>(M, m->dD, m->dPointD, m->dGradOld, m->dFixed, m->dPatchGrad, m->dGradDb);
        k_fgrad_faces<<<agrid(m, nF), ATPB, 0, s>>>(M, m->dD, m->dPointD, m->dGradD, m->dFixed, m->dPatchGrad, limitCoeff, m->dGradDf);
    }
    MCK(cudaEventRecord(m->ev1, s));
    MCK(cudaGetLastError());
    auto d2h = [&](double* h, const double* d, size_t n) { return n && h ? cudaMemcpyAsync(h, d, n*sizeof(double), cudaMemcpyDeviceToHost, s) : cudaSuccess; };
    MCK(d2h(gradD, m->dGradD, 9*(size_t)nC)); MCK(d2h(gradDb, m->dGradDb, 9*(size_t)nBF)); MCK(d2h(gradDf, m->dGradDf, 9*(size_t)nF));
    MCK(cudaStreamSynchronize(s));
    m->haveGradD = m->haveGradDf = true;
    if (kernel_ms) { float ms = 0.f; cudaEventElapsedTime(&ms, m->ev0, m->ev1); *kernel_ms = ms/reps; }
    return GPUPCG_OK;
}

// internal (csrc/mesh_view.h): what the solver side needs of an assembled D equation; the mesh's stream is idle on return
extern "C" int gpupcgi_mesh_view(gpupcg_mesh m, gpupcg_mesh_view* v)
{
    if (!m || !v) return GPUPCG_ERR_ARG;
    if (!m->assembled) { m->err = "no assembled D equation on the device (gpupcg_assemble_D first)"; return GPUPCG_ERR_STATE; }
    MCK(cudaSetDevice(m->device));
    MCK(cudaStreamSynchronize(m->stream));
    v->device = m->device; v->nCells = m->M.nCells; v->nIF = m->M.nIF; v->nBF = m->M.nFaces - m->M.nIF;
    v->upper = m->dUpper; v->diag = m->dDiag; v->source = m->dSource; v->ic = m->dIC; v->bc = m->dBC;
    v->cellBfStart = m->dCellBfStart; v->cellBf = m->dCellBf;
    return GPUPCG_OK;
}

// a device-resident field on demand: which 0 D [nCells*3], 1 gradD [nCells*9], 2 gradDf [nFaces*9], 3 gradDb [nBoundaryFaces*9]
extern "C" int gpupcg_mesh_fetch(gpupcg_mesh m, int which, double* out)
{
    if (!m || !out) return GPUPCG_ERR_ARG;
    const MeshDev& M = m->M;
    const double* src = nullptr; size_t n = 0;
    switch (which)
    {
        case 0: src = m->dD; n = 3*(size_t)M.nCells; break;
        case 1: src = m->dGradD; n = 9*(size_t)M.nCells; break;
        case 2: src = m->dGradDf; n = 9*(size_t)M.nFaces; break;
        case 3: src = m->dGradDb; n = 9*(size_t)(M.nFaces - M.nIF); break;
        default: m->err = "gpupcg_mesh_fetch: unknown field id"; return GPUPCG_ERR_ARG;
    }
    if ((which == 1 && !m->haveGradD) || (which >= 2 && !m->haveGradDf) || !src)
    {
        m->err = "gpupcg_mesh_fetch: the device does not hold that field yet";
        return GPUPCG_ERR_STATE;
    }
    MCK(cudaSetDevice(m->device));
    if (n) MCK(cudaMemcpyAsync(out, src, n*sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    return GPUPCG_OK;
}

// ====================================================================================================================
// Round 2: the whole outer loop of evolve() on resident fields (no per-iteration PCIe traffic)
// ====================================================================================================================
namespace
{
struct FieldRef { double** p; size_t n; bool* have; };

FieldRef field_ref(gpupcg_mesh m, int id)
{
    const MeshDev& M = m->M;
    const size_t nC = (size_t)M.nCells, nF = (size_t)M.nFaces, nIF = (size_t)M.nIF, nBF = nF - nIF, nP = (size_t)M.nPoints;
    switch (id)
    {
        case GPUPCG_F_MU: return {&m->dMu, nC, nullptr};
        case GPUPCG_F_LAMBDA: return {&m->dLambda, nC, nullptr};
        case GPUPCG_F_RHO: return {&m->dRho, nC, nullptr};
        case GPUPCG_F_D: return {&m->dD, 3*nC, nullptr};
        case GPUPCG_F_DOLD: return {&m->dDold, 3*nC, &m->haveDold};
        case GPUPCG_F_DOLDOLD: return {&m->dDoldold, 3*nC, &m->haveDoldold};
        case GPUPCG_F_GRADD: return {&m->dGradD, 9*nC, &m->haveGradD};
        case GPUPCG_F_GRADDF: return {&m->dGradDf, 9*nF, &m->haveGradDf};
        case GPUPCG_F_TRACTION: return {&m->dTraction, 3*nBF, nullptr};
        case GPUPCG_F_PRESSURE: return {&m->dPressure, nBF, nullptr};
        case GPUPCG_F_FIXEDVALUE: return {&m->dFixed, 3*nBF, nullptr};
        case GPUPCG_F_THREEK: return {&m->dThreeK, nC, &m->haveThermal};
        case GPUPCG_F_ALPHA: return {&m->dAlpha, nC, nullptr};
        case GPUPCG_F_DT: return {&m->dDT, nC, nullptr};
        case GPUPCG_F_DTB: return {&m->dDTb, nBF, nullptr};
        case GPUPCG_F_POINTD: return {&m->dPointD, 3*nP, nullptr};
        case GPUPCG_F_DB: return {&m->dDb, 3*nBF, &m->haveDb};
        case GPUPCG_F_GRADDB: return {&m->dGradDb, 9*nBF, nullptr};
        case GPUPCG_F_U: return {&m->dU, 3*nC, nullptr};
        case GPUPCG_F_EPSILON: return {&m->dEps, 6*nC, nullptr};
        case GPUPCG_F_SIGMA: return {&m->dSig, 6*nC, nullptr};
        case GPUPCG_F_UPPER: return {&m->dUpper, nIF, nullptr};
        case GPUPCG_F_DIAG: return {&m->dDiag, nC, nullptr};
        case GPUPCG_F_SOURCE: return {&m->dSource, 3*nC, nullptr};
        case GPUPCG_F_INTERNALCOEFFS: return {&m->dIC, 3*nBF, nullptr};
        case GPUPCG_F_BOUNDARYCOEFFS: return {&m->dBC, 3*nBF, nullptr};
        case GPUPCG_F_PATCHGRADIENT: return {&m->dPatchGrad, 3*nBF, nullptr};
        case GPUPCG_F_DPREV: return {&m->dDprev, 3*nC, nullptr};
        case GPUPCG_F_DBPREV: return {&m->dDbPrev, 3*nBF, nullptr};
        case GPUPCG_F_DOOO: return {&m->dDooo, 3*nC, nullptr};
        case GPUPCG_F_DOOOO: return {&m->dDoooo, 3*nC, nullptr};
        case GPUPCG_F_GRADDFB_START: return {&m->dGradDfBStart, 9*nBF, nullptr};
        default: return {nullptr, 0, nullptr};
    }
}

// a field that is allocated on first use (zero-filled)
int ensure(gpupcg_mesh m, double*& p, size_t n)
{
    if (p) return GPUPCG_OK;
    MCK(alloc(m, p, n));
    MCK(cudaMemsetAsync(p, 0, sizeof(double)*std::max<size_t>(n, 1), m->stream));
    return GPUPCG_OK;
}

// host-side scalars of one use of a backward scheme, IEEE double, in the reference's order (backwardD2dt2Scheme.C:251-253)
BwW backward_weights(double deltaT, double e)
{
    BwW w;
    w.c = 1 + deltaT/(deltaT + e);
    w.c00 = deltaT*deltaT/(e*(deltaT + e));
    w.c0 = w.c + w.c00;
    return w;
}

AsmCtl make_ctl(const gpupcg_eqn_controls* c)
{
    AsmCtl T{};
    if (!c) return T;
    T.d2dt2 = c->d2dt2Scheme; T.ddt = c->ddtScheme; T.K = c->K; T.damping = c->K > A_SMALL;
    T.nonLinear = c->nonLinear != 0; T.thermal = c->thermal != 0;
    if (T.d2dt2 == 1)
    {
        // host-side scalars of [EXT] EulerD2dt2Scheme::fvmD2dt2, IEEE double, in its order
        T.coefft = (c->deltaT + c->deltaT0)/(2*c->deltaT);
        T.coefft00 = (c->deltaT + c->deltaT0)/(2*c->deltaT0);
        const double s = c->deltaT + c->deltaT0;
        T.rDeltaT2 = 4.0/(s*s);
    }
    if (T.ddt == 1 || T.ddt == 2) T.rDeltaT = 1.0/c->deltaT;
    if (T.ddt == 2)
    {
        const BwW w = backward_weights(c->deltaT, c->eDamp);
        T.cbD = w.c; T.cb0D = w.c0; T.cb00D = w.c00;
    }
    return T;
}
} // namespace

extern "C" int gpupcg_mesh_field_size(gpupcg_mesh m, int id, long long* n)
{
    if (!m || !n) return GPUPCG_ERR_ARG;
    FieldRef r = field_ref(m, id);
    if (!r.p) { m->err = "gpupcg_mesh_field_size: unknown field id"; return GPUPCG_ERR_ARG; }
    *n = (long long)r.n;
    return GPUPCG_OK;
}

extern "C" int gpupcg_mesh_set_field(gpupcg_mesh m, int id, const double* host)
{
    if (!m || !host) return GPUPCG_ERR_ARG;
    FieldRef r = field_ref(m, id);
    if (!r.p) { m->err = "gpupcg_mesh_set_field: unknown field id"; return GPUPCG_ERR_ARG; }
    if (id == GPUPCG_F_POINTD && !m->M.points) { m->err = "gpupcg_mesh_set_field: gpupcg_mesh_set_points first"; return GPUPCG_ERR_STATE; }
    MCK(cudaSetDevice(m->device));
    int rc = ensure(m, *r.p, r.n); if (rc) return rc;
    if (r.n) MCK(cudaMemcpyAsync(*r.p, host, sizeof(double)*r.n, cudaMemcpyHostToDevice, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    if (r.have) *r.have = true;
    return GPUPCG_OK;
}

extern "C" int gpupcg_mesh_get_field(gpupcg_mesh m, int id, double* host)
{
    if (!m || !host) return GPUPCG_ERR_ARG;
    FieldRef r = field_ref(m, id);
    if (!r.p) { m->err = "gpupcg_mesh_get_field: unknown field id"; return GPUPCG_ERR_ARG; }
    if (!*r.p) { m->err = "gpupcg_mesh_get_field: the device does not hold that field yet"; return GPUPCG_ERR_STATE; }
    MCK(cudaSetDevice(m->device));
    if (r.n) MCK(cudaMemcpyAsync(host, *r.p, sizeof(double)*r.n, cudaMemcpyDeviceToHost, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    return GPUPCG_OK;
}

// leastSquaresVolPointInterpolation data of the mesh: what the reference's volToPoint_ object holds
// (pointCells / pointBndFaces stencils, origins(), invLsMatrices(), mirror planes), flattened.
extern "C" int gpupcg_mesh_set_lsq(gpupcg_mesh m, const int* pointStart, const int* entry, const double* invLs,
                                   const double* origin, const unsigned char* mirror)
{
    if (!m || !pointStart || !entry || !invLs || !origin || !mirror) return GPUPCG_ERR_ARG;
    if (!m->M.points) { m->err = "gpupcg_mesh_set_lsq: gpupcg_mesh_set_points has not been called"; return GPUPCG_ERR_STATE; }
    const int nP = m->M.nPoints, nC = m->M.nCells, nBF = m->M.nFaces - m->M.nIF;
    const int rows = pointStart[nP];
    for (int p = 0; p < nP; p++)
        if (pointStart[p + 1] < pointStart[p] || (mirror[p] && ((pointStart[p + 1] - pointStart[p]) & 1)))
        { m->err = "gpupcg_mesh_set_lsq: pointStart not ascending / odd row count at a mirrored point"; return GPUPCG_ERR_ARG; }
    for (int r = 0; r < rows; r++)
        if (entry[r] >= nC || -1 - entry[r] >= nBF) { m->err = "gpupcg_mesh_set_lsq: stencil entry out of range"; return GPUPCG_ERR_ARG; }
    MCK(cudaSetDevice(m->device));
    LsqDev& L = m->lsq;
    L.nPoints = nP;
    MCK(up(m, L.start, pointStart, (size_t)nP + 1)); MCK(up(m, L.entry, entry, (size_t)rows));
    MCK(up(m, L.inv, invLs, 3*(size_t)rows)); MCK(up(m, L.origin, origin, 3*(size_t)nP));
    MCK(up(m, L.mirror, mirror, (size_t)nP));
    MCK(cudaStreamSynchronize(m->stream));
    return GPUPCG_OK;
}

// The point patch fields of pointD that act in pf.correctBoundaryConditions() at the end of interpolate()
// (leastSquaresVolPointInterpolate.C:301): symmetryPlane (kind 1, vec = the patch's point normal) and fixedValue (kind 2).
extern "C" int gpupcg_mesh_set_point_constraints(gpupcg_mesh m, int nConstrained, const int* point, const int* start,
                                                 const int* kind, const double* vec)
{
    if (!m || nConstrained < 0 || (nConstrained > 0 && (!point || !start || !kind || !vec))) return GPUPCG_ERR_ARG;
    if (!m->M.points) { m->err = "gpupcg_mesh_set_point_constraints: gpupcg_mesh_set_points has not been called"; return GPUPCG_ERR_STATE; }
    PointBcDev& B = m->pointBc;
    B.n = 0;
    if (nConstrained == 0) return GPUPCG_OK;
    if (start[0] != 0) { m->err = "gpupcg_mesh_set_point_constraints: start[0] != 0"; return GPUPCG_ERR_ARG; }
    for (int i = 0; i < nConstrained; i++)
    {
        if (point[i] < 0 || point[i] >= m->M.nPoints || start[i + 1] < start[i] || (i > 0 && point[i] <= point[i - 1]))
        { m->err = "gpupcg_mesh_set_point_constraints: points must be distinct, ascending and in range, start ascending"; return GPUPCG_ERR_ARG; }
    }
    const int rows = start[nConstrained];
    for (int r = 0; r < rows; r++)
        if (kind[r] != 1 && kind[r] != 2)
        { m->err = "gpupcg_mesh_set_point_constraints: kind must be 1 (symmetryPlane) or 2 (fixedValue)"; return GPUPCG_ERR_UNSUPPORTED; }
    MCK(cudaSetDevice(m->device));
    MCK(up(m, B.point, point, (size_t)nConstrained)); MCK(up(m, B.start, start, (size_t)nConstrained + 1));
    MCK(up(m, B.kind, kind, (size_t)rows)); MCK(up(m, B.vec, vec, 3*(size_t)rows));
    MCK(cudaStreamSynchronize(m->stream));
    B.n = nConstrained;
    return GPUPCG_OK;
}

// The other scalar systems of the boundary (TEqn, pEqn): assembly_scalar.cuh.  Host arrays in, host arrays out (all optional
// outputs); the geometry, addressing and patch table are those of the mesh handle (D's patch types only tell `empty`).
extern "C" int gpupcg_assemble_scalar(gpupcg_mesh m, const gpupcg_scalar_eqn* e, double* upper, double* diag, double* source,
                                      double* internalCoeffs, double* boundaryCoeffs, double* kernel_ms)
{
    if (!m || !e) return GPUPCG_ERR_ARG;
    const MeshDev& M = m->M;
    const int nC = M.nCells, nF = M.nFaces, nIF = M.nIF, nBF = nF - nIF;
    if ((e->gammaCells == nullptr) == (e->gammaFaces == nullptr))
    { m->err = "gpupcg_assemble_scalar: give gammaCells or gammaFaces (exactly one)"; return GPUPCG_ERR_ARG; }
    if (e->sign != 1 && e->sign != -1) { m->err = "gpupcg_assemble_scalar: sign must be +1 (ddt - laplacian) or -1 (laplacian == div(phi))"; return GPUPCG_ERR_ARG; }
    if (e->ddtScheme != 0 && e->ddtScheme != 1) { m->err = "gpupcg_assemble_scalar: ddt scheme 0 steadyState or 1 Euler"; return GPUPCG_ERR_UNSUPPORTED; }
    if (e->ddtScheme == 1 && (!(e->deltaT > 0.0) || !e->rate || !e->psiOld || e->sign != 1))
    { m->err = "gpupcg_assemble_scalar: Euler ddt needs deltaT > 0, rate, psiOld and the (ddt - laplacian) form"; return GPUPCG_ERR_ARG; }
    if (nBF > 0 && (!e->bfKind || !e->bfValue)) { m->err = "gpupcg_assemble_scalar: NULL boundary tables"; return GPUPCG_ERR_ARG; }
    if (e->refCell >= nC) { m->err = "gpupcg_assemble_scalar: reference cell out of range"; return GPUPCG_ERR_ARG; }
    if (e->psi && (!e->lsP || !e->lsN || (nBF > 0 && !e->lsB)))
    { m->err = "gpupcg_assemble_scalar: the skewCorrected correction needs the least-squares vectors lsP, lsN, lsB"; return GPUPCG_ERR_ARG; }
    for (int bf = 0; bf < nBF; bf++)
        if (e->bfKind[bf] < 0 || e->bfKind[bf] > 1)
        { m->err = "gpupcg_assemble_scalar: boundary kind must be 0 (zeroGradient) or 1 (fixedValue)"; return GPUPCG_ERR_UNSUPPORTED; }
    MCK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    double *dGamma = nullptr, *dGms = nullptr, *dLup = nullptr, *dUp = nullptr, *dDg = nullptr, *dSrc = nullptr, *dIc = nullptr, *dBc = nullptr;
    double *dRate = nullptr, *dOld = nullptr, *dPhi = nullptr, *dVal = nullptr; int* dKind = nullptr;
    std::vector<void*> tmp;
    auto dev = [&](void** p, size_t bytes, const void* host) -> cudaError_t
    {
        cudaError_t er = cudaMalloc(p, std::max<size_t>(bytes, 8));
        if (er != cudaSuccess) return er;
        tmp.push_back(*p);
        return host ? cudaMemcpyAsync(*p, host, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
    };
    auto freeAll = [&]() { for (void* p : tmp) cudaFree(p); };
#define SCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { freeAll(); m->err = std::string("gpupcg_assemble_scalar: ") + cudaGetErrorString(e_); return GPUPCG_ERR_CUDA; } } while (0)
    const bool onFaces = e->gammaFaces != nullptr;
    SCK(dev((void**)&dGamma, sizeof(double)*(onFaces ? nF : nC), onFaces ? e->gammaFaces : e->gammaCells));
    SCK(dev((void**)&dGms, sizeof(double)*nF, nullptr)); SCK(dev((void**)&dLup, sizeof(double)*nIF, nullptr));
    SCK(dev((void**)&dUp, sizeof(double)*nIF, nullptr)); SCK(dev((void**)&dDg, sizeof(double)*nC, nullptr));
    SCK(dev((void**)&dSrc, sizeof(double)*nC, nullptr)); SCK(dev((void**)&dIc, sizeof(double)*nBF, nullptr));
    SCK(dev((void**)&dBc, sizeof(double)*nBF, nullptr));
    SCK(dev((void**)&dKind, sizeof(int)*nBF, e->bfKind)); SCK(dev((void**)&dVal, sizeof(double)*nBF, e->bfValue));
    if (e->ddtScheme == 1) { SCK(dev((void**)&dRate, sizeof(double)*nC, e->rate)); SCK(dev((void**)&dOld, sizeof(double)*nC, e->psiOld)); }
    if (e->phi) SCK(dev((void**)&dPhi, sizeof(double)*nF, e->phi));
    double *dPsi = nullptr, *dLsP = nullptr, *dLsN = nullptr, *dLsB = nullptr, *dGrad = nullptr, *dCorr = nullptr;
    if (e->psi)
    {
        SCK(dev((void**)&dPsi, sizeof(double)*nC, e->psi));
        SCK(dev((void**)&dLsP, sizeof(double)*3*nIF, e->lsP)); SCK(dev((void**)&dLsN, sizeof(double)*3*nIF, e->lsN));
        SCK(dev((void**)&dLsB, sizeof(double)*3*nBF, nBF ? e->lsB : nullptr));
        SCK(dev((void**)&dGrad, sizeof(double)*3*nC, nullptr)); SCK(dev((void**)&dCorr, sizeof(double)*nIF, nullptr));
    }
    const ScalarEqnDev E{e->sign, e->ddtScheme, e->ddtScheme == 1 ? 1.0/e->deltaT : 0.0, onFaces ? 1 : 0};
    SCK(cudaEventRecord(m->ev0, s));
    k_sc_faces<<<agrid(m, nF), ATPB, 0, s>>>(M, E, dGamma, dGms, dLup, dUp);
    if (e->psi)
    {
        k_sc_lsgrad<<<agrid(m, nC), ATPB, 0, s>>>(M, dPsi, dKind, dVal, dLsP, dLsN, dLsB, dGrad);
        if (nIF > 0) k_sc_corr_faces<<<agrid(m, nIF), ATPB, 0, s>>>(M, dGms, dPsi, dGrad, e->limitCoeff, dCorr);
    }
    k_sc_cells<<<agrid(m, nC), ATPB, 0, s>>>(M, E, dLup, dRate, dOld, dPhi, dDg, dSrc, dCorr);
    if (nBF > 0) k_sc_patches<<<agrid(m, nBF), ATPB, 0, s>>>(M, E, dKind, dVal, dGms, dIc, dBc);
    if (e->refCell >= 0) k_sc_reference<<<1, 32, 0, s>>>(e->refCell, e->refValue, dDg, dSrc);
    SCK(cudaEventRecord(m->ev1, s));
    SCK(cudaGetLastError());
    if (upper) SCK(cudaMemcpyAsync(upper, dUp, sizeof(double)*nIF, cudaMemcpyDeviceToHost, s));
    if (diag) SCK(cudaMemcpyAsync(diag, dDg, sizeof(double)*nC, cudaMemcpyDeviceToHost, s));
    if (source) SCK(cudaMemcpyAsync(source, dSrc, sizeof(double)*nC, cudaMemcpyDeviceToHost, s));
    if (internalCoeffs && nBF) SCK(cudaMemcpyAsync(internalCoeffs, dIc, sizeof(double)*nBF, cudaMemcpyDeviceToHost, s));
    if (boundaryCoeffs && nBF) SCK(cudaMemcpyAsync(boundaryCoeffs, dBc, sizeof(double)*nBF, cudaMemcpyDeviceToHost, s));
    SCK(cudaStreamSynchronize(s));
    if (kernel_ms) { float ms = 0.f; cudaEventElapsedTime(&ms, m->ev0, m->ev1); *kernel_ms = ms; }
    freeAll();
#undef SCK
    return GPUPCG_OK;
}

// runTime++ : D.storeOldTimes()  ->  D.oldTime().oldTime() = D.oldTime(); D.oldTime() = D
extern "C" int gpupcg_mesh_new_time_step(gpupcg_mesh m)
{
    if (!m) return GPUPCG_ERR_ARG;
    MCK(cudaSetDevice(m->device));
    const size_t n = 3*(size_t)m->M.nCells;
    int rc = ensure(m, m->dDoldold, n); if (rc) return rc;
    MCK(cudaMemcpyAsync(m->dDoldold, m->dDold, sizeof(double)*n, cudaMemcpyDeviceToDevice, m->stream));
    MCK(cudaMemcpyAsync(m->dDold, m->dD, sizeof(double)*n, cudaMemcpyDeviceToDevice, m->stream));
    m->haveDold = m->haveDoldold = true;
    MCK(cudaStreamSynchronize(m->stream));
    return GPUPCG_OK;
}

extern "C" int gpupcg_mesh_copy_field(gpupcg_mesh m, int dstId, int srcId)
{
    if (!m) return GPUPCG_ERR_ARG;
    MCK(cudaSetDevice(m->device));
    FieldRef d = field_ref(m, dstId), r = field_ref(m, srcId);
    if (!d.p || !r.p || d.n != r.n) { m->err = "gpupcg_mesh_copy_field: unknown field id or sizes differ"; return GPUPCG_ERR_ARG; }
    if (!*r.p) { m->err = "gpupcg_mesh_copy_field: the source field is not on the device"; return GPUPCG_ERR_STATE; }
    int rc = ensure(m, *d.p, d.n); if (rc) return rc;
    MCK(cudaMemcpyAsync(*d.p, *r.p, sizeof(double)*d.n, cudaMemcpyDeviceToDevice, m->stream));
    if (d.have) *d.have = true;
    MCK(cudaStreamSynchronize(m->stream));
    return GPUPCG_OK;
}

static int check_ctl(gpupcg_mesh m, const gpupcg_eqn_controls* c)
{
    if (!c) return GPUPCG_OK;
    if (c->d2dt2Scheme < 0 || c->d2dt2Scheme > 2 || c->ddtScheme < 0 || c->ddtScheme > 2)
    { m->err = "unsupported d2dt2 / ddt scheme (0 steadyState, 1 Euler, 2 backward)"; return GPUPCG_ERR_UNSUPPORTED; }
    if ((c->d2dt2Scheme >= 1 || c->ddtScheme >= 1) && !(c->deltaT > 0.0 && c->deltaT0 > 0.0))
    { m->err = "Euler / backward time schemes need deltaT > 0 and deltaT0 > 0"; return GPUPCG_ERR_ARG; }
    if (c->d2dt2Scheme == 2 && !(c->eD2 > 0.0 && c->eDdtD > 0.0 && c->eDdtD0 > 0.0 && c->eDdtD00 > 0.0))
    { m->err = "backward d2dt2 needs the effective old time steps eD2, eDdtD, eDdtD0, eDdtD00 (> 0; GREAT during start-up)"; return GPUPCG_ERR_ARG; }
    if (c->ddtScheme == 2 && !(c->eU > 0.0) && !(c->eDamp > 0.0))
    { m->err = "backward ddt needs the effective old time steps eU / eDamp (> 0; GREAT during start-up)"; return GPUPCG_ERR_ARG; }
    if (c->thermal && !m->haveThermal)
    { m->err = "thermal stress requested but threeK / alpha / DT / DTb have not been set (gpupcg_mesh_set_field)"; return GPUPCG_ERR_STATE; }
    return GPUPCG_OK;
}

// the D equation from the RESIDENT fields, every branch of unsTotalLagrangianStress.C:846-890; stream = m->stream
static int assemble_resident(gpupcg_mesh m, const gpupcg_eqn_controls* c, const double* g)
{
    const MeshDev& M = m->M;
    const int nC = M.nCells, nF = M.nFaces, nBF = nF - M.nIF;
    int rc = check_ctl(m, c); if (rc) return rc;
    if (!m->haveGradD || !m->haveGradDf) { m->err = "no grad(D) / gradDf on the device yet"; return GPUPCG_ERR_STATE; }
    const AsmCtl T = make_ctl(c);
    cudaStream_t s = m->stream;
    if (T.d2dt2 >= 1 || (T.damping && T.ddt >= 1))
    {
        rc = ensure(m, m->dDoldold, 3*(size_t)nC); if (rc) return rc;
    }
    if (T.d2dt2 == 2)
    {
        rc = ensure(m, m->dDooo, 3*(size_t)nC); if (rc) return rc;
        rc = ensure(m, m->dDoooo, 3*(size_t)nC); if (rc) return rc;
        rc = ensure(m, m->dTimeA, 4*(size_t)nC); if (rc) return rc;
        BackwardCtl W;
        W.rDt = 1.0/c->deltaT;
        W.d2 = backward_weights(c->deltaT, c->eD2); W.ddtD = backward_weights(c->deltaT, c->eDdtD);
        W.ddtD0 = backward_weights(c->deltaT, c->eDdtD0); W.ddtD00 = backward_weights(c->deltaT, c->eDdtD00);
        W.sc = W.d2.c*W.rDt;
        k_time_backward<<<agrid(m, nC), ATPB, 0, s>>>(nC, W, M.V, m->dRho, m->dDold, m->dDoldold, m->dDooo, m->dDoooo, m->dTimeA);
    }
    if (T.thermal) { rc = ensure(m, m->dAlpha, nC); if (rc) return rc; rc = ensure(m, m->dDT, nC); if (rc) return rc; rc = ensure(m, m->dDTb, nBF); if (rc) return rc; }
    const double limitCoeff = c ? c->limitCoeff : 1.0;
    if (m->chunks.size() != 1) { m->err = "resident assembly needs the one-chunk layout (unset GPUPCG_ASM_CHUNK_CELLS)"; return GPUPCG_ERR_UNSUPPORTED; }
    const AsmChunk& K = m->chunks[0];
    double *fluxE = m->dFluxE, *fluxC = m->dFluxE + 3*(size_t)m->ringSlots, *lup = m->dFluxE + 6*(size_t)m->ringSlots;
    k_asm_faces<<<agrid(m, K.nOwn + K.nExtra), ATPB, 0, s>>>(M, K, m->dExtraFace, m->dMu, m->dLambda, m->dD, m->dGradD, m->dGradDf,
                                                             limitCoeff, m->dUpper, lup, m->dGms, fluxE, fluxC);
    k_asm_cells<<<agrid(m, nC), ATPB, 0, s>>>(M, K, T, m->dDold, m->dDoldold, m->dCellSlot, m->dRho, g[0], g[1], g[2], lup, m->dUpper,
                                              fluxE, fluxC, m->dDiag, m->dSource, m->dTimeA);
    if (T.nonLinear || T.thermal)
    {
        if (T.nonLinear) { rc = ensure(m, m->dFluxX, 3*(size_t)nF); if (rc) return rc; rc = ensure(m, m->dFluxY, 3*(size_t)nF); if (rc) return rc; }
        if (T.thermal) { rc = ensure(m, m->dFluxT, 3*(size_t)nF); if (rc) return rc; }
        k_asm_faces_extra<<<agrid(m, nF), ATPB, 0, s>>>(M, T, m->dMu, m->dLambda, m->dGradDf, m->dThreeK, m->dAlpha, m->dDT, m->dDTb,
                                                        m->dFluxX, m->dFluxY, m->dFluxT);
        k_asm_cells_extra<<<agrid(m, nC), ATPB, 0, s>>>(M, T, m->dFluxX, m->dFluxY, m->dFluxT, m->dSource);
    }
    if (nBF > 0)
        k_asm_patches<<<agrid(m, nBF), ATPB, 0, s>>>(M, T, m->dThreeK, m->dAlpha, m->dDTb, m->dMu, m->dLambda, m->dD, m->dGradD, m->dGradDf,
                                                     m->dGms, m->dTraction, m->dPressure, m->dFixed, m->dIC, m->dBC, m->dPatchGrad);
    MCK(cudaGetLastError());
    m->assembled = true;
    return GPUPCG_OK;
}

extern "C" int gpupcg_assemble_D_resident(gpupcg_mesh m, const gpupcg_eqn_controls* ctl, const double* g)
{
    if (!m || !g) return GPUPCG_ERR_ARG;
    MCK(cudaSetDevice(m->device));
    int rc = assemble_resident(m, ctl, g); if (rc) return rc;
    MCK(cudaStreamSynchronize(m->stream));
    return GPUPCG_OK;
}

// D.correctBoundaryConditions() on the resident D: boundary values -> Db (slots of empty faces untouched)
static int evaluate_resident(gpupcg_mesh m)
{
    const MeshDev& M = m->M;
    const int nBF = M.nFaces - M.nIF;
    int rc = ensure(m, m->dDb, 3*(size_t)nBF); if (rc) return rc;
    if (nBF > 0) k_patch_evaluate<<<agrid(m, nBF), ATPB, 0, m->stream>>>(M, m->dD, m->dGradD, m->dFixed, m->dPatchGrad, m->dDb);
    MCK(cudaGetLastError());
    m->haveDb = true;
    return GPUPCG_OK;
}

extern "C" int gpupcg_patch_evaluate(gpupcg_mesh m)
{
    if (!m) return GPUPCG_ERR_ARG;
    MCK(cudaSetDevice(m->device));
    int rc = ensure(m, m->dGradD, 9*(size_t)m->M.nCells); if (rc) return rc;
    rc = evaluate_resident(m); if (rc) return rc;
    MCK(cudaStreamSynchronize(m->stream));
    return GPUPCG_OK;
}

// volToPoint_.interpolate(D_, pointD_) on the resident D and D.boundaryField()   (unsTotalLagrangianStress.C:924)
static int vol_to_point_resident(gpupcg_mesh m)
{
    if (!m->lsq.start) { m->err = "gpupcg_mesh_set_lsq has not been called"; return GPUPCG_ERR_STATE; }
    if (!m->haveDb) { m->err = "no boundary values of D on the device (gpupcg_patch_evaluate / set_field DB first)"; return GPUPCG_ERR_STATE; }
    k_lsq_interpolate<<<agrid(m, m->M.nPoints), ATPB, 0, m->stream>>>(m->lsq, m->M.points, m->dD, m->dDb, m->dPointD);
    MCK(cudaGetLastError());
    if (m->pointBc.n > 0)
    {
        k_point_constraints<<<agrid(m, m->pointBc.n), ATPB, 0, m->stream>>>(m->pointBc, m->dPointD);
        MCK(cudaGetLastError());
    }
    return GPUPCG_OK;
}

extern "C" int gpupcg_vol_to_point(gpupcg_mesh m, int reps, double* kernel_ms)
{
    if (!m) return GPUPCG_ERR_ARG;
    MCK(cudaSetDevice(m->device));
    if (reps < 1) reps = 1;
    MCK(cudaEventRecord(m->ev0, m->stream));
    for (int r = 0; r < reps; r++) { int rc = vol_to_point_resident(m); if (rc) return rc; }
    MCK(cudaEventRecord(m->ev1, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    if (kernel_ms) { float ms = 0.f; cudaEventElapsedTime(&ms, m->ev0, m->ev1); *kernel_ms = ms/reps; }
    return GPUPCG_OK;
}

static int gradients_resident(gpupcg_mesh m, double limitCoeff)
{
    const MeshDev& M = m->M;
    const int nC = M.nCells, nF = M.nFaces, nBF = nF - M.nIF;
    cudaStream_t s = m->stream;
    MCK(cudaMemcpyAsync(m->dGradOld, m->dGradD, sizeof(double)*9*(size_t)nC, cudaMemcpyDeviceToDevice, s));
    k_grad_cells<<<std::max(1, std::min((nC + GTPB - 1)/GTPB, m->numSMs*32)), GTPB, 0, s>>>(M, m->dPointD, m->dGradD);
    if (nBF > 0)
        k_grad_patches<<<agrid(m, nBF), ATPB, 0, s>>>(M, m->dD, m->dPointD, m->dGradOld, m->dFixed, m->dPatchGrad, m->dGradDb);
    k_fgrad_faces<<<agrid(m, nF), ATPB, 0, s>>>(M, m->dD, m->dPointD, m->dGradD, m->dFixed, m->dPatchGrad, limitCoeff, m->dGradDf);
    MCK(cudaGetLastError());
    m->haveGradD = m->haveGradDf = true;
    return GPUPCG_OK;
}

extern "C" int gpupcg_gradients_resident(gpupcg_mesh m, double limitCoeff)
{
    if (!m) return GPUPCG_ERR_ARG;
    if (!m->M.points) { m->err = "gpupcg_gradients_resident: gpupcg_mesh_set_points has not been called"; return GPUPCG_ERR_STATE; }
    MCK(cudaSetDevice(m->device));
    int rc = gradients_resident(m, limitCoeff); if (rc) return rc;
    MCK(cudaStreamSynchronize(m->stream));
    return GPUPCG_OK;
}

extern "C" int gpupcg_sigma_resident(gpupcg_mesh m, const gpupcg_eqn_controls* c)
{
    if (!m) return GPUPCG_ERR_ARG;
    if (!m->haveGradD) { m->err = "gpupcg_sigma_resident: no grad(D) on the device"; return GPUPCG_ERR_STATE; }
    MCK(cudaSetDevice(m->device));
    int rc = check_ctl(m, c); if (rc) return rc;
    const int nC = m->M.nCells;
    const AsmCtl T = make_ctl(c);
    cudaStream_t s = m->stream;
    rc = ensure(m, m->dU, 3*(size_t)nC); if (rc) return rc;
    // U_ = fvc::ddt(D_)   (:973)
    if (T.ddt == 2)
    {
        if (!(c->eU > 0.0)) { m->err = "gpupcg_sigma_resident: backward ddt needs eU > 0"; return GPUPCG_ERR_ARG; }
        rc = ensure(m, m->dDoldold, 3*(size_t)nC); if (rc) return rc;
        k_ddt_backward<<<agrid(m, 3*(long long)nC), ATPB, 0, s>>>(3*(long long)nC, T.rDeltaT, backward_weights(c->deltaT, c->eU),
                                                                 m->dD, m->dDold, m->dDoldold, m->dU);
    }
    else
    k_ddt<<<agrid(m, 3*(long long)nC), ATPB, 0, s>>>(3*(long long)nC, T.ddt, T.rDeltaT, m->dD, m->dDold, m->dU);
    // nonLinear here is the dictionary switch itself (:979), not `nonLinear && !enforceLinear`: the caller passes it in c->nonLinear
    k_sigma_ex<<<agrid(m, nC), ATPB, 0, s>>>(nC, T.nonLinear, T.thermal, m->dMu, m->dLambda, m->dGradD, m->dThreeK, m->dAlpha, m->dDT,
                                             m->dEps, m->dSig);
    MCK(cudaGetLastError());
    MCK(cudaStreamSynchronize(s));
    return GPUPCG_OK;
}

// ---- internal: the stages of one outer iteration, driven by gpupcg_evolve_D (gpupcg_vec.cuh) -------------------------
extern "C" int gpupcgi_mesh_bind_stream(gpupcg_mesh m, void* stream, void** previous)
{
    if (!m) return GPUPCG_ERR_ARG;
    if (previous) *previous = (void*)m->stream;
    m->stream = (cudaStream_t)stream;
    return GPUPCG_OK;
}

// D_.storePrevIter() (:844) incl. the boundary field, then the assembly
extern "C" int gpupcgi_mesh_outer_begin(gpupcg_mesh m, const gpupcg_eqn_controls* c, const double* g)
{
    const MeshDev& M = m->M;
    const size_t n = 3*(size_t)M.nCells, nb = 3*(size_t)(M.nFaces - M.nIF);
    int rc = ensure(m, m->dDb, nb); if (rc) return rc;
    rc = ensure(m, m->dDbPrev, nb); if (rc) return rc;
    if (!m->haveDb) { rc = evaluate_resident(m); if (rc) return rc; }
    MCK(cudaMemcpyAsync(m->dDprev, m->dD, sizeof(double)*n, cudaMemcpyDeviceToDevice, m->stream));
    if (nb) MCK(cudaMemcpyAsync(m->dDbPrev, m->dDb, sizeof(double)*nb, cudaMemcpyDeviceToDevice, m->stream));
    return assemble_resident(m, c, g);
}

// incremental loop: DSigmaf_ is formed from gradDDf_ at the start of an outer iteration (unsIncrTotalLagrangianStress.C:1013-1018)
// and not again after the last solve; keep the boundary rows it was formed from
extern "C" int gpupcgi_mesh_snapshot_gradDfB(gpupcg_mesh m)
{
    const MeshDev& M = m->M;
    const size_t nb9 = 9*(size_t)(M.nFaces - M.nIF);
    int rc = ensure(m, m->dGradDfBStart, nb9); if (rc) return rc;
    if (nb9) MCK(cudaMemcpyAsync(m->dGradDfBStart, m->dGradDf + 9*(size_t)M.nIF, sizeof(double)*nb9, cudaMemcpyDeviceToDevice, m->stream));
    return GPUPCG_OK;
}

// after DEqn.solve(): correctBoundaryConditions, D.relax() (:910), volToPoint (:924), grad / fGrad (:925-926), det bounds
// (:929-937), residual() (:949).  hostScal (pinned, 4 doubles) <- {residual, maxDD, minDet, maxDet}, asynchronously.
extern "C" int gpupcgi_mesh_outer_end(gpupcg_mesh m, const gpupcg_eqn_controls* c, double alpha, int doRelax, int detBounds,
                                      double* hostScal, int incremental)
{
    const MeshDev& M = m->M;
    const int nC = M.nCells;
    const long long nb3 = 3*(long long)(M.nFaces - M.nIF);
    cudaStream_t s = m->stream;
    int rc = evaluate_resident(m); if (rc) return rc;
    if (doRelax && nb3 > 0) k_relax_boundary<<<agrid(m, nb3), ATPB, 0, s>>>(nb3, alpha, m->dDb, m->dDbPrev);
    const int nb = std::min(1000, agrid(m, nC));        // partials: dRed[0, 1000) and dRed[1024, 2024); results at dRed[2040, 2048)
    k_relax_maxdd<<<nb, ATPB, 0, s>>>(nC, alpha, doRelax, m->dD, m->dDprev, incremental ? nullptr : m->dDold, m->dRed);
    k_residual<<<nb, ATPB, 0, s>>>(nC, nb, m->dD, m->dDprev, m->dRed, m->dRed + 1024);
    MCK(cudaGetLastError());
    rc = vol_to_point_resident(m); if (rc) return rc;
    rc = gradients_resident(m, c ? c->limitCoeff : 1.0); if (rc) return rc;
    double* out = m->dRed + 2048 - 8;       // the tail of the 2 x 1024 partials buffer is never reached (nb <= 1024 - 8 checked below)
    if (nb > 1016) { m->err = "internal: partials buffer too small"; return GPUPCG_ERR_STATE; }
    k_final_minmax<<<1, ATPB, 0, s>>>(nb, m->dRed + 1024, nullptr, out);            // out[0] = residual
    k_final_minmax<<<1, ATPB, 0, s>>>(nb, m->dRed, nullptr, out + 2);               // out[2] = maxDD
    if (detBounds)
    {
        const int nbd = std::min(1000, agrid(m, M.nFaces));
        k_det_minmax<<<nbd, ATPB, 0, s>>>(M, m->dGradDf, m->dRed, m->dRed + 1024);
        k_final_minmax<<<1, ATPB, 0, s>>>(nbd, m->dRed + 1024, m->dRed, out + 4);   // out[4] = maxDet, out[5] = minDet
    }
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(hostScal, out, sizeof(double)*6, cudaMemcpyDeviceToHost, s));
    return GPUPCG_OK;
}

extern "C" int gpupcgi_mesh_D(gpupcg_mesh m, double** D) { if (!m || !D) return GPUPCG_ERR_ARG; *D = m->dD; return GPUPCG_OK; }
extern "C" void gpupcgi_mesh_set_error(gpupcg_mesh m, const char* msg) { if (m) m->err = msg; }
