/*
 * oracle/fvm_oracle.c  --  CPU ORACLE of the D-equation ASSEMBLY, TEST INFRASTRUCTURE ONLY (see ldu_oracle.c).
 *
 * Restates, operand order included, what one outer iteration of
 *   unsTotalLagrangianStress::evolve()                       unsTotalLagrangianStress.C:846-859
 *     fvVectorMatrix DEqn( rho*fvm::d2dt2(D) == fvm::laplacian(2*muf+lambdaf, D, "laplacian(DD,D)")
 *                          + fvc::div(Sf & (-(muf+lambdaf)*gradDf + muf*gradDf.T() + lambdaf*(I*tr(gradDf)))) + rho*g )
 * builds for a steadyState d2dt2 scheme, linear elasticity (nonLinear no), no thermal stress:
 *   reference-owned pieces, followed line by line:
 *     constitutiveModel::mu / lambda                                         constitutiveModel.C:193-257
 *     skewCorrectedSnGrad<vector>::correction                                skewCorrectedVectorSnGrad.C:50-294
 *     tractionDisplacementFvPatchVectorField::updateCoeffs (linear branch)   tractionDisplacementFvPatchVectorField.C:148-257
 *     fixedDisplacementFvPatchVectorField::gradientBoundaryCoeffs            fixedDisplacementFvPatchVectorField.C:129-157
 *     sigma/epsilon                                                          unsTotalLagrangianStress.C:975-990
 *   foam-extend-3.1 pieces [EXT], restated from the published algorithm (PARITY UNPINNED):
 *     surfaceInterpolationScheme::interpolate (linear), gaussLaplacianScheme::fvmLaplacian, lduMatrix::negSumDiag,
 *     fvc::surfaceIntegrate, fvMatrix operator+/-/==, fixedValue / fixedGradient gradient coefficients.
 *
 * Field algebra is evaluated the way OpenFOAM's expression templates do: one elementwise temporary per operator,
 * left to right; vector & tensor = sum_i v_i T_ij accumulated x, y, z; tr(T) = xx + yy + zz; mag = sqrt(x*x+y*y+z*z).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_SMALL 1.0e-15

typedef struct orc_mesh
{
    int nCells, nFaces, nInternalFaces, nPatches;
    const int* owner;           /* [nFaces] */
    const int* neighbour;       /* [nInternalFaces] */
    const int* patchStart;      /* [nPatches] first face */
    const int* patchSize;       /* [nPatches] */
    const int* patchType;       /* 0 tractionDisplacement (fixedGradient), 1 fixedDisplacement (fixedValue), 2 symmetryDisplacement */
    const double* Sf;           /* [nFaces*3] */
    const double* magSf;        /* [nFaces] */
    const double* Cf;           /* [nFaces*3] */
    const double* C;            /* [nCells*3] */
    const double* V;            /* [nCells] */
    const double* weights;      /* [nInternalFaces] */
    const double* deltaCoeffs;  /* [nFaces] */
    /* point connectivity, only needed by the vertex-based gradients */
    int nPoints;
    const int* faceStart;       /* [nFaces+1] */
    const int* facePoints;
    const double* points;       /* [nPoints*3] */
} orc_mesh;

/* constitutiveModel::mu, lambda (constitutiveModel.C:193-257) */
void orc_lame(double E, double nu, int planeStress, double* mu, double* lambda)
{
    *mu = E/(2.0*(1.0 + nu));
    if (planeStress) *lambda = nu*E/((1.0 + nu)*(1.0 - nu));
    else *lambda = nu*E/((1.0 + nu)*(1.0 - 2.0*nu));
}

/* fvc::interpolate (linear): internal  w*(P - N) + N ; boundary = patch value (here: the owner cell's value,
 * the rheology law fields are uniform per material)                                                        */
void orc_interpolate(const orc_mesh* M, const double* vol, double* face)
{
    for (int f = 0; f < M->nInternalFaces; f++)
    {
        const double P = vol[M->owner[f]], N = vol[M->neighbour[f]];
        face[f] = M->weights[f]*(P - N) + N;
    }
    for (int f = M->nInternalFaces; f < M->nFaces; f++) face[f] = vol[M->owner[f]];
}

static inline double dot3(const double* a, const double* b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
static inline double mag3(const double* a) { return sqrt(a[0]*a[0] + a[1]*a[1] + a[2]*a[2]); }
/* vector & tensor (row-major T[9]): r_j = v.x*T.xj + v.y*T.yj + v.z*T.zj */
static inline void vdotT(const double* v, const double* T, double* r)
{
    r[0] = v[0]*T[0] + v[1]*T[3] + v[2]*T[6];
    r[1] = v[0]*T[1] + v[1]*T[4] + v[2]*T[7];
    r[2] = v[0]*T[2] + v[1]*T[5] + v[2]*T[8];
}

/* Sf & ( -(muf+lambdaf)*G + muf*G.T() + lambdaf*(I*tr(G)) )      unsTotalLagrangianStress.C:850-857 */
static void explicit_flux(const double* Sf, double muf, double lambdaf, const double* G, double* flux)
{
    const double a = -(muf + lambdaf);
    const double tr = G[0] + G[4] + G[8];
    const double sph = lambdaf*tr;
    double T[9];
    /* ((a*G) + (muf*G^T)) + sphericalTensor(sph) */
    T[0] = (a*G[0] + muf*G[0]) + sph;  T[1] = a*G[1] + muf*G[3];          T[2] = a*G[2] + muf*G[6];
    T[3] = a*G[3] + muf*G[1];          T[4] = (a*G[4] + muf*G[4]) + sph;  T[5] = a*G[5] + muf*G[7];
    T[6] = a*G[6] + muf*G[2];          T[7] = a*G[7] + muf*G[5];          T[8] = (a*G[8] + muf*G[8]) + sph;
    vdotT(Sf, T, flux);
}

/* skewCorrectedSnGrad<vector>::correction for one internal face (skewCorrectedVectorSnGrad.C:124-129,228-233,256-286) */
static void skew_correction(const orc_mesh* M, int f, const double* D, const double* gradD, double limitCoeff, double* ssf)
{
    const int P = M->owner[f], N = M->neighbour[f];
    const double* Sf = M->Sf + 3*f;
    const double* Cf = M->Cf + 3*f;
    const double magSqrSf = Sf[0]*Sf[0] + Sf[1]*Sf[1] + Sf[2]*Sf[2];
    double kP[3], kN[3], t[3], a[3], b[3];
    for (int d = 0; d < 3; d++) kP[d] = Cf[d] - M->C[3*P + d];
    {
        const double s = dot3(Sf, kP);
        for (int d = 0; d < 3; d++) kP[d] -= Sf[d]*s/magSqrSf;
    }
    for (int d = 0; d < 3; d++) kN[d] = Cf[d] - M->C[3*N + d];
    {
        const double s = dot3(Sf, kN);
        for (int d = 0; d < 3; d++) kN[d] -= Sf[d]*s/magSqrSf;
    }
    vdotT(kN, gradD + 9*N, a);
    vdotT(kP, gradD + 9*P, b);
    const double dc = M->deltaCoeffs[f];
    for (int d = 0; d < 3; d++) ssf[d] = (a[d] - b[d])*dc;
    /* limiter = min(limitCoeff*mag(orthSnGrad + ssf)/((1 - limitCoeff)*mag(ssf) + SMALL), 1) */
    for (int d = 0; d < 3; d++) t[d] = dc*(D[3*N + d] - D[3*P + d]) + ssf[d];
    double lim = limitCoeff*mag3(t)/((1 - limitCoeff)*mag3(ssf) + ORC_SMALL);
    if (lim > 1.0) lim = 1.0;
    for (int d = 0; d < 3; d++) ssf[d] *= lim;
}


/* symmetryDisplacementFvPatchVectorField::snGrad()   (symmetryDisplacementFvPatchVectorField.C:134-165):
 *   UP = D_P + (k & gradD_P);  snGrad = (transform(I - 2.0*sqr(nHat), UP) - UP)*(deltaCoeffs/2.0)              */
static void symmetry_snGrad(const orc_mesh* M, int f, const double* D, const double* gradD, double* sn)
{
    const int c = M->owner[f];
    double n[3], delta[3], k[3], kg[3], UP[3], S[6], T[6], tv[3];
    for (int d = 0; d < 3; d++) { n[d] = M->Sf[3*f + d]/M->magSf[f]; delta[d] = M->Cf[3*f + d] - M->C[3*c + d]; }
    const double nd = dot3(n, delta);
    for (int d = 0; d < 3; d++) k[d] = delta[d] - n[d]*nd;
    vdotT(k, gradD + 9*c, kg);
    for (int d = 0; d < 3; d++) UP[d] = D[3*c + d] + kg[d];
    /* sqr(n) as symmTensor xx xy xz yy yz zz ; T = I - 2.0*sqr(n) */
    S[0] = n[0]*n[0]; S[1] = n[0]*n[1]; S[2] = n[0]*n[2]; S[3] = n[1]*n[1]; S[4] = n[1]*n[2]; S[5] = n[2]*n[2];
    T[0] = 1.0 - 2.0*S[0]; T[1] = -(2.0*S[1]); T[2] = -(2.0*S[2]); T[3] = 1.0 - 2.0*S[3]; T[4] = -(2.0*S[4]); T[5] = 1.0 - 2.0*S[5];
    tv[0] = T[0]*UP[0] + T[1]*UP[1] + T[2]*UP[2];
    tv[1] = T[1]*UP[0] + T[3]*UP[1] + T[4]*UP[2];
    tv[2] = T[2]*UP[0] + T[4]*UP[1] + T[5]*UP[2];
    const double h = M->deltaCoeffs[f]/2.0;
    for (int d = 0; d < 3; d++) sn[d] = (tv[d] - UP[d])*h;
}

/*
 * Assemble the D equation.  Inputs per cell: mu, lambda, rho, D[3], gradD[9]; per face: gradDf[9] (internal +
 * boundary); per boundary face (index f - nInternalFaces): traction[3], pressure, fixedValue[3]; g[3].
 * Outputs: upper[nInternalFaces], diag[nCells] (before addBoundaryDiag), source[nCells*3] (before
 * addBoundarySource), internalCoeffs / boundaryCoeffs [nBoundaryFaces*3].
 */
void orc_assemble_D
(
    const orc_mesh* M, const double* mu, const double* lambda, const double* rho, const double* g,
    const double* D, const double* gradD, const double* gradDf,
    const double* traction, const double* pressure, const double* fixedValue, double limitCoeff,
    double* upper, double* diag, double* source, double* internalCoeffs, double* boundaryCoeffs,
    double* patchGradient       /* optional [nBoundaryFaces*3]: gradient() of the tractionDisplacement patches after updateCoeffs */
)
{
    const int nC = M->nCells, nF = M->nFaces, nIF = M->nInternalFaces;
    double* muf = (double*)malloc(sizeof(double)*nF);
    double* lambdaf = (double*)malloc(sizeof(double)*nF);
    double* gms = (double*)malloc(sizeof(double)*nF);           /* (2*muf + lambdaf)*magSf */
    double* fluxE = (double*)malloc(sizeof(double)*3*nF);
    double* fluxC = (double*)calloc(3*(size_t)(nIF > 0 ? nIF : 1), sizeof(double));
    double* sumE = (double*)calloc(3*(size_t)nC, sizeof(double));
    double* sumC = (double*)calloc(3*(size_t)nC, sizeof(double));
    double* Ldiag = (double*)calloc(nC, sizeof(double));

    orc_interpolate(M, mu, muf);
    orc_interpolate(M, lambda, lambdaf);
    for (int f = 0; f < nF; f++) gms[f] = (2*muf[f] + lambdaf[f])*M->magSf[f];

    /* gaussLaplacianScheme::fvmLaplacianUncorrected: upper = deltaCoeffs*gammaMagSf ; negSumDiag */
    for (int f = 0; f < nIF; f++)
    {
        const double up = M->deltaCoeffs[f]*gms[f];
        Ldiag[M->owner[f]] -= up;
        Ldiag[M->neighbour[f]] -= up;
        upper[f] = 0.0 - up;                        /* M0 - R : upper() -= B.upper() on a zero field */
    }
    for (int c = 0; c < nC; c++) diag[c] = 0.0 - Ldiag[c];

    /* corrected(): corrFlux = gammaMagSf*correction(D) ; explicit stress flux on every face */
    for (int f = 0; f < nIF; f++)
    {
        double ssf[3];
        skew_correction(M, f, D, gradD, limitCoeff, ssf);
        for (int d = 0; d < 3; d++) fluxC[3*f + d] = gms[f]*ssf[d];
    }
    for (int f = 0; f < nF; f++) explicit_flux(M->Sf + 3*f, muf[f], lambdaf[f], gradDf + 9*f, fluxE + 3*f);

    /* fvc::surfaceIntegrate: += owner, -= neighbour, boundary faces, then /V */
    for (int f = 0; f < nIF; f++)
    {
        for (int d = 0; d < 3; d++)
        {
            sumC[3*M->owner[f] + d] += fluxC[3*f + d];
            sumC[3*M->neighbour[f] + d] -= fluxC[3*f + d];
            sumE[3*M->owner[f] + d] += fluxE[3*f + d];
            sumE[3*M->neighbour[f] + d] -= fluxE[3*f + d];
        }
    }
    for (int f = nIF; f < nF; f++)
    {
        for (int d = 0; d < 3; d++) sumE[3*M->owner[f] + d] += fluxE[3*f + d];      /* corrFlux is zero on non-coupled patches */
    }
    for (int c = 0; c < nC; c++)
    {
        const double V = M->V[c];
        for (int d = 0; d < 3; d++)
        {
            const double dc = sumC[3*c + d]/V, de = sumE[3*c + d]/V;
            const double rg = rho[c]*g[d];
            /* laplacian: source -= V*div(corrFlux);  (L + div): source -= V*div;  (+ rho*g): source -= V*rho*g;
               M0 == R  ->  source = 0 - R.source                                                              */
            const double R = ((0.0 - V*dc) - V*de) - V*rg;
            source[3*c + d] = 0.0 - R;
        }
    }

    /* boundary coefficients */
    for (int p = 0; p < M->nPatches; p++)
    {
        for (int i = 0; i < M->patchSize[p]; i++)
        {
            const int f = M->patchStart[p] + i, bf = f - nIF, c = M->owner[f];
            const double* Sf = M->Sf + 3*f;
            const double magSf = M->magSf[f];
            const double dcf = M->deltaCoeffs[f];
            double n[3], gIC[3], gBC[3];
            for (int d = 0; d < 3; d++) n[d] = Sf[d]/magSf;
            if (M->patchType[p] == 1)
            {
                /* fixedValue: gradientInternalCoeffs = -pTraits::one*deltaCoeffs ;
                   fixedDisplacement::gradientBoundaryCoeffs = deltaCoeffs*(*this - (k & gradU.patchInternalField())) */
                double delta[3], k[3], kg[3];
                for (int d = 0; d < 3; d++) delta[d] = M->Cf[3*f + d] - M->C[3*c + d];
                const double nd = dot3(n, delta);
                for (int d = 0; d < 3; d++) k[d] = delta[d] - n[d]*nd;
                vdotT(k, gradD + 9*c, kg);
                for (int d = 0; d < 3; d++)
                {
                    gIC[d] = -dcf;
                    gBC[d] = dcf*(fixedValue[3*bf + d] - kg[d]);
                }
            }
            else if (M->patchType[p] == 2)
            {
                /* symmetryDisplacement ([EXT] transformFvPatchField over symmetryFvPatchField):
                   gradientInternalCoeffs = -deltaCoeffs*snGradTransformDiag(), snGradTransformDiag = (|nx|,|ny|,|nz|)
                   gradientBoundaryCoeffs = snGrad() - cmptMultiply(gradientInternalCoeffs, patchInternalField)     */
                double sn[3];
                symmetry_snGrad(M, f, D, gradD, sn);
                for (int d = 0; d < 3; d++)
                {
                    gIC[d] = -dcf*fabs(n[d]);
                    gBC[d] = sn[d] - gIC[d]*D[3*c + d];
                }
            }
            else
            {
                /* tractionDisplacement (fixedGradient): gradientInternalCoeffs = 0 ; gradientBoundaryCoeffs = gradient()
                   gradient = t - (n & (mu*gradDf.T() - (mu + lambda)*gradDf)) - n*lambda*tr(gradDf); /= (2.0*mu + lambda) */
                const double m = muf[f], la = lambdaf[f];
                const double* G = gradDf + 9*f;
                const double ml = m + la;
                double T[9], nT[3], t[3];
                T[0] = m*G[0] - ml*G[0]; T[1] = m*G[3] - ml*G[1]; T[2] = m*G[6] - ml*G[2];
                T[3] = m*G[1] - ml*G[3]; T[4] = m*G[4] - ml*G[4]; T[5] = m*G[7] - ml*G[5];
                T[6] = m*G[2] - ml*G[6]; T[7] = m*G[5] - ml*G[7]; T[8] = m*G[8] - ml*G[8];
                vdotT(n, T, nT);
                const double tr = G[0] + G[4] + G[8];
                const double den = 2.0*m + la;
                for (int d = 0; d < 3; d++)
                {
                    t[d] = traction[3*bf + d] - pressure[bf]*n[d];
                    gIC[d] = 0.0;
                    gBC[d] = ((t[d] - nT[d]) - n[d]*la*tr)/den;
                    if (patchGradient) patchGradient[3*bf + d] = gBC[d];
                }
            }
            for (int d = 0; d < 3; d++)
            {
                const double Lic = gms[f]*gIC[d];            /* internalCoeffs =  gammaMagSf*gradientInternalCoeffs  */
                const double Lbc = (-gms[f])*gBC[d];         /* boundaryCoeffs = -gammaMagSf*gradientBoundaryCoeffs  */
                internalCoeffs[3*bf + d] = 0.0 - Lic;
                boundaryCoeffs[3*bf + d] = 0.0 - Lbc;
            }
        }
    }
    free(muf); free(lambdaf); free(gms); free(fluxE); free(fluxC); free(sumE); free(sumC); free(Ldiag);
}

/* epsilon = symm(gradD); sigma = 2*mu*epsilon + I*(lambda*tr(epsilon))      unsTotalLagrangianStress.C:975-990
 * symmTensor order: xx xy xz yy yz zz                                                                       */
void orc_sigma(int nCells, const double* mu, const double* lambda, const double* gradD, double* epsilon, double* sigma)
{
    for (int c = 0; c < nCells; c++)
    {
        const double* G = gradD + 9*c;
        double e[6];
        e[0] = G[0]; e[1] = 0.5*(G[1] + G[3]); e[2] = 0.5*(G[2] + G[6]);
        e[3] = G[4]; e[4] = 0.5*(G[5] + G[7]); e[5] = G[8];
        const double tr = e[0] + e[3] + e[5];
        const double sph = lambda[c]*tr;
        const double tm = 2*mu[c];
        for (int k = 0; k < 6; k++) epsilon[6*c + k] = e[k];
        sigma[6*c + 0] = tm*e[0] + sph; sigma[6*c + 1] = tm*e[1]; sigma[6*c + 2] = tm*e[2];
        sigma[6*c + 3] = tm*e[3] + sph; sigma[6*c + 4] = tm*e[4]; sigma[6*c + 5] = tm*e[5] + sph;
    }
}


/* ======================================================================================================== */
/*  Vertex-based gradients of the reference: fvc::grad(vf, pf) and fvc::fGrad(vf, pf)                        */
/*  (src/fluidStructureInteraction/numerics/fvc/fvcGradf.C), followed line by line.                          */
/* ======================================================================================================== */
static inline void cross3(const double* a, const double* b, double* r)
{
    r[0] = a[1]*b[2] - a[2]*b[1];
    r[1] = a[2]*b[0] - a[0]*b[2];
    r[2] = a[0]*b[1] - a[1]*b[0];
}

/* accumulate one face's contribution  sign*(St*ttcf)  and  sign*(St & Ct)  (fvcGradf.C:541-610 / 623-684) */
static void grad_face(const orc_mesh* M, int f, const double* pf, double sign, double* G, double* V)
{
    const int* fp = M->facePoints + M->faceStart[f];
    const int n = M->faceStart[f + 1] - M->faceStart[f];
    const double* X = M->points;
    if (n == 3)
    {
        /* SF = curFace.normal(points)*curFace.average(points, pfI) ; SR = normal & centre   ([EXT] face.C, triangle) */
        const double *a = X + 3*fp[0], *b = X + 3*fp[1], *c = X + 3*fp[2];
        double ab[3], ac[3], nrm[3], avg[3], ctr[3];
        for (int d = 0; d < 3; d++) { ab[d] = b[d] - a[d]; ac[d] = c[d] - a[d]; }
        cross3(ab, ac, nrm);
        for (int d = 0; d < 3; d++)
        {
            nrm[d] = 0.5*nrm[d];
            avg[d] = (1.0/3.0)*(pf[3*fp[0] + d] + pf[3*fp[1] + d] + pf[3*fp[2] + d]);
            ctr[d] = (1.0/3.0)*(a[d] + b[d] + c[d]);
        }
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
            {
                const double t = nrm[i]*avg[j];
                if (sign > 0) G[3*i + j] += t; else G[3*i + j] -= t;
            }
        const double sr = dot3(nrm, ctr);
        if (sign > 0) *V += sr; else *V -= sr;
        return;
    }
    double centre[3] = {0, 0, 0}, cf[3] = {0, 0, 0};
    for (int p = 0; p < n; p++)
        for (int d = 0; d < 3; d++) { centre[d] += X[3*fp[p] + d]; cf[d] += pf[3*fp[p] + d]; }
    for (int d = 0; d < 3; d++) { centre[d] /= n; cf[d] /= n; }
    for (int p = 0; p < n; p++)
    {
        const int p0 = fp[p], p1 = fp[(p + 1) % n];
        double ttcf[3], a[3], b[3], St[3], Ct[3];
        for (int d = 0; d < 3; d++)
        {
            ttcf[d] = (pf[3*p0 + d] + pf[3*p1 + d] + cf[d])/3.0;
            a[d] = X[3*p0 + d] - centre[d];
            b[d] = X[3*p1 + d] - centre[d];
        }
        cross3(a, b, St);
        for (int d = 0; d < 3; d++)
        {
            St[d] /= 2.0;
            Ct[d] = (centre[d] + X[3*p0 + d] + X[3*p1 + d])/3;
        }
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
            {
                const double t = St[i]*ttcf[j];
                if (sign > 0) G[3*i + j] += t; else G[3*i + j] -= t;
            }
        const double sc = dot3(St, Ct);
        if (sign > 0) *V += sc; else *V -= sc;
    }
}

/* tangential face gradient from point values: sum_edges Le*fe / mag   (fvcGradf.C:97-142, 144-198, 410-482) */
static void face_tangential_grad(const orc_mesh* M, int f, const double* pf, double* G)
{
    const int* fp = M->facePoints + M->faceStart[f];
    const int n = M->faceStart[f + 1] - M->faceStart[f];
    const double* X = M->points;
    double nf[3];
    for (int d = 0; d < 3; d++) nf[d] = M->Sf[3*f + d]/M->magSf[f];
    for (int k = 0; k < 9; k++) G[k] = 0.0;
    for (int p = 0; p < n; p++)
    {
        const int s = fp[p], e_ = fp[(p + 1) % n];
        double e[3], Le[3], fe[3];
        for (int d = 0; d < 3; d++) e[d] = X[3*e_ + d] - X[3*s + d];
        const double ne = dot3(nf, e);
        for (int d = 0; d < 3; d++) e[d] -= nf[d]*ne;
        cross3(e, nf, Le);
        for (int d = 0; d < 3; d++) fe[d] = 0.5*(pf[3*s + d] + pf[3*e_ + d]);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) G[3*i + j] += Le[i]*fe[j];
    }
    for (int k = 0; k < 9; k++) G[k] /= M->magSf[f];
}

/* boundary snGrad of D as the patch field classes give it:
 *   fixedDisplacement::snGrad()  (fixedDisplacementFvPatchVectorField.C:96-127)
 *   fixedGradient::snGrad() = gradient()  (tractionDisplacement: the gradient set by the last updateCoeffs)   */
static void patch_snGrad(const orc_mesh* M, int ptype, int f, const double* D, const double* gradD,
                         const double* fixedValue, const double* patchGradient, double* sn)
{
    const int bf = f - M->nInternalFaces, c = M->owner[f];
    if (ptype == 1)
    {
        double n[3], delta[3], k[3], kg[3];
        for (int d = 0; d < 3; d++) { n[d] = M->Sf[3*f + d]/M->magSf[f]; delta[d] = M->Cf[3*f + d] - M->C[3*c + d]; }
        const double nd = dot3(n, delta);
        for (int d = 0; d < 3; d++) k[d] = delta[d] - n[d]*nd;
        vdotT(k, gradD + 9*c, kg);
        for (int d = 0; d < 3; d++) sn[d] = (fixedValue[3*bf + d] - (D[3*c + d] + kg[d]))*M->deltaCoeffs[f];
    }
    else if (ptype == 2) symmetry_snGrad(M, f, D, gradD, sn);
    else
    {
        for (int d = 0; d < 3; d++) sn[d] = patchGradient[3*bf + d];
    }
}

/* fvc::grad(D, pointD)  (fvcGradf.C:485-717).  gradDold = the registered grad(D) at call time (used by
 * fixedDisplacement::snGrad inside gaussGrad::correctBoundaryConditions); patchGradient = fixedGradient values.
 * Outputs: gradD[nCells*9], gradDb[nBoundaryFaces*9] (boundary field of the gradient).                        */
void orc_grad(const orc_mesh* M, const double* D, const double* pointD, const double* gradDold,
              const double* fixedValue, const double* patchGradient, double* gradD, double* gradDb)
{
    const int nC = M->nCells, nIF = M->nInternalFaces, nF = M->nFaces;
    double* V = (double*)calloc(nC, sizeof(double));
    memset(gradD, 0, sizeof(double)*9*(size_t)nC);
    for (int f = 0; f < nIF; f++)
    {
        /* the reference adds to the owner and subtracts from the neighbour triangle by triangle */
        grad_face(M, f, pointD, +1.0, gradD + 9*M->owner[f], V + M->owner[f]);
        grad_face(M, f, pointD, -1.0, gradD + 9*M->neighbour[f], V + M->neighbour[f]);
    }
    for (int f = nIF; f < nF; f++) grad_face(M, f, pointD, +1.0, gradD + 9*M->owner[f], V + M->owner[f]);
    for (int c = 0; c < nC; c++)
    {
        V[c] /= 3;
        for (int k = 0; k < 9; k++) gradD[9*c + k] /= V[c];
    }
    /* boundary: tangential gradient from the patch points, then gaussGrad::correctBoundaryConditions:
       patchGrad += n*(snGrad - (n & patchGrad))                                                             */
    for (int p = 0; p < M->nPatches; p++)
    {
        for (int i = 0; i < M->patchSize[p]; i++)
        {
            const int f = M->patchStart[p] + i, bf = f - nIF;
            double* G = gradDb + 9*bf;
            double n[3], sn[3], nG[3];
            face_tangential_grad(M, f, pointD, G);
            for (int d = 0; d < 3; d++) n[d] = M->Sf[3*f + d]/M->magSf[f];
            patch_snGrad(M, M->patchType[p], f, D, gradDold, fixedValue, patchGradient, sn);
            vdotT(n, G, nG);
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) G[3*a + b] += n[a]*(sn[b] - nG[b]);
        }
    }
    free(V);
}

/* fvc::fGrad(D, pointD)  (fvcGradf.C:47-231): tangential part from the points + n*snGrad(D), snGrad = the
 * skewCorrected scheme (orthogonal part + limited correction from the registered grad(D)) on internal faces
 * and the patch field's snGrad() on boundary faces.                                                           */
void orc_fgrad(const orc_mesh* M, const double* D, const double* pointD, const double* gradD,
               const double* fixedValue, const double* patchGradient, double limitCoeff, double* gradDf)
{
    const int nIF = M->nInternalFaces;
    for (int f = 0; f < M->nFaces; f++) face_tangential_grad(M, f, pointD, gradDf + 9*(size_t)f);
    for (int f = 0; f < nIF; f++)
    {
        double ssf[3], sn[3], n[3];
        skew_correction(M, f, D, gradD, limitCoeff, ssf);
        const int P = M->owner[f], N = M->neighbour[f];
        for (int d = 0; d < 3; d++)
        {
            sn[d] = M->deltaCoeffs[f]*(D[3*N + d] - D[3*P + d]) + ssf[d];
            n[d] = M->Sf[3*f + d]/M->magSf[f];
        }
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) gradDf[9*(size_t)f + 3*a + b] += n[a]*sn[b];
    }
    for (int p = 0; p < M->nPatches; p++)
    {
        for (int i = 0; i < M->patchSize[p]; i++)
        {
            const int f = M->patchStart[p] + i;
            double sn[3], n[3];
            patch_snGrad(M, M->patchType[p], f, D, gradD, fixedValue, patchGradient, sn);
            for (int d = 0; d < 3; d++) n[d] = M->Sf[3*f + d]/M->magSf[f];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) gradDf[9*(size_t)f + 3*a + b] += n[a]*sn[b];
        }
    }
}


/* GeometricField::relax(alpha) [EXT GeometricField.C]:  D = prevIter + alpha*(D - prevIter)   (unsTotalLagrangianStress.C:910)
 * then unsTotalLagrangianStress::residual()  (unsTotalLagrangianStressSolve.C:647-672):
 *   maxDD = gMax(mag(D - D.oldTime()));  res = gMax(mag(D - D.prevIter())/(maxDD + SMALL))                            */
static double mag3o(const double* v) { return sqrt(v[0]*v[0] + v[1]*v[1] + v[2]*v[2]); }

double orc_relax_residual(int nCells, double alpha, int doRelax, double* D, const double* Dprev, const double* Dold)
{
    if (doRelax)
    {
        for (int i = 0; i < 3*nCells; i++) D[i] = Dprev[i] + alpha*(D[i] - Dprev[i]);
    }
    double maxDD = -1.0e300;        /* gMax of an empty list is -VGREAT; never empty here */
    for (int c = 0; c < nCells; c++)
    {
        const double d[3] = {D[3*c] - Dold[3*c], D[3*c + 1] - Dold[3*c + 1], D[3*c + 2] - Dold[3*c + 2]};
        const double m = mag3o(d);
        if (m > maxDD) maxDD = m;
    }
    double res = -1.0e300;
    for (int c = 0; c < nCells; c++)
    {
        const double d[3] = {D[3*c] - Dprev[3*c], D[3*c + 1] - Dprev[3*c + 1], D[3*c + 2] - Dprev[3*c + 2]};
        const double r = mag3o(d)/(maxDD + 1.0e-15);
        if (r > res) res = r;
    }
    return res;
}
