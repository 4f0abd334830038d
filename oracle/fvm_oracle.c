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
    const int* patchType;       /* 0 tractionDisplacement (fixedGradient), 1 fixedDisplacement (fixedValue), 2 symmetryDisplacement,
                                   3 symmetryPlane [EXT], 4 empty, 5 fixedValue [EXT] */
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
            if (M->patchType[p] == 1 || M->patchType[p] == 5)
            {
                /* fixedValue: gradientInternalCoeffs = -pTraits::one*deltaCoeffs ;
                   fixedDisplacement::gradientBoundaryCoeffs = deltaCoeffs*(*this - (k & gradU.patchInternalField()));
                   plain [EXT] fixedValueFvPatchField (type 5): gradientBoundaryCoeffs = deltaCoeffs*(*this), no k term */
                double delta[3], k[3], kg[3];
                for (int d = 0; d < 3; d++) delta[d] = M->Cf[3*f + d] - M->C[3*c + d];
                const double nd = dot3(n, delta);
                for (int d = 0; d < 3; d++) k[d] = delta[d] - n[d]*nd;
                vdotT(k, gradD + 9*c, kg);
                if (M->patchType[p] == 5) kg[0] = kg[1] = kg[2] = 0.0;
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
    if (ptype == 1 || ptype == 5)
    {
        /* type 5 = [EXT] fixedValueFvPatchField::snGrad(): deltaCoeffs*(*this - patchInternalField()) */
        double n[3], delta[3], k[3], kg[3];
        for (int d = 0; d < 3; d++) { n[d] = M->Sf[3*f + d]/M->magSf[f]; delta[d] = M->Cf[3*f + d] - M->C[3*c + d]; }
        const double nd = dot3(n, delta);
        for (int d = 0; d < 3; d++) k[d] = delta[d] - n[d]*nd;
        vdotT(k, gradD + 9*c, kg);
        if (ptype == 5) kg[0] = kg[1] = kg[2] = 0.0;
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


/* ======================================================================================================== */
/*  Round 2: the branches of evolve() every shipped case except plateHole needs                              */
/*    d2dt2(D) Euler                    [EXT] EulerD2dt2Scheme<Type>::fvmD2dt2(vf) (same weights as the         */
/*                                      reference's own backwardD2dt2Scheme.C:306-331), then rho_* ...          */
/*    damping  K*rho*fvm::ddt(D)        unsTotalLagrangianStress.C:861-865; [EXT] EulerDdtScheme::fvmDdt        */
/*    nonLinear                         :867-881, sigma :979-982, traction BC tractionDisplacement...C:199-231   */
/*    thermal                           :883-890, sigma :986-989, traction BC :233-251                          */
/*    U = fvc::ddt(D)                   :973                                                                    */
/*    patch types 3 = symmetryPlane [EXT symmetryFvPatchField], 4 = empty (2-D cases)                          */
/*    boundary values of D (evaluate()) tractionDisplacement...C:259-300, symmetryDisplacement...C:169-207       */
/*    least-squares vol->point          numerics/leastSquaresVolPointInterpolation/*                            */
/*  Field algebra is evaluated operator by operator, left to right, as OpenFOAM's field expressions are.       */
/* ======================================================================================================== */
typedef struct orc_controls
{
    int d2dt2Scheme;        /* 0 steadyState, 1 Euler                                                    */
    int ddtScheme;          /* 0 steadyState, 1 Euler (fvm::ddt of the damping term, fvc::ddt for U)     */
    double deltaT, deltaT0;
    double K;               /* damping coefficient, active when K > SMALL                                */
    int nonLinear;          /* nonLinear && !enforceLinear                                               */
    int thermal;            /* thermalStress()                                                           */
    double limitCoeff;      /* snGrad scheme: skewCorrected <limitCoeff>                                 */
    /* `backward` schemes (d2dt2Scheme / ddtScheme == 2; the shipped 3dTube and HronTurek solids):
     *   d2dt2(D)  numerics/backwardD2dt2Scheme/backwardD2dt2Scheme.C:230-280 (fvmD2dt2(vf), then rho_* ... at
     *             unsTotalLagrangianStress.C:848), built on [EXT] backwardDdtScheme::fvmDdt / fvcDdt
     *   ddt(D)    [EXT] backwardDdtScheme::fvmDdt (damping term :864) and ::fvcDdt (U, :973)
     * Each use takes its own "deltaT0": GREAT until enough old-time levels exist -- backwardD2dt2Scheme.C:58-76
     * (timeIndex of the 2nd and 3rd old level equal), [EXT] backwardDdtScheme::deltaT0_(vf) (vf.nOldTimes() < 2) -- so the
     * caller, who owns the time levels (oracle/evolve.py: OldTimes), passes the effective value of every use:
     *   eD2 fvmD2dt2(D); eDdtD the fvmDdt(D) inside it; eDdtD0, eDdtD00 the fvcDdt(D.oldTime()) and
     *   fvcDdt(D.oldTime().oldTime()) inside it; eDamp the fvm::ddt(D) of the damping term.
     * Dooo, Doooo = the 3rd and 4th old-time level of D (cell fields). */
    double eD2, eDdtD, eDdtD0, eDdtD00, eDamp;
    const double* Dooo;
    const double* Doooo;
} orc_controls;

/* coefft, coefft00, coefft0 of the backward schemes for time step deltaT and effective old time step e */
static void backward_weights(double deltaT, double e, double* c, double* c00, double* c0)
{
    *c = 1 + deltaT/(deltaT + e);
    *c00 = deltaT*deltaT/(e*(deltaT + e));
    *c0 = *c + *c00;
}

/* (A & B)_ij = A_ik B_kj, accumulated k = x, y, z   [EXT TensorI.H operator&] */
static void TdotT(const double* A, const double* B, double* R)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[3*i + j] = A[3*i]*B[j] + A[3*i + 1]*B[3 + j] + A[3*i + 2]*B[6 + j];
}
static void transposeT(const double* A, double* R)
{
    R[0] = A[0]; R[1] = A[3]; R[2] = A[6]; R[3] = A[1]; R[4] = A[4]; R[5] = A[7]; R[6] = A[2]; R[7] = A[5]; R[8] = A[8];
}
/* symm(T) = (xx, 0.5*(xy+yx), 0.5*(xz+zx), yy, 0.5*(yz+zy), zz)   [EXT TensorI.H] */
static void symmT(const double* T, double* S)
{
    S[0] = T[0]; S[1] = 0.5*(T[1] + T[3]); S[2] = 0.5*(T[2] + T[6]); S[3] = T[4]; S[4] = 0.5*(T[5] + T[7]); S[5] = T[8];
}
/* symmTensor (xx xy xz yy yz zz) & tensor -> tensor   [EXT SymmTensorI.H / TensorI.H] */
static void SdotT(const double* S, const double* T, double* R)
{
    const double F[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};
    TdotT(F, T, R);
}
/* T & v */
static void Tdotv(const double* T, const double* v, double* r)
{
    r[0] = T[0]*v[0] + T[1]*v[1] + T[2]*v[2];
    r[1] = T[3]*v[0] + T[4]*v[1] + T[5]*v[2];
    r[2] = T[6]*v[0] + T[7]*v[1] + T[8]*v[2];
}
static double detT(const double* t)
{
    return (t[0]*t[4]*t[8] + t[1]*t[5]*t[6] + t[2]*t[3]*t[7] - t[0]*t[5]*t[7] - t[1]*t[3]*t[8] - t[2]*t[4]*t[6]);
}
/* inv(t) = cofactor matrix / det(t)   [EXT TensorI.H inv(t, det(t))]; hinv(t) == inv(t) for det(t) > 1e-10 [EXT tensor.C] */
static void invT(const double* t, double* r)
{
    const double d = detT(t);
    r[0] = (t[4]*t[8] - t[7]*t[5])/d; r[1] = (t[2]*t[7] - t[1]*t[8])/d; r[2] = (t[1]*t[5] - t[2]*t[4])/d;
    r[3] = (t[6]*t[5] - t[3]*t[8])/d; r[4] = (t[0]*t[8] - t[2]*t[6])/d; r[5] = (t[3]*t[2] - t[0]*t[5])/d;
    r[6] = (t[3]*t[7] - t[4]*t[6])/d; r[7] = (t[1]*t[6] - t[0]*t[7])/d; r[8] = (t[0]*t[4] - t[3]*t[1])/d;
}

/* the two extra explicit fluxes of the non-linear branch, unsTotalLagrangianStress.C:869-880:
 *   X = muf*(Sf & (G & G.T())) + 0.5*lambdaf*tr(G & G.T())*Sf
 *   Y = Sf & (sigmaf & G),  Ef = symm(G) + 0.5*symm(G & G.T()),  sigmaf = 2*muf*Ef + I*(lambdaf*tr(Ef))        */
static void nonlinear_fluxes(const double* Sf, double muf, double laf, const double* G, double* X, double* Y)
{
    double GT[9], GG[9], v[3], sG[6], sGG[6], Ef[6], sg[6], SG[9];
    transposeT(G, GT);
    TdotT(G, GT, GG);
    vdotT(Sf, GG, v);
    const double trGG = GG[0] + GG[4] + GG[8];
    const double c = (0.5*laf)*trGG;
    for (int d = 0; d < 3; d++) X[d] = muf*v[d] + c*Sf[d];
    symmT(G, sG);
    symmT(GG, sGG);
    for (int k = 0; k < 6; k++) Ef[k] = sG[k] + 0.5*sGG[k];
    const double trE = Ef[0] + Ef[3] + Ef[5];
    const double tm = 2*muf, sph = laf*trE;
    for (int k = 0; k < 6; k++) sg[k] = tm*Ef[k];
    sg[0] += sph; sg[3] += sph; sg[5] += sph;
    SdotT(sg, G, SG);
    vdotT(Sf, SG, Y);
}

static void patch_nk(const orc_mesh* M, int f, double* n, double* k)
{
    const int c = M->owner[f];
    double delta[3];
    for (int d = 0; d < 3; d++) { n[d] = M->Sf[3*f + d]/M->magSf[f]; delta[d] = M->Cf[3*f + d] - M->C[3*c + d]; }
    const double nd = dot3(n, delta);
    for (int d = 0; d < 3; d++) k[d] = delta[d] - n[d]*nd;
}

/* [EXT] symmetryFvPatchField<vector>::snGrad(): (transform(I - 2.0*sqr(nHat), pif) - pif)*(deltaCoeffs/2.0)   (patch type 3) */
static void symmetryPlane_snGrad(const orc_mesh* M, int f, const double* D, double* sn)
{
    const int c = M->owner[f];
    double n[3], S[6], T[6], tv[3];
    const double* UP = D + 3*c;
    for (int d = 0; d < 3; d++) n[d] = M->Sf[3*f + d]/M->magSf[f];
    S[0] = n[0]*n[0]; S[1] = n[0]*n[1]; S[2] = n[0]*n[2]; S[3] = n[1]*n[1]; S[4] = n[1]*n[2]; S[5] = n[2]*n[2];
    T[0] = 1.0 - 2.0*S[0]; T[1] = -(2.0*S[1]); T[2] = -(2.0*S[2]); T[3] = 1.0 - 2.0*S[3]; T[4] = -(2.0*S[4]); T[5] = 1.0 - 2.0*S[5];
    tv[0] = T[0]*UP[0] + T[1]*UP[1] + T[2]*UP[2];
    tv[1] = T[1]*UP[0] + T[3]*UP[1] + T[4]*UP[2];
    tv[2] = T[2]*UP[0] + T[4]*UP[1] + T[5]*UP[2];
    const double h = M->deltaCoeffs[f]/2.0;
    for (int d = 0; d < 3; d++) sn[d] = (tv[d] - UP[d])*h;
}

/* tractionDisplacementFvPatchVectorField::updateCoeffs, all branches (:148-257): gradient() of one face */
static void traction_gradient(const orc_mesh* M, const orc_controls* ctl, int f, double m, double la, const double* G,
                              const double* traction, double pressure, double threeK, double alpha, double DTb, double* grad)
{
    double n[3], t[3], T[9], nT[3];
    for (int d = 0; d < 3; d++) { n[d] = M->Sf[3*f + d]/M->magSf[f]; t[d] = traction[d]; }
    if (ctl->nonLinear)
    {
        double F[9], invF[9], nC[3], Jn[3], tv[3];
        for (int k = 0; k < 9; k++) F[k] = G[k];
        F[0] = 1.0 + G[0]; F[4] = 1.0 + G[4]; F[8] = 1.0 + G[8];       /* I + gradDf */
        const double J = detT(F);
        invT(F, invF);
        Tdotv(invF, n, nC);                                            /* invF & n */
        for (int d = 0; d < 3; d++) Jn[d] = J*nC[d];
        const double SoS0 = mag3(Jn);
        const double mc = mag3(nC);
        for (int d = 0; d < 3; d++) nC[d] /= mc;
        for (int d = 0; d < 3; d++) t[d] -= pressure*nC[d];
        vdotT(t, invF, tv);                                            /* t & invF */
        for (int d = 0; d < 3; d++) t[d] = tv[d]*SoS0;
    }
    else
    {
        for (int d = 0; d < 3; d++) t[d] -= pressure*n[d];
    }
    const double ml = m + la;
    T[0] = m*G[0] - ml*G[0]; T[1] = m*G[3] - ml*G[1]; T[2] = m*G[6] - ml*G[2];
    T[3] = m*G[1] - ml*G[3]; T[4] = m*G[4] - ml*G[4]; T[5] = m*G[7] - ml*G[5];
    T[6] = m*G[2] - ml*G[6]; T[7] = m*G[5] - ml*G[7]; T[8] = m*G[8] - ml*G[8];
    vdotT(n, T, nT);
    const double tr = G[0] + G[4] + G[8];
    for (int d = 0; d < 3; d++) grad[d] = (t[d] - nT[d]) - n[d]*la*tr;
    if (ctl->nonLinear)
    {
        /* gradient() -= (n & (mu*(gradDf & gradDf.T()))) + 0.5*n*lambda*tr(gradDf & gradDf.T()) */
        double GT[9], GG[9], mGG[9], a[3];
        transposeT(G, GT);
        TdotT(G, GT, GG);
        for (int k = 0; k < 9; k++) mGG[k] = m*GG[k];
        vdotT(n, mGG, a);
        const double trGG = GG[0] + GG[4] + GG[8];
        for (int d = 0; d < 3; d++) grad[d] -= a[d] + ((0.5*n[d])*la)*trGG;
    }
    if (ctl->thermal)
    {
        for (int d = 0; d < 3; d++) grad[d] += ((n[d]*threeK)*alpha)*DTb;
    }
    const double den = 2.0*m + la;
    for (int d = 0; d < 3; d++) grad[d] /= den;
}

/*
 * The D equation with every branch of evolve() (unsTotalLagrangianStress.C:846-890).  Extra inputs over orc_assemble_D:
 * Dold = D.oldTime(), Doldold = D.oldTime().oldTime() (cell fields), threeK / alpha / DT (cell fields; boundary value of
 * threeKf, alphaf = the owner cell's value as for muf), DTb [nBoundaryFaces] boundary values of DT.  Unused ones may be NULL.
 */
void orc_assemble_D_ex
(
    const orc_mesh* M, const orc_controls* ctl, const double* mu, const double* lambda, const double* rho, const double* g,
    const double* D, const double* Dold, const double* Doldold, const double* gradD, const double* gradDf,
    const double* traction, const double* pressure, const double* fixedValue,
    const double* threeK, const double* alpha, const double* DT, const double* DTb,
    double* upper, double* diag, double* source, double* internalCoeffs, double* boundaryCoeffs, double* patchGradient
)
{
    const int nC = M->nCells, nF = M->nFaces, nIF = M->nInternalFaces;
    double* muf = (double*)malloc(sizeof(double)*nF);
    double* lambdaf = (double*)malloc(sizeof(double)*nF);
    double* gms = (double*)malloc(sizeof(double)*nF);
    double* fluxE = (double*)calloc(3*(size_t)nF, sizeof(double));
    double* fluxC = (double*)calloc(3*(size_t)(nIF > 0 ? nIF : 1), sizeof(double));
    double* fluxN = (double*)calloc(3*(size_t)nF, sizeof(double));     /* X + Y of the non-linear branch, per face */
    double* fluxX = (double*)calloc(3*(size_t)nF, sizeof(double));
    double* fluxY = (double*)calloc(3*(size_t)nF, sizeof(double));
    double* fluxT = (double*)calloc(3*(size_t)nF, sizeof(double));
    double* sumE = (double*)calloc(3*(size_t)nC, sizeof(double));
    double* sumC = (double*)calloc(3*(size_t)nC, sizeof(double));
    double* sumX = (double*)calloc(3*(size_t)nC, sizeof(double));
    double* sumY = (double*)calloc(3*(size_t)nC, sizeof(double));
    double* sumT = (double*)calloc(3*(size_t)nC, sizeof(double));
    double* Ldiag = (double*)calloc(nC, sizeof(double));
    unsigned char* emptyFace = (unsigned char*)calloc(nF > 0 ? nF : 1, 1);
    for (int p = 0; p < M->nPatches; p++)
        if (M->patchType[p] == 4)
            for (int i = 0; i < M->patchSize[p]; i++) emptyFace[M->patchStart[p] + i] = 1;

    orc_interpolate(M, mu, muf);
    orc_interpolate(M, lambda, lambdaf);
    for (int f = 0; f < nF; f++) gms[f] = (2*muf[f] + lambdaf[f])*M->magSf[f];

    for (int f = 0; f < nIF; f++)
    {
        const double up = M->deltaCoeffs[f]*gms[f];
        Ldiag[M->owner[f]] -= up;
        Ldiag[M->neighbour[f]] -= up;
        upper[f] = 0.0 - up;
    }
    for (int f = 0; f < nIF; f++)
    {
        double ssf[3];
        skew_correction(M, f, D, gradD, ctl->limitCoeff, ssf);
        for (int d = 0; d < 3; d++) fluxC[3*f + d] = gms[f]*ssf[d];
    }
    for (int f = 0; f < nF; f++)
    {
        if (emptyFace[f]) continue;             /* empty fvsPatchFields have no faces */
        explicit_flux(M->Sf + 3*f, muf[f], lambdaf[f], gradDf + 9*(size_t)f, fluxE + 3*f);
        if (ctl->nonLinear) nonlinear_fluxes(M->Sf + 3*f, muf[f], lambdaf[f], gradDf + 9*(size_t)f, fluxX + 3*f, fluxY + 3*f);
        if (ctl->thermal)
        {
            /* mesh().Sf()*threeKf()*alphaf()*DTf()  :888 */
            double tk, al, dt;
            if (f < nIF)
            {
                const int P = M->owner[f], N = M->neighbour[f];
                const double w = M->weights[f];
                tk = w*(threeK[P] - threeK[N]) + threeK[N];
                al = w*(alpha[P] - alpha[N]) + alpha[N];
                dt = w*(DT[P] - DT[N]) + DT[N];
            }
            else { tk = threeK[M->owner[f]]; al = alpha[M->owner[f]]; dt = DTb[f - nIF]; }
            for (int d = 0; d < 3; d++) fluxT[3*f + d] = ((M->Sf[3*f + d]*tk)*al)*dt;
        }
    }
    for (int f = 0; f < nIF; f++)
    {
        const int P = M->owner[f], N = M->neighbour[f];
        for (int d = 0; d < 3; d++)
        {
            sumC[3*P + d] += fluxC[3*f + d]; sumC[3*N + d] -= fluxC[3*f + d];
            sumE[3*P + d] += fluxE[3*f + d]; sumE[3*N + d] -= fluxE[3*f + d];
            sumX[3*P + d] += fluxX[3*f + d]; sumX[3*N + d] -= fluxX[3*f + d];
            sumY[3*P + d] += fluxY[3*f + d]; sumY[3*N + d] -= fluxY[3*f + d];
            sumT[3*P + d] += fluxT[3*f + d]; sumT[3*N + d] -= fluxT[3*f + d];
        }
    }
    for (int f = nIF; f < nF; f++)
    {
        if (emptyFace[f]) continue;
        const int P = M->owner[f];
        for (int d = 0; d < 3; d++)
        {
            sumE[3*P + d] += fluxE[3*f + d]; sumX[3*P + d] += fluxX[3*f + d];
            sumY[3*P + d] += fluxY[3*f + d]; sumT[3*P + d] += fluxT[3*f + d];
        }
    }
    /* time-scheme coefficients */
    double coefft = 0, coefft00 = 0, rDeltaT2 = 0, rDeltaT = 0;
    if (ctl->d2dt2Scheme == 1)
    {
        coefft = (ctl->deltaT + ctl->deltaT0)/(2*ctl->deltaT);
        coefft00 = (ctl->deltaT + ctl->deltaT0)/(2*ctl->deltaT0);
        const double s = ctl->deltaT + ctl->deltaT0;
        rDeltaT2 = 4.0/(s*s);
    }
    if (ctl->ddtScheme == 1) rDeltaT = 1.0/ctl->deltaT;
    const int damping = ctl->K > ORC_SMALL;
    for (int c = 0; c < nC; c++)
    {
        const double V = M->V[c];
        /* A = rho*fvm::d2dt2(D) */
        double dA = 0.0, sA[3] = {0.0, 0.0, 0.0};
        if (ctl->d2dt2Scheme == 1)
        {
            dA = ((coefft*rDeltaT2)*V)*rho[c];
            const double t1 = rDeltaT2*V;
            for (int d = 0; d < 3; d++)
                sA[d] = (t1*((coefft + coefft00)*Dold[3*c + d] - coefft00*Doldold[3*c + d]))*rho[c];
        }
        else if (ctl->d2dt2Scheme == 2)
        {
            /* backwardD2dt2Scheme::fvmD2dt2(vf)  :230-280
             *   fvm = coefft*rDeltaT * backwardDdtScheme.fvmDdt(vf)
             *         [EXT fvmDdt: diag = (coefft*rDeltaT)*V, source = rDeltaT*V*(coefft0*vf.oldTime() - coefft00*vf.oldTime().oldTime())]
             *   fvm.source() += rDeltaT*V*(coefft0*fvcDdt(vf.oldTime()) - coefft00*fvcDdt(vf.oldTime().oldTime()))
             *         [EXT fvcDdt(w) = rDeltaT*(coefft*w - coefft0*w.oldTime() + coefft00*w.oldTime().oldTime())]
             * then rho_* (diag and source scaled by rho, as for Euler)                                                  */
            const double dT = ctl->deltaT, rDt = 1.0/dT;
            double c2, c200, c20, cb, cb00, cb0, c3, c300, c30, c4, c400, c40;
            backward_weights(dT, ctl->eD2, &c2, &c200, &c20);
            backward_weights(dT, ctl->eDdtD, &cb, &cb00, &cb0);
            backward_weights(dT, ctl->eDdtD0, &c3, &c300, &c30);
            backward_weights(dT, ctl->eDdtD00, &c4, &c400, &c40);
            const double sc = c2*rDt;
            dA = (((cb*rDt)*V)*sc)*rho[c];
            const double rV = rDt*V;
            for (int d = 0; d < 3; d++)
            {
                const double D0 = Dold[3*c + d], D00 = Doldold[3*c + d], D000 = ctl->Dooo[3*c + d], D0000 = ctl->Doooo[3*c + d];
                const double s1 = (rV*(cb0*D0 - cb00*D00))*sc;
                const double X = rDt*(((c3*D0) - (c30*D00)) + (c300*D000));
                const double Y = rDt*(((c4*D00) - (c40*D000)) + (c400*D0000));
                const double s2 = s1 + rV*(c20*X - c200*Y);
                sA[d] = s2*rho[c];
            }
        }
        double dg = dA - Ldiag[c];
        double src[3];
        for (int d = 0; d < 3; d++)
        {
            const double dc = sumC[3*c + d]/V, de = sumE[3*c + d]/V;
            const double rg = rho[c]*g[d];
            const double R = ((0.0 - V*dc) - V*de) - V*rg;
            src[d] = sA[d] - R;
        }
        if (damping)
        {
            /* DEqn += K*rho_*fvm::ddt(D_)    :864 */
            const double sf = ctl->K*rho[c];
            double dM = 0.0, sM[3] = {0.0, 0.0, 0.0};
            if (ctl->ddtScheme == 1)
            {
                dM = (rDeltaT*V)*sf;
                for (int d = 0; d < 3; d++) sM[d] = ((rDeltaT*Dold[3*c + d])*V)*sf;
            }
            else if (ctl->ddtScheme == 2)
            {
                /* [EXT] backwardDdtScheme::fvmDdt(vf), then scaled by K*rho */
                const double rDt = 1.0/ctl->deltaT;
                double cb, cb00, cb0;
                backward_weights(ctl->deltaT, ctl->eDamp, &cb, &cb00, &cb0);
                dM = ((cb*rDt)*V)*sf;
                for (int d = 0; d < 3; d++) sM[d] = ((rDt*V)*(cb0*Dold[3*c + d] - cb00*Doldold[3*c + d]))*sf;
            }
            dg += dM;
            for (int d = 0; d < 3; d++) src[d] += sM[d];
        }
        if (ctl->nonLinear)
        {
            /* DEqn -= fvc::div(X) + fvc::div(Y)   ->  source += V*(divX + divY) */
            for (int d = 0; d < 3; d++) src[d] += V*(sumX[3*c + d]/V + sumY[3*c + d]/V);
        }
        if (ctl->thermal)
        {
            /* DEqn += fvc::div(...)  ->  source -= V*div */
            for (int d = 0; d < 3; d++) src[d] -= V*(sumT[3*c + d]/V);
        }
        diag[c] = dg;
        for (int d = 0; d < 3; d++) source[3*c + d] = src[d];
    }

    /* boundary coefficients */
    for (int p = 0; p < M->nPatches; p++)
    {
        for (int i = 0; i < M->patchSize[p]; i++)
        {
            const int f = M->patchStart[p] + i, bf = f - nIF, c = M->owner[f];
            const double dcf = M->deltaCoeffs[f];
            double n[3], gIC[3], gBC[3];
            for (int d = 0; d < 3; d++) n[d] = M->Sf[3*f + d]/M->magSf[f];
            const int pt = M->patchType[p];
            if (pt == 4)
            {
                for (int d = 0; d < 3; d++) { internalCoeffs[3*bf + d] = 0.0; boundaryCoeffs[3*bf + d] = 0.0; }
                continue;
            }
            if (pt == 1 || pt == 5)
            {
                double k[3], kg[3];
                patch_nk(M, f, n, k);
                vdotT(k, gradD + 9*(size_t)c, kg);
                if (pt == 5) kg[0] = kg[1] = kg[2] = 0.0;           /* [EXT] fixedValueFvPatchField: no k term */
                for (int d = 0; d < 3; d++) { gIC[d] = -dcf; gBC[d] = dcf*(fixedValue[3*bf + d] - kg[d]); }
            }
            else if (pt == 2 || pt == 3)
            {
                double sn[3];
                if (pt == 2) symmetry_snGrad(M, f, D, gradD, sn); else symmetryPlane_snGrad(M, f, D, sn);
                for (int d = 0; d < 3; d++) { gIC[d] = -dcf*fabs(n[d]); gBC[d] = sn[d] - gIC[d]*D[3*c + d]; }
            }
            else
            {
                traction_gradient(M, ctl, f, muf[f], lambdaf[f], gradDf + 9*(size_t)f, traction + 3*bf, pressure[bf],
                                  ctl->thermal ? threeK[c] : 0.0, ctl->thermal ? alpha[c] : 0.0, ctl->thermal ? DTb[bf] : 0.0, gBC);
                for (int d = 0; d < 3; d++) { gIC[d] = 0.0; if (patchGradient) patchGradient[3*bf + d] = gBC[d]; }
            }
            for (int d = 0; d < 3; d++)
            {
                const double Lic = gms[f]*gIC[d];
                const double Lbc = (-gms[f])*gBC[d];
                /* rho*d2dt2 and the damping matrix carry zero patch coefficients: 0 - L (then += 0) */
                internalCoeffs[3*bf + d] = 0.0 - Lic;
                boundaryCoeffs[3*bf + d] = 0.0 - Lbc;
            }
        }
    }
    free(muf); free(lambdaf); free(gms); free(fluxE); free(fluxC); free(fluxN); free(fluxX); free(fluxY); free(fluxT);
    free(sumE); free(sumC); free(sumX); free(sumY); free(sumT); free(Ldiag); free(emptyFace);
}

/* D.correctBoundaryConditions() at the end of fvMatrix::solve: evaluate() of every patch -> boundary values Db[nBF*3]
 *   tractionDisplacement  D_P + (k & gradD_P) + gradient()/deltaCoeffs                tractionDisplacement...C:259-300
 *   fixedDisplacement     the fixed value
 *   symmetryDisplacement  (UP + transform(I - 2.0*sqr(nHat), UP))/2.0, UP = D_P + (k & gradD_P)   symmetryDisplacement...C:169-207
 *   symmetryPlane [EXT]   (pif + transform(I - 2.0*sqr(nHat), pif))/2.0
 *   empty                 no faces (slots left untouched)                                                         */
void orc_patch_evaluate(const orc_mesh* M, const double* D, const double* gradD, const double* fixedValue,
                        const double* patchGradient, double* Db)
{
    const int nIF = M->nInternalFaces;
    for (int p = 0; p < M->nPatches; p++)
    {
        const int pt = M->patchType[p];
        for (int i = 0; i < M->patchSize[p]; i++)
        {
            const int f = M->patchStart[p] + i, bf = f - nIF, c = M->owner[f];
            double n[3], k[3], kg[3];
            if (pt == 4) continue;
            if (pt == 1 || pt == 5) { for (int d = 0; d < 3; d++) Db[3*bf + d] = fixedValue[3*bf + d]; continue; }
            patch_nk(M, f, n, k);
            if (pt == 0)
            {
                vdotT(k, gradD + 9*(size_t)c, kg);
                for (int d = 0; d < 3; d++) Db[3*bf + d] = (D[3*c + d] + kg[d]) + patchGradient[3*bf + d]/M->deltaCoeffs[f];
                continue;
            }
            double UP[3], S[6], T[6], tv[3];
            if (pt == 2) { vdotT(k, gradD + 9*(size_t)c, kg); for (int d = 0; d < 3; d++) UP[d] = D[3*c + d] + kg[d]; }
            else { for (int d = 0; d < 3; d++) UP[d] = D[3*c + d]; }
            S[0] = n[0]*n[0]; S[1] = n[0]*n[1]; S[2] = n[0]*n[2]; S[3] = n[1]*n[1]; S[4] = n[1]*n[2]; S[5] = n[2]*n[2];
            T[0] = 1.0 - 2.0*S[0]; T[1] = -(2.0*S[1]); T[2] = -(2.0*S[2]); T[3] = 1.0 - 2.0*S[3]; T[4] = -(2.0*S[4]); T[5] = 1.0 - 2.0*S[5];
            tv[0] = T[0]*UP[0] + T[1]*UP[1] + T[2]*UP[2];
            tv[1] = T[1]*UP[0] + T[3]*UP[1] + T[4]*UP[2];
            tv[2] = T[2]*UP[0] + T[4]*UP[1] + T[5]*UP[2];
            for (int d = 0; d < 3; d++) Db[3*bf + d] = (UP[d] + tv[d])/2.0;
        }
    }
}

/* boundary snGrad with the two extra patch types (used by orc_grad_ex / orc_fgrad_ex) */
static int patch_snGrad_ex(const orc_mesh* M, int ptype, int f, const double* D, const double* gradD,
                           const double* fixedValue, const double* patchGradient, double* sn)
{
    if (ptype == 4) return 0;
    if (ptype == 3) { symmetryPlane_snGrad(M, f, D, sn); return 1; }
    patch_snGrad(M, ptype, f, D, gradD, fixedValue, patchGradient, sn);
    return 1;
}

/* fvc::grad / fvc::fGrad with empty and symmetryPlane patches: empty faces take part in the cell-gradient face fans
 * (fvcGradf.C:613-686 loops over the polyPatches) but carry no boundary gradient and no face gradient (their fvPatch has
 * size 0): those slots are left as they are.                                                                         */
void orc_grad_ex(const orc_mesh* M, const double* D, const double* pointD, const double* gradDold,
                 const double* fixedValue, const double* patchGradient, double* gradD, double* gradDb)
{
    const int nC = M->nCells, nIF = M->nInternalFaces, nF = M->nFaces;
    double* V = (double*)calloc(nC, sizeof(double));
    memset(gradD, 0, sizeof(double)*9*(size_t)nC);
    for (int f = 0; f < nIF; f++)
    {
        grad_face(M, f, pointD, +1.0, gradD + 9*(size_t)M->owner[f], V + M->owner[f]);
        grad_face(M, f, pointD, -1.0, gradD + 9*(size_t)M->neighbour[f], V + M->neighbour[f]);
    }
    for (int f = nIF; f < nF; f++) grad_face(M, f, pointD, +1.0, gradD + 9*(size_t)M->owner[f], V + M->owner[f]);
    for (int c = 0; c < nC; c++)
    {
        V[c] /= 3;
        for (int k = 0; k < 9; k++) gradD[9*(size_t)c + k] /= V[c];
    }
    for (int p = 0; p < M->nPatches; p++)
    {
        for (int i = 0; i < M->patchSize[p]; i++)
        {
            const int f = M->patchStart[p] + i, bf = f - nIF;
            double* G = gradDb + 9*(size_t)bf;
            double n[3], sn[3], nG[3];
            if (!patch_snGrad_ex(M, M->patchType[p], f, D, gradDold, fixedValue, patchGradient, sn)) continue;
            face_tangential_grad(M, f, pointD, G);
            for (int d = 0; d < 3; d++) n[d] = M->Sf[3*f + d]/M->magSf[f];
            vdotT(n, G, nG);
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) G[3*a + b] += n[a]*(sn[b] - nG[b]);
        }
    }
    free(V);
}

void orc_fgrad_ex(const orc_mesh* M, const double* D, const double* pointD, const double* gradD,
                  const double* fixedValue, const double* patchGradient, double limitCoeff, double* gradDf)
{
    const int nIF = M->nInternalFaces;
    for (int f = 0; f < nIF; f++)
    {
        double ssf[3], sn[3], n[3];
        face_tangential_grad(M, f, pointD, gradDf + 9*(size_t)f);
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
            if (!patch_snGrad_ex(M, M->patchType[p], f, D, gradD, fixedValue, patchGradient, sn)) continue;
            face_tangential_grad(M, f, pointD, gradDf + 9*(size_t)f);
            for (int d = 0; d < 3; d++) n[d] = M->Sf[3*f + d]/M->magSf[f];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) gradDf[9*(size_t)f + 3*a + b] += n[a]*sn[b];
        }
    }
}

/* epsilon / sigma with the non-linear and thermal terms (unsTotalLagrangianStress.C:975-990):
 *   epsilon = symm(gradD) [+ 0.5*symm(gradD & gradD.T())];  sigma = 2*mu*epsilon + I*(lambda*tr(epsilon))
 *   [- symmTensor(1,0,0,1,0,1)*threeK*alpha*DT]                                                              */
void orc_sigma_ex(int nCells, int nonLinear, int thermal, const double* mu, const double* lambda, const double* gradD,
                  const double* threeK, const double* alpha, const double* DT, double* epsilon, double* sigma)
{
    for (int c = 0; c < nCells; c++)
    {
        const double* G = gradD + 9*(size_t)c;
        double e[6];
        symmT(G, e);
        if (nonLinear)
        {
            double GT[9], GG[9], s[6];
            transposeT(G, GT); TdotT(G, GT, GG); symmT(GG, s);
            for (int k = 0; k < 6; k++) e[k] += 0.5*s[k];
        }
        const double tr = e[0] + e[3] + e[5];
        const double sph = lambda[c]*tr;
        const double tm = 2*mu[c];
        double sg[6];
        for (int k = 0; k < 6; k++) sg[k] = tm*e[k];
        sg[0] += sph; sg[3] += sph; sg[5] += sph;
        if (thermal)
        {
            static const double one[6] = {1, 0, 0, 1, 0, 1};
            for (int k = 0; k < 6; k++) sg[k] -= ((one[k]*threeK[c])*alpha[c])*DT[c];
        }
        for (int k = 0; k < 6; k++) { epsilon[6*(size_t)c + k] = e[k]; sigma[6*(size_t)c + k] = sg[k]; }
    }
}

/* min / max of det(I + gradDf) over the (non-empty) faces   unsTotalLagrangianStress.C:929-946 */
void orc_det_minmax(const orc_mesh* M, const double* gradDf, double* minDet, double* maxDet)
{
    double mn = 1.0e300, mx = -1.0e300;
    unsigned char* emptyFace = (unsigned char*)calloc(M->nFaces > 0 ? M->nFaces : 1, 1);
    for (int p = 0; p < M->nPatches; p++)
        if (M->patchType[p] == 4)
            for (int i = 0; i < M->patchSize[p]; i++) emptyFace[M->patchStart[p] + i] = 1;
    for (int f = 0; f < M->nFaces; f++)
    {
        if (emptyFace[f]) continue;
        double F[9];
        for (int k = 0; k < 9; k++) F[k] = gradDf[9*(size_t)f + k];
        F[0] = 1.0 + F[0]; F[4] = 1.0 + F[4]; F[8] = 1.0 + F[8];
        const double d = detT(F);
        if (d < mn) mn = d;
        if (d > mx) mx = d;
    }
    free(emptyFace);
    *minDet = mn; *maxDet = mx;
}

/* U = fvc::ddt(D)  [EXT EulerDdtScheme::fvcDdt]: rDeltaT*(D - D.oldTime()); steadyState: 0    :973 */
void orc_ddt(int nCells, int ddtScheme, double deltaT, const double* D, const double* Dold, double* U)
{
    const double rDeltaT = ddtScheme == 1 ? 1.0/deltaT : 0.0;
    for (int i = 0; i < 3*nCells; i++) U[i] = ddtScheme == 1 ? rDeltaT*(D[i] - Dold[i]) : 0.0;
}

/* U = fvc::ddt(D), backward [EXT backwardDdtScheme::fvcDdt]: rDeltaT*(coefft*D - coefft0*D.oldTime() + coefft00*D.oldTime().oldTime()),
 * e = the effective deltaT0 of this use (GREAT while D has fewer than two old levels) */
void orc_ddt_backward(int nCells, double deltaT, double e, const double* D, const double* Dold, const double* Doldold, double* U)
{
    const double rDt = 1.0/deltaT;
    double c, c00, c0;
    backward_weights(deltaT, e, &c, &c00, &c0);
    for (int i = 0; i < 3*nCells; i++) U[i] = rDt*(((c*D[i]) - (c0*Dold[i])) + (c00*Doldold[i]));
}

/* GeometricField::relax on the boundary field: operator== assigns prevIter + alpha*(this - prevIter) on every patch */
void orc_relax_boundary(int n3, double alpha, double* Db, const double* DbPrev)
{
    for (int i = 0; i < n3; i++) Db[i] = DbPrev[i] + alpha*(Db[i] - DbPrev[i]);
}


/* ======================================================================================================== */
/*  leastSquaresVolPointInterpolation (numerics/leastSquaresVolPointInterpolation/), serial case:            */
/*    stencil of a point = its cells (mesh.pointCells(), ascending [EXT primitiveMeshPointCells.C]) followed by */
/*    its boundary faces that are neither empty nor wedge (makePointFaces, ...Interpolation.C:54-205) in the      */
/*    order of labelHashSet::toc() [EXT HashTable: 128 buckets, bucket = label & 127, newest first in a bucket];  */
/*    points on an empty patch mirror the whole stencil about the plane through the point with the patch's        */
/*    point normal (makeMirrorPlaneTransformation :1829-1940, T = I);  weights are all 1 (makeWeights :1116-1133);*/
/*    origin = mean of the stencil points (makeOrigins :1405-1421);  invLs = inv(M^T M) M^T (makeInvLsMatrices    */
/*    :1718-1826);  interpolate = leastSquaresVolPointInterpolate.C:89-299.                                       */
/* ======================================================================================================== */
typedef struct orc_lsq
{
    int nPoints;
    int* start;         /* [nPoints+1] stencil rows per point INCLUDING mirrored rows */
    int* entry;         /* per row: cell id >= 0, or -1 - boundaryFaceIndex (face - nInternalFaces) */
    double* inv;        /* [3*rows]: row j of point p holds invLs[0][j], invLs[1][j], invLs[2][j] */
    double* origin;     /* [3*nPoints] */
    unsigned char* mirror;  /* [nPoints] */
    int nSrc;           /* scratch size: max rows */
} orc_lsq;

void orc_lsq_free(orc_lsq* L)
{
    if (!L) return;
    free(L->start); free(L->entry); free(L->inv); free(L->origin); free(L->mirror); free(L);
}

orc_lsq* orc_lsq_build(const orc_mesh* M)
{
    const int nP = M->nPoints, nF = M->nFaces, nIF = M->nInternalFaces, nC = M->nCells;
    orc_lsq* L = (orc_lsq*)calloc(1, sizeof(orc_lsq));
    L->nPoints = nP;
    /* patch type per boundary face */
    int* bfType = (int*)malloc(sizeof(int)*(nF - nIF > 0 ? nF - nIF : 1));
    for (int p = 0; p < M->nPatches; p++)
        for (int i = 0; i < M->patchSize[p]; i++) bfType[M->patchStart[p] + i - nIF] = M->patchType[p];
    /* cell -> faces to enumerate cell points; pointCells: for cellI ascending, labels of the cell's points (each once) */
    int* cfs = (int*)calloc(nC + 1, sizeof(int));
    for (int f = 0; f < nF; f++) cfs[M->owner[f] + 1]++;
    for (int f = 0; f < nIF; f++) cfs[M->neighbour[f] + 1]++;
    for (int c = 0; c < nC; c++) cfs[c + 1] += cfs[c];
    int* cf = (int*)malloc(sizeof(int)*(cfs[nC] > 0 ? cfs[nC] : 1));
    int* cur = (int*)malloc(sizeof(int)*(nC > 0 ? nC : 1));
    for (int c = 0; c < nC; c++) cur[c] = cfs[c];
    for (int f = 0; f < nF; f++) { cf[cur[M->owner[f]]++] = f; if (f < nIF) cf[cur[M->neighbour[f]]++] = f; }
    int* pcCount = (int*)calloc(nP + 1, sizeof(int));
    int* lastCell = (int*)malloc(sizeof(int)*(nP > 0 ? nP : 1));
    for (int p = 0; p < nP; p++) lastCell[p] = -1;
    for (int c = 0; c < nC; c++)
        for (int k = cfs[c]; k < cfs[c + 1]; k++)
            for (int q = M->faceStart[cf[k]]; q < M->faceStart[cf[k] + 1]; q++)
            {
                const int pt = M->facePoints[q];
                if (lastCell[pt] != c) { lastCell[pt] = c; pcCount[pt + 1]++; }
            }
    for (int p = 0; p < nP; p++) pcCount[p + 1] += pcCount[p];
    int* pc = (int*)malloc(sizeof(int)*(pcCount[nP] > 0 ? pcCount[nP] : 1));
    int* pcur = (int*)malloc(sizeof(int)*(nP > 0 ? nP : 1));
    for (int p = 0; p < nP; p++) { pcur[p] = pcCount[p]; lastCell[p] = -1; }
    for (int c = 0; c < nC; c++)
        for (int k = cfs[c]; k < cfs[c + 1]; k++)
            for (int q = M->faceStart[cf[k]]; q < M->faceStart[cf[k] + 1]; q++)
            {
                const int pt = M->facePoints[q];
                if (lastCell[pt] != c) { lastCell[pt] = c; pc[pcur[pt]++] = c; }
            }
    /* point -> boundary faces (ascending face index = insertion order into the hash set) */
    int* pbCount = (int*)calloc(nP + 1, sizeof(int));
    for (int f = nIF; f < nF; f++)
    {
        const int t = bfType[f - nIF];
        if (t == 4) continue;
        for (int q = M->faceStart[f]; q < M->faceStart[f + 1]; q++) pbCount[M->facePoints[q] + 1]++;
    }
    for (int p = 0; p < nP; p++) pbCount[p + 1] += pbCount[p];
    int* pb = (int*)malloc(sizeof(int)*(pbCount[nP] > 0 ? pbCount[nP] : 1));
    for (int p = 0; p < nP; p++) pcur[p] = pbCount[p];
    for (int f = nIF; f < nF; f++)
    {
        if (bfType[f - nIF] == 4) continue;
        for (int q = M->faceStart[f]; q < M->faceStart[f + 1]; q++) pb[pcur[M->facePoints[q]]++] = f;
    }
    /* labelHashSet::toc(): bucket (f & 127) ascending, newest (largest f) first inside a bucket */
    for (int p = 0; p < nP; p++)
    {
        int* a = pb + pbCount[p];
        const int n = pbCount[p + 1] - pbCount[p];
        for (int i = 1; i < n; i++)
        {
            const int v = a[i];
            int j = i - 1;
            while (j >= 0 && (((a[j] & 127) > (v & 127)) || ((a[j] & 127) == (v & 127) && a[j] < v))) { a[j + 1] = a[j]; j--; }
            a[j + 1] = v;
        }
    }
    /* mirror planes: points of empty patches, normal = sum of the unit normals of the patch faces at the point (patch
       face order), / (mag + VSMALL)   [EXT PrimitivePatch::calcPointNormals]; later patches overwrite earlier ones */
    double* mn = (double*)calloc(3*(size_t)(nP > 0 ? nP : 1), sizeof(double));
    L->mirror = (unsigned char*)calloc(nP > 0 ? nP : 1, 1);
    for (int p = 0; p < M->nPatches; p++)
    {
        if (M->patchType[p] != 4) continue;
        for (int i = 0; i < M->patchSize[p]; i++)
        {
            const int f = M->patchStart[p] + i;
            for (int q = M->faceStart[f]; q < M->faceStart[f + 1]; q++) { const int pt = M->facePoints[q]; mn[3*pt] = mn[3*pt + 1] = mn[3*pt + 2] = 0.0; }
        }
        for (int i = 0; i < M->patchSize[p]; i++)
        {
            const int f = M->patchStart[p] + i;
            const double ms = M->magSf[f];
            for (int q = M->faceStart[f]; q < M->faceStart[f + 1]; q++)
            {
                const int pt = M->facePoints[q];
                for (int d = 0; d < 3; d++) mn[3*pt + d] += M->Sf[3*f + d]/ms;
            }
        }
        for (int i = 0; i < M->patchSize[p]; i++)
        {
            const int f = M->patchStart[p] + i;
            for (int q = M->faceStart[f]; q < M->faceStart[f + 1]; q++)
            {
                const int pt = M->facePoints[q];
                if (L->mirror[pt] == (unsigned char)(p + 1)) continue;
                L->mirror[pt] = (unsigned char)(p + 1);
                const double m = mag3(mn + 3*pt) + 1.0e-300;
                for (int d = 0; d < 3; d++) mn[3*pt + d] /= m;
            }
        }
    }
    for (int p = 0; p < nP; p++) L->mirror[p] = (L->mirror[p] && mag3(mn + 3*p) > ORC_SMALL) ? 1 : 0;
    /* rows */
    L->start = (int*)calloc(nP + 1, sizeof(int));
    for (int p = 0; p < nP; p++)
    {
        const int n = (pcCount[p + 1] - pcCount[p]) + (pbCount[p + 1] - pbCount[p]);
        L->start[p + 1] = L->start[p] + (L->mirror[p] ? 2*n : n);
        if (L->start[p + 1] - L->start[p] > L->nSrc) L->nSrc = L->start[p + 1] - L->start[p];
    }
    const int rows = L->start[nP];
    L->entry = (int*)malloc(sizeof(int)*(rows > 0 ? rows : 1));
    L->inv = (double*)calloc(3*(size_t)(rows > 0 ? rows : 1), sizeof(double));
    L->origin = (double*)calloc(3*(size_t)(nP > 0 ? nP : 1), sizeof(double));
    double* X = (double*)malloc(sizeof(double)*3*(size_t)(L->nSrc > 0 ? L->nSrc : 1));
    for (int p = 0; p < nP; p++)
    {
        const int r0 = L->start[p], m = L->start[p + 1] - r0;
        const int n = L->mirror[p] ? m/2 : m;
        int k = 0;
        for (int i = pcCount[p]; i < pcCount[p + 1]; i++, k++)
        {
            L->entry[r0 + k] = pc[i];
            for (int d = 0; d < 3; d++) X[3*k + d] = M->C[3*pc[i] + d];
        }
        for (int i = pbCount[p]; i < pbCount[p + 1]; i++, k++)
        {
            L->entry[r0 + k] = -1 - (pb[i] - nIF);
            for (int d = 0; d < 3; d++) X[3*k + d] = M->Cf[3*pb[i] + d];
        }
        if (L->mirror[p])
        {
            /* allMirrorPoints = p + transform(I - 2*n*n, allPoints - p) */
            const double* nn = mn + 3*p;
            const double* pp = M->points + 3*p;
            double T[9], two[3];
            for (int d = 0; d < 3; d++) two[d] = 2*nn[d];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) T[3*a + b] = (a == b ? 1.0 : 0.0) - two[a]*nn[b];
            for (int i = 0; i < n; i++)
            {
                double dv[3], tv[3];
                for (int d = 0; d < 3; d++) dv[d] = X[3*i + d] - pp[d];
                Tdotv(T, dv, tv);
                for (int d = 0; d < 3; d++) X[3*(n + i) + d] = pp[d] + tv[d];
                L->entry[r0 + n + i] = L->entry[r0 + i];
            }
        }
        /* origin = sum(sqr(W)*X)/sum(sqr(W)), W = 1 */
        double o[3] = {0.0, 0.0, 0.0}, sw = 0.0;
        for (int i = 0; i < m; i++) { for (int d = 0; d < 3; d++) o[d] += 1.0*X[3*i + d]; sw += 1.0; }
        for (int d = 0; d < 3; d++) { o[d] /= sw; L->origin[3*p + d] = o[d]; }
        /* M = (X - o)*W; lsM = M^T M; inv; invLs[i][j] = sum_k inv(i,k)*M[j][k]*W[j] */
        for (int i = 0; i < m; i++) for (int d = 0; d < 3; d++) X[3*i + d] = (X[3*i + d] - o[d])*1.0;
        double ls[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, il[9];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++)
                for (int i = 0; i < m; i++) ls[3*a + b] += X[3*i + a]*X[3*i + b];
        invT(ls, il);
        for (int a = 0; a < 3; a++)
            for (int j = 0; j < m; j++)
            {
                double s = 0.0;
                for (int b = 0; b < 3; b++) s += il[3*a + b]*X[3*j + b]*1.0;
                L->inv[3*(size_t)(r0 + j) + a] = s;
            }
    }
    free(X); free(mn); free(pb); free(pbCount); free(pc); free(pcur); free(pcCount); free(lastCell); free(cur); free(cf); free(cfs); free(bfType);
    return L;
}

int orc_lsq_rows(const orc_lsq* L) { return L->start[L->nPoints]; }
void orc_lsq_export(const orc_lsq* L, int* start, int* entry, double* inv, double* origin, unsigned char* mirror)
{
    const int rows = L->start[L->nPoints];
    memcpy(start, L->start, sizeof(int)*(L->nPoints + 1));
    memcpy(entry, L->entry, sizeof(int)*rows);
    memcpy(inv, L->inv, sizeof(double)*3*(size_t)rows);
    memcpy(origin, L->origin, sizeof(double)*3*(size_t)L->nPoints);
    memcpy(mirror, L->mirror, L->nPoints);
}

/* interpolate(vf, pf): vfI = D [nCells*3], boundary values Db [nBF*3]   leastSquaresVolPointInterpolate.C:89-299 */
void orc_lsq_interpolate(const orc_lsq* L, const orc_mesh* M, const double* D, const double* Db, double* pointD)
{
    double* src = (double*)malloc(sizeof(double)*3*(size_t)(L->nSrc > 0 ? L->nSrc : 1));
    for (int p = 0; p < L->nPoints; p++)
    {
        const int r0 = L->start[p], m = L->start[p + 1] - r0;
        const int n = L->mirror[p] ? m/2 : m;
        double avg[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < n; i++)
        {
            const int e = L->entry[r0 + i];
            const double* v = e >= 0 ? D + 3*(size_t)e : Db + 3*(size_t)(-1 - e);
            for (int d = 0; d < 3; d++) { src[3*i + d] = v[d]; avg[d] += 1.0*v[d]; }
        }
        if (L->mirror[p])
        {
            static const double I9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            double ta[3];
            for (int i = 0; i < n; i++) Tdotv(I9, src + 3*i, src + 3*(n + i));
            Tdotv(I9, avg, ta);
            for (int d = 0; d < 3; d++) avg[d] += ta[d];
        }
        const double sw = (double)m;
        for (int d = 0; d < 3; d++) avg[d] /= sw;
        double co[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int i = 0; i < m; i++) for (int d = 0; d < 3; d++) src[3*i + d] -= avg[d];
        for (int a = 0; a < 3; a++)
            for (int j = 0; j < m; j++)
            {
                const double w = L->inv[3*(size_t)(r0 + j) + a];
                for (int d = 0; d < 3; d++) co[a][d] += w*src[3*j + d];
            }
        const double dr[3] = {M->points[3*p] - L->origin[3*p], M->points[3*p + 1] - L->origin[3*p + 1], M->points[3*p + 2] - L->origin[3*p + 2]};
        for (int d = 0; d < 3; d++) pointD[3*(size_t)p + d] = ((avg[d] + co[0][d]*dr[0]) + co[1][d]*dr[1]) + co[2][d]*dr[2];
    }
    free(src);
}

/* pf.correctBoundaryConditions() at the end of interpolate() (leastSquaresVolPointInterpolate.C:301): evaluate() of the
 * point patch fields of pointD, patches in boundary order, each patch over its meshPoints():
 *   kind 1  symmetryPlane [EXT SymmetryPointPatchField::evaluate]: pf = transform(I - nHat*nHat, pf), nHat = the patch's
 *           pointNormals() [EXT PrimitivePatch: sum of the unit normals of the patch faces at the point / (mag + VSMALL)]
 *   kind 2  fixedValue    [EXT ValuePointPatchField::evaluate]: pf = the patch value
 * (calculated / empty point patches do nothing).  Constrained point i = point[i] owns rows start[i] .. start[i+1], in
 * patch order; vec[row] = nHat or the value.                                                                             */
void orc_point_constraints(int n, const int* point, const int* start, const int* kind, const double* vec, double* pointD)
{
    for (int i = 0; i < n; i++)
    {
        double* v = pointD + 3*(size_t)point[i];
        for (int r = start[i]; r < start[i + 1]; r++)
        {
            const double* a = vec + 3*(size_t)r;
            if (kind[r] == 2) { v[0] = a[0]; v[1] = a[1]; v[2] = a[2]; continue; }
            double T[9], t[3];
            T[0] = 1.0 - a[0]*a[0]; T[1] = -(a[0]*a[1]); T[2] = -(a[0]*a[2]);
            T[3] = -(a[1]*a[0]); T[4] = 1.0 - a[1]*a[1]; T[5] = -(a[1]*a[2]);
            T[6] = -(a[2]*a[0]); T[7] = -(a[2]*a[1]); T[8] = 1.0 - a[2]*a[2];
            Tdotv(T, v, t);
            v[0] = t[0]; v[1] = t[1]; v[2] = t[2];
        }
    }
}

/* ======================================================================================================== */
/*  The other scalar systems of the boundary: TEqn and pEqn                          */
/*    TEqn  solvers/thermalStressFoam/TEqn.H:9-15      rhoC*fvm::ddt(T) == fvm::laplacian(k, T)                */
/*    pEqn  flowModels/consistentIcoFlow/consistentIcoFlow.C:211-224  fvm::laplacian(rAUf, p) == fvc::div(phi) */
/*  [EXT] gaussLaplacianScheme::fvmLaplacianUncorrected, lduMatrix::negSumDiag, EulerDdtScheme::fvmDdt,        */
/*  fvMatrix operator*(volScalarField), operator==, fvc::surfaceIntegrate, fvMatrix::setReference              */
/* ======================================================================================================== */
/*  The explicit correction of `laplacian(gamma,psi) Gauss linear skewCorrected <limitCoeff>` for a scalar                 */
/*  (numerics/skewCorrectedSnGrad/skewCorrectedSnGrad.C:57-294, non-coupled patches): per internal face                   */
/*     kP = Cf - C[own]; kP -= Sf*(Sf & kP)/sqr(magSf); kN likewise                                        (:137-141)    */
/*     corr = ((kN & grad[nei]) - (kP & grad[own]))*deltaCoeffs, grad = the gradScheme named snGradCorr(psi) (:197-222)  */
/*     limiter = min(limitCoeff*mag(orthSnGrad + corr)/((1 - limitCoeff)*mag(corr) + SMALL), 1); corr *= limiter (:255-288)*/
/*  and [EXT] gaussLaplacianScheme::fvmLaplacian: fvm.source() -= V*fvc::div(gammaMagSf*corr).                            */
/*  grad = [EXT] leastSquaresGrad: lsGrad[own] += ownLs[f]*(psi[nei] - psi[own]); lsGrad[nei] -= neiLs[f]*(...), faces    */
/*  ascending, then the patches: lsGrad[faceCells] += patchLs*(psi_b - psi[faceCells]) (psi_b: the fixed value, or the    */
/*  cell value on zeroGradient patches).  The least-squares vectors are mesh data (leastSquaresVectors::pVectors() /      */
/*  nVectors()), an INPUT here as weights and deltaCoeffs are.  Lsrc[c] = 0 - V*div(gms*corr)[c] is returned in `lsrc`.   */
static void scalar_skew_correction(const orc_mesh* M, const unsigned char* emptyFace, const double* gms, const int* bfKind,
                                   const double* bfValue, const double* psi, const double* lsP, const double* lsN,
                                   const double* lsB, double limitCoeff, double* lsrc)
{
    const int nC = M->nCells, nF = M->nFaces, nIF = M->nInternalFaces;
    double* g = (double*)calloc(3*(size_t)(nC > 0 ? nC : 1), sizeof(double));
    double* div = (double*)calloc(nC > 0 ? nC : 1, sizeof(double));
    for (int f = 0; f < nIF; f++)
    {
        const int P = M->owner[f], N = M->neighbour[f];
        const double d = psi[N] - psi[P];
        for (int k = 0; k < 3; k++) { g[3*P + k] += lsP[3*f + k]*d; g[3*N + k] -= lsN[3*f + k]*d; }
    }
    for (int f = nIF; f < nF; f++)
    {
        if (emptyFace[f]) continue;
        const int bf = f - nIF, P = M->owner[f];
        const double pb = bfKind[bf] == 1 ? bfValue[bf] : psi[P];
        const double d = pb - psi[P];
        for (int k = 0; k < 3; k++) g[3*P + k] += lsB[3*bf + k]*d;
    }
    for (int f = 0; f < nIF; f++)
    {
        const int P = M->owner[f], N = M->neighbour[f];
        const double* Sf = M->Sf + 3*f;
        const double m2 = M->magSf[f]*M->magSf[f];
        double kP[3], kN[3];
        for (int k = 0; k < 3; k++) { kP[k] = M->Cf[3*f + k] - M->C[3*P + k]; kN[k] = M->Cf[3*f + k] - M->C[3*N + k]; }
        const double sP = Sf[0]*kP[0] + Sf[1]*kP[1] + Sf[2]*kP[2];
        for (int k = 0; k < 3; k++) kP[k] -= (Sf[k]*sP)/m2;
        const double sN = Sf[0]*kN[0] + Sf[1]*kN[1] + Sf[2]*kN[2];
        for (int k = 0; k < 3; k++) kN[k] -= (Sf[k]*sN)/m2;
        const double a = kN[0]*g[3*N] + kN[1]*g[3*N + 1] + kN[2]*g[3*N + 2];
        const double b = kP[0]*g[3*P] + kP[1]*g[3*P + 1] + kP[2]*g[3*P + 2];
        double corr = (a - b)*M->deltaCoeffs[f];
        const double orth = M->deltaCoeffs[f]*(psi[N] - psi[P]);
        double lim = limitCoeff*fabs(orth + corr)/((1 - limitCoeff)*fabs(corr) + ORC_SMALL);
        if (lim > 1.0) lim = 1.0;
        corr *= lim;
        const double flux = gms[f]*corr;
        div[P] += flux; div[N] -= flux;
    }
    for (int c = 0; c < nC; c++) lsrc[c] = 0.0 - M->V[c]*(div[c]/M->V[c]);
    free(g); free(div);
}

void orc_assemble_scalar(const orc_mesh* M, int sign, const double* gammaCells, const double* gammaFaces, const int* bfKind,
                         const double* bfValue, int ddtScheme, double deltaT, const double* rate, const double* psiOld,
                         const double* phi, int refCell, double refValue,
                         double* upper, double* diag, double* source, double* internalCoeffs, double* boundaryCoeffs)
{
    void orc_assemble_scalar_corrected(const orc_mesh*, int, const double*, const double*, const int*, const double*, int, double,
                                       const double*, const double*, const double*, int, double, const double*, const double*,
                                       const double*, const double*, double, double*, double*, double*, double*, double*);
    orc_assemble_scalar_corrected(M, sign, gammaCells, gammaFaces, bfKind, bfValue, ddtScheme, deltaT, rate, psiOld, phi, refCell,
                                  refValue, NULL, NULL, NULL, NULL, 1.0, upper, diag, source, internalCoeffs, boundaryCoeffs);
}

/* psi != NULL: with the explicit skewCorrected correction (psi = the current field, lsP / lsN [nInternalFaces*3], lsB
 * [nBoundaryFaces*3] the least-squares vectors) */
void orc_assemble_scalar_corrected(const orc_mesh* M, int sign, const double* gammaCells, const double* gammaFaces, const int* bfKind,
                         const double* bfValue, int ddtScheme, double deltaT, const double* rate, const double* psiOld,
                         const double* phi, int refCell, double refValue,
                         const double* psi, const double* lsP, const double* lsN, const double* lsB, double limitCoeff,
                         double* upper, double* diag, double* source, double* internalCoeffs, double* boundaryCoeffs)
{
    const int nC = M->nCells, nF = M->nFaces, nIF = M->nInternalFaces;
    unsigned char* emptyFace = (unsigned char*)calloc(nF > 0 ? nF : 1, 1);
    for (int p = 0; p < M->nPatches; p++)
        if (M->patchType[p] == 4)
            for (int i = 0; i < M->patchSize[p]; i++) emptyFace[M->patchStart[p] + i] = 1;
    double* gms = (double*)calloc(nF > 0 ? nF : 1, sizeof(double));
    double* Ldiag = (double*)calloc(nC > 0 ? nC : 1, sizeof(double));
    double* Lupper = (double*)calloc(nIF > 0 ? nIF : 1, sizeof(double));
    double* div = (double*)calloc(nC > 0 ? nC : 1, sizeof(double));
    for (int f = 0; f < nF; f++)
    {
        if (emptyFace[f]) continue;
        double gf;
        if (gammaFaces) gf = gammaFaces[f];
        else if (f < nIF) { const int P = M->owner[f], N = M->neighbour[f]; gf = M->weights[f]*(gammaCells[P] - gammaCells[N]) + gammaCells[N]; }
        else gf = gammaCells[M->owner[f]];
        gms[f] = gf*M->magSf[f];
    }
    /* fvmLaplacianUncorrected: upper = deltaCoeffs*gammaMagSf; negSumDiag: Diag[l] -= Lower; Diag[u] -= Upper, face by face */
    for (int f = 0; f < nIF; f++)
    {
        Lupper[f] = M->deltaCoeffs[f]*gms[f];
        Ldiag[M->owner[f]] -= Lupper[f];
        Ldiag[M->neighbour[f]] -= Lupper[f];
    }
    /* surfaceIntegrate(phi): internal faces, then patches; then /V */
    if (phi)
    {
        for (int f = 0; f < nIF; f++) { div[M->owner[f]] += phi[f]; div[M->neighbour[f]] -= phi[f]; }
        for (int f = nIF; f < nF; f++) if (!emptyFace[f]) div[M->owner[f]] += phi[f];
        for (int c = 0; c < nC; c++) div[c] /= M->V[c];
    }
    double* lsrc = (double*)calloc(nC > 0 ? nC : 1, sizeof(double));        /* source of the laplacian matrix: 0, or -V*div(correction) */
    if (psi) scalar_skew_correction(M, emptyFace, gms, bfKind, bfValue, psi, lsP, lsN, lsB, limitCoeff, lsrc);
    const double rDeltaT = ddtScheme == 1 ? 1.0/deltaT : 0.0;
    for (int c = 0; c < nC; c++)
    {
        const double V = M->V[c];
        double dA = 0.0, sA = 0.0;
        if (ddtScheme == 1) { dA = (rDeltaT*V)*rate[c]; sA = ((rDeltaT*psiOld[c])*V)*rate[c]; }
        if (sign > 0) { diag[c] = dA - Ldiag[c]; source[c] = sA - lsrc[c]; }
        else { diag[c] = Ldiag[c]; source[c] = phi ? lsrc[c] + V*div[c] : lsrc[c]; }
    }
    free(lsrc);
    for (int f = 0; f < nIF; f++) upper[f] = sign > 0 ? 0.0 - Lupper[f] : Lupper[f];
    for (int f = nIF; f < nF; f++)
    {
        const int bf = f - nIF;
        double Lic = 0.0, Lbc = 0.0;
        if (bfKind[bf] == 1 && !emptyFace[f])
        {
            /* fixedValue: gradientInternalCoeffs = -deltaCoeffs; gradientBoundaryCoeffs = deltaCoeffs*value */
            Lic = gms[f]*(-M->deltaCoeffs[f]);
            Lbc = (-gms[f])*(M->deltaCoeffs[f]*bfValue[bf]);
        }
        internalCoeffs[bf] = sign > 0 ? 0.0 - Lic : Lic;
        boundaryCoeffs[bf] = sign > 0 ? 0.0 - Lbc : Lbc;
    }
    if (refCell >= 0)
    {
        source[refCell] += diag[refCell]*refValue;
        diag[refCell] += diag[refCell];
    }
    free(emptyFace); free(gms); free(Ldiag); free(Lupper); free(div);
}
