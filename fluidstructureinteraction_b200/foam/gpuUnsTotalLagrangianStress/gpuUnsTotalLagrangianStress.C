/*---------------------------------------------------------------------------*\
    gpuUnsTotalLagrangianStress.C -- see the header.  NOT COMPILED here (needs a foam-extend-3.1 tree + the reference's
    libfluidStructureInteraction); host/gpuStressModel.C is the tested twin over the API stand-in.
\*---------------------------------------------------------------------------*/

#include "gpuUnsTotalLagrangianStress.H"
#include "addToRunTimeSelectionTable.H"
#include "tractionDisplacementFvPatchVectorField.H"
#include "fixedDisplacementFvPatchVectorField.H"
#include "symmetryDisplacementFvPatchVectorField.H"
#include "symmetryFvPatchFields.H"
#include "emptyFvPatchFields.H"
#include "fixedValueFvPatchFields.H"
#include "SymmetryPointPatchField.H"
#include "ValuePointPatchField.H"
#include "emptyPolyPatch.H"

namespace Foam
{
namespace stressModels
{

defineTypeNameAndDebug(gpuUnsTotalLagrangianStress, 0);
addToRunTimeSelectionTable(stressModel, gpuUnsTotalLagrangianStress, dictionary);   // cf. unsTotalLagrangianStress.C:53-54


void gpuUnsTotalLagrangianStress::check(const int rc, const char* what) const
{
    if (rc != GPUPCG_OK)
    {
        const char* m = gpupcg_mesh_last_error(gm_);
        FatalErrorIn("gpuUnsTotalLagrangianStress") << what << " failed (" << rc << "): "
            << ((m && *m) ? m : gpupcg_last_error(solver_)) << abort(FatalError);
    }
}


label gpuUnsTotalLagrangianStress::patchTypeCode(const fvPatchVectorField& pf) const
{
    if (isA<tractionDisplacementFvPatchVectorField>(pf)) return 0;
    if (isA<fixedDisplacementFvPatchVectorField>(pf)) return 1;
    if (isA<symmetryDisplacementFvPatchVectorField>(pf)) return 2;
    if (isA<symmetryFvPatchVectorField>(pf)) return 3;
    if (isA<emptyFvPatchVectorField>(pf)) return 4;
    if (isA<fixedValueFvPatchVectorField>(pf)) return 5;
    FatalErrorIn("gpuUnsTotalLagrangianStress") << "patch field " << pf.type() << " on " << pf.patch().name()
        << " is not available on the GPU" << abort(FatalError);
    return -1;
}


void gpuUnsTotalLagrangianStress::createDeviceCase()
{
    const fvMesh& mesh = this->mesh();
    const label nC = mesh.nCells(), nF = mesh.nFaces(), nIF = mesh.nInternalFaces(), nBF = nF - nIF;
    if (Pstream::parRun())
    {
        FatalErrorIn("gpuUnsTotalLagrangianStress") << "decomposed cases: the device-resident assembly has no processor patches yet; "
            << "use the stock model with `solver gpuPCG;`" << abort(FatalError);
    }

    // geometry in face order (internal faces, then the patches as they lie in the polyMesh)
    vectorField Sf(nF), Cf(nF); scalarField magSf(nF), dc(nF);
    SubField<vector>(Sf, nIF) = mesh.Sf().internalField(); SubField<vector>(Cf, nIF) = mesh.Cf().internalField();
    SubField<scalar>(magSf, nIF) = mesh.magSf().internalField(); SubField<scalar>(dc, nIF) = mesh.deltaCoeffs().internalField();
    labelList pStart(mesh.boundary().size()), pSize(mesh.boundary().size()), pType(mesh.boundary().size());
    forAll(mesh.boundary(), patchI)
    {
        const fvPatch& p = mesh.boundary()[patchI];
        const polyPatch& pp = p.patch();
        pStart[patchI] = pp.start(); pSize[patchI] = pp.size();
        pType[patchI] = patchTypeCode(D().boundaryField()[patchI]);
        if (!isA<emptyPolyPatch>(pp))
        {
            SubField<vector>(Sf, p.size(), pp.start()) = mesh.Sf().boundaryField()[patchI];
            SubField<vector>(Cf, p.size(), pp.start()) = mesh.Cf().boundaryField()[patchI];
            SubField<scalar>(magSf, p.size(), pp.start()) = mesh.magSf().boundaryField()[patchI];
            SubField<scalar>(dc, p.size(), pp.start()) = mesh.deltaCoeffs().boundaryField()[patchI];
        }
        else
        {
            // fvPatch of an empty patch has no faces: the device wants the polyMesh geometry of those faces (fan loops of fvc::grad)
            SubField<vector>(Sf, pp.size(), pp.start()) = pp.faceAreas();
            SubField<vector>(Cf, pp.size(), pp.start()) = pp.faceCentres();
            SubField<scalar>(magSf, pp.size(), pp.start()) = mag(pp.faceAreas());
            SubField<scalar>(dc, pp.size(), pp.start()) = 1.0/mag(pp.faceCentres() - vectorField(mesh.C().internalField(), pp.faceCells()));
        }
    }
    check
    (
        gpupcg_mesh_create
        (
            &gm_, 0 /* device: serial run (decomposed cases are refused above) */, nC, nF, nIF,
            mesh.faceOwner().begin(), mesh.faceNeighbour().begin(), pStart.size(), pStart.begin(), pSize.begin(), pType.begin(),
            reinterpret_cast<const double*>(Sf.begin()), magSf.begin(), reinterpret_cast<const double*>(Cf.begin()),
            reinterpret_cast<const double*>(mesh.C().internalField().begin()), mesh.V().field().begin(),
            mesh.weights().internalField().begin(), dc.begin()
        ),
        "gpupcg_mesh_create"
    );

    // faces as CSR
    labelList faceStart(nF + 1, 0);
    forAll(mesh.faces(), faceI) faceStart[faceI + 1] = faceStart[faceI] + mesh.faces()[faceI].size();
    labelList facePoints(faceStart[nF]);
    forAll(mesh.faces(), faceI) forAll(mesh.faces()[faceI], pI) facePoints[faceStart[faceI] + pI] = mesh.faces()[faceI][pI];
    check(gpupcg_mesh_set_points(gm_, mesh.nPoints(), reinterpret_cast<const double*>(mesh.points().begin()), faceStart.begin(),
                                 facePoints.begin()), "gpupcg_mesh_set_points");

    // material
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_MU, rheology().mu()().internalField().begin()), "mu");
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_LAMBDA, rheology().lambda()().internalField().begin()), "lambda");
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_RHO, rheology().rho()().internalField().begin()), "rho");

    // fields: D and its old levels, gradients, boundary loads
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_D, reinterpret_cast<const double*>(D().internalField().begin())), "D");
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_DOLD, reinterpret_cast<const double*>(D().oldTime().internalField().begin())), "D0");
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_DOLDOLD, reinterpret_cast<const double*>(D().oldTime().oldTime().internalField().begin())), "D00");
    {
        scalarField z(9*max(nC, nF), 0.0);
        check(gpupcg_mesh_set_field(gm_, GPUPCG_F_GRADD, z.begin()), "gradD");
        check(gpupcg_mesh_set_field(gm_, GPUPCG_F_GRADDF, z.begin()), "gradDf");
        check(gpupcg_mesh_set_field(gm_, GPUPCG_F_PATCHGRADIENT, z.begin()), "patchGradient");
    }
    vectorField fixedValue(nBF, vector::zero);
    forAll(D().boundaryField(), patchI)
    {
        if (pType[patchI] == 1 || pType[patchI] == 5)
        {
            SubField<vector>(fixedValue, pSize[patchI], pStart[patchI] - nIF) = D().boundaryField()[patchI];
        }
    }
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_FIXEDVALUE, reinterpret_cast<const double*>(fixedValue.begin())), "fixedValue");
    uploadPatchLoads();

    // volToPoint_: stencils (cells of the point, then its boundary faces), inverse least-squares matrices column by column,
    // origins, mirror planes (leastSquaresVolPointInterpolation.C: pointBndFaces :54-205, origins :1138-1421, invLsMatrices
    // :1425-1826, mirrorPlaneTransformation :1829-1940)
    {
        const leastSquaresVolPointInterpolation& v2p = volToPoint();
        const labelListList& pointCells = mesh.pointCells();
        const labelListList& pointBndFaces = v2p.pointBndFaces();
        const PtrList<scalarRectangularMatrix>& invLs = v2p.invLsMatrices();
        labelList start(mesh.nPoints() + 1, 0);
        forAll(pointCells, pI) start[pI + 1] = start[pI] + invLs[pI].m();
        labelList entry(start[mesh.nPoints()]);
        scalarField inv(3*start[mesh.nPoints()]);
        List<unsigned char> mirror(mesh.nPoints(), 0);
        forAll(pointCells, pI)
        {
            const label nReal = pointCells[pI].size() + pointBndFaces[pI].size();
            const label m = invLs[pI].m();
            mirror[pI] = (m == 2*nReal) ? 1 : 0;
            for (label j = 0; j < m; j++)
            {
                const label r = j % nReal;
                entry[start[pI] + j] = (r < pointCells[pI].size()) ? pointCells[pI][r]
                                                                   : -1 - (pointBndFaces[pI][r - pointCells[pI].size()] - nIF);
                for (label i = 0; i < 3; i++) inv[3*(start[pI] + j) + i] = invLs[pI][i][j];
            }
        }
        check(gpupcg_mesh_set_lsq(gm_, start.begin(), entry.begin(), inv.begin(),
                                  reinterpret_cast<const double*>(v2p.origins().begin()), mirror.begin()), "gpupcg_mesh_set_lsq");
    }

    // the point patch fields of pointD that act in correctBoundaryConditions(): symmetryPlane, fixedValue
    {
        const pointVectorField& pD = pointD();
        DynamicList<label> pts; DynamicList<label> kinds; DynamicList<vector> vecs;
        forAll(pD.boundaryField(), patchI)
        {
            const pointPatchVectorField& ppf = pD.boundaryField()[patchI];
            const labelList& mp = ppf.patch().meshPoints();
            const bool sym = ppf.type() == "symmetryPlane", fix = ppf.type() == "fixedValue";
            if (!sym && !fix) continue;
            const vectorField& nHat = ppf.patch().pointNormals();
            forAll(mp, i)
            {
                pts.append(mp[i]); kinds.append(sym ? 1 : 2);
                vecs.append(sym ? nHat[i] : ppf.patchInternalField()()[i]);
            }
        }
        // group by point, patch order preserved
        labelList order; sortedOrder(labelList(pts), order);        // stable
        labelList cPoint, cStart(1, 0), cKind; vectorField cVec;
        DynamicList<label> dPoint, dStart, dKind; DynamicList<vector> dVec;
        dStart.append(0);
        forAll(order, k)
        {
            const label i = order[k];
            if (dPoint.size() == 0 || dPoint[dPoint.size() - 1] != pts[i]) { if (dPoint.size()) dStart.append(dKind.size()); dPoint.append(pts[i]); }
            dKind.append(kinds[i]); dVec.append(vecs[i]);
        }
        dStart.append(dKind.size());
        cPoint = dPoint; cStart = dStart; cKind = dKind; cVec = vectorField(dVec);
        check(gpupcg_mesh_set_point_constraints(gm_, cPoint.size(), cPoint.begin(), cStart.begin(), cKind.begin(),
                                                reinterpret_cast<const double*>(cVec.begin())), "gpupcg_mesh_set_point_constraints");
    }
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_POINTD, reinterpret_cast<const double*>(pointD().internalField().begin())), "pointD");
    {
        vectorField Db(nBF, vector::zero);
        check(gpupcg_mesh_set_field(gm_, GPUPCG_F_DB, reinterpret_cast<const double*>(Db.begin())), "Db");
        check(gpupcg_patch_evaluate(gm_), "gpupcg_patch_evaluate");
    }
    check(gpupcg_create(&solver_, 0, nC, nIF, mesh.faceOwner().begin(), mesh.faceNeighbour().begin(), 0, NULL, NULL, NULL), "gpupcg_create");
}


void gpuUnsTotalLagrangianStress::uploadPatchLoads()
{
    const fvMesh& mesh = this->mesh();
    const label nIF = mesh.nInternalFaces(), nBF = mesh.nFaces() - nIF;
    vectorField traction(nBF, vector::zero); scalarField pressure(nBF, 0.0);
    forAll(D().boundaryField(), patchI)
    {
        if (isA<tractionDisplacementFvPatchVectorField>(D().boundaryField()[patchI]))
        {
            const tractionDisplacementFvPatchVectorField& p =
                refCast<const tractionDisplacementFvPatchVectorField>(D().boundaryField()[patchI]);
            const label s = p.patch().patch().start() - nIF;
            SubField<vector>(traction, p.size(), s) = p.traction();
            SubField<scalar>(pressure, p.size(), s) = p.pressure();
        }
    }
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_TRACTION, reinterpret_cast<const double*>(traction.begin())), "traction");
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_PRESSURE, pressure.begin()), "pressure");
}


gpuUnsTotalLagrangianStress::gpuUnsTotalLagrangianStress(const fvMesh& mesh)
:
    unsTotalLagrangianStress(mesh),
    gm_(NULL),
    solver_(NULL),
    timeIndex_(-1)
{}


gpuUnsTotalLagrangianStress::~gpuUnsTotalLagrangianStress()
{
    if (solver_) gpupcg_destroy(solver_);
    if (gm_) gpupcg_mesh_destroy(gm_);
}


bool gpuUnsTotalLagrangianStress::evolve()
{
    const fvMesh& mesh = this->mesh();
    if (!gm_) createDeviceCase();               // first call: everything the constructor of the stock model has read
    if (timeIndex_ != runTime().timeIndex())
    {
        if (timeIndex_ >= 0) check(gpupcg_mesh_new_time_step(gm_), "gpupcg_mesh_new_time_step");
        timeIndex_ = runTime().timeIndex();
    }
    uploadPatchLoads();                         // setTraction / setPressure / function objects may have changed them

    // --- controls: stressProperties (:773-810), fvSolution solvers.D + relaxationFactors.D, fvSchemes
    gpupcg_evolve_controls c;
    memset(&c, 0, sizeof(c));
    c.nCorrectors = readInt(stressProperties().lookup("nCorrectors"));
    c.convergenceTolerance = readScalar(stressProperties().lookup("convergenceTolerance"));
    c.relConvergenceTolerance = stressProperties().lookupOrDefault<scalar>("relConvergenceTolerance", 0);
    const dictionary& sd = mesh.solutionDict().solverDict("D");
    if (word(sd.lookup("solver")) != "PCG" && word(sd.lookup("solver")) != "gpuPCG")
    {
        FatalErrorIn("gpuUnsTotalLagrangianStress::evolve()") << "solvers.D.solver must be PCG or gpuPCG (preconditioner DIC)" << abort(FatalError);
    }
    c.tolerance = sd.lookupOrDefault<scalar>("tolerance", 1e-6); c.relTol = sd.lookupOrDefault<scalar>("relTol", 0);
    c.minIter = sd.lookupOrDefault<label>("minIter", 0); c.maxIter = sd.lookupOrDefault<label>("maxIter", 1000);
    c.doRelax = mesh.solutionDict().relax("D") ? 1 : 0;
    c.relaxD = c.doRelax ? mesh.solutionDict().relaxationFactor("D") : 1.0;
    const Vector<label>& sD = mesh.solutionD();
    for (direction d = 0; d < 3; d++) c.validCmpt[d] = (sD[d] != -1) ? 1 : 0;
    if (stressProperties().found("g"))
    {
        const dimensionedVector g(stressProperties().lookup("g"));
        for (direction d = 0; d < 3; d++) c.g[d] = g.value()[d];
    }
    const word d2 = word(mesh.schemesDict().d2dt2Scheme("d2dt2(D)")), d1 = word(mesh.schemesDict().ddtScheme("ddt(D)"));
    if ((d2 != "steadyState" && d2 != "Euler" && d2 != "backward") || (d1 != "steadyState" && d1 != "Euler" && d1 != "backward"))
    {
        FatalErrorIn("gpuUnsTotalLagrangianStress::evolve()") << "d2dt2(D) " << d2 << " / ddt(D) " << d1
            << ": steadyState, Euler and backward are available on the GPU" << abort(FatalError);
    }
    c.eqn.d2dt2Scheme = (d2 == "Euler") ? 1 : (d2 == "backward") ? 2 : 0;
    c.eqn.ddtScheme = (d1 == "Euler") ? 1 : (d1 == "backward") ? 2 : 0;
    c.eqn.K = stressProperties().found("K") ? dimensionedScalar(stressProperties().lookup("K")).value() : 0.0;
    c.eqn.deltaT = runTime().deltaT().value(); c.eqn.deltaT0 = runTime().deltaT0().value();
    if (c.eqn.d2dt2Scheme == 2 || c.eqn.ddtScheme == 2)
    {
        // `backward`: the schemes decide from the state of D's old-time levels whether a use sees deltaT0 or GREAT
        // (backwardD2dt2Scheme.C:58-76; backwardDdtScheme::deltaT0_).  Here D_ IS a GeometricField, so the state is read off
        // it in the order evolve()'s expression touches it (the fvMatrix constructor's non-const access first), and the four
        // levels go to the device as they are.
        volVectorField& Dref = D();
        Dref.boundaryField();                                       // non-const access: storeOldTimes()
        const scalar dT0 = c.eqn.deltaT0;
        c.eqn.eD2 = c.eqn.eDdtD = c.eqn.eDdtD0 = c.eqn.eDdtD00 = c.eqn.eDamp = c.eqn.eU = dT0;
        if (c.eqn.d2dt2Scheme == 2)
        {
            const volVectorField& D000 = Dref.oldTime().oldTime().oldTime();
            c.eqn.eD2 = (Dref.oldTime().oldTime().timeIndex() == D000.timeIndex()) ? GREAT : dT0;
            c.eqn.eDdtD = (Dref.nOldTimes() < 2) ? GREAT : dT0;
            c.eqn.eDdtD0 = (Dref.oldTime().nOldTimes() < 2) ? GREAT : dT0;
            c.eqn.eDdtD00 = (Dref.oldTime().oldTime().nOldTimes() < 2) ? GREAT : dT0;
            const volVectorField& D0000 = D000.oldTime();
            check(gpupcg_mesh_set_field(gm_, GPUPCG_F_DOOO, reinterpret_cast<const double*>(D000.internalField().begin())), "D000");
            check(gpupcg_mesh_set_field(gm_, GPUPCG_F_DOOOO, reinterpret_cast<const double*>(D0000.internalField().begin())), "D0000");
        }
        if (c.eqn.K > SMALL && c.eqn.ddtScheme == 2) c.eqn.eDamp = (Dref.nOldTimes() < 2) ? GREAT : dT0;
        check(gpupcg_mesh_set_field(gm_, GPUPCG_F_DOLD, reinterpret_cast<const double*>(Dref.oldTime().internalField().begin())), "D0");
        check(gpupcg_mesh_set_field(gm_, GPUPCG_F_DOLDOLD, reinterpret_cast<const double*>(Dref.oldTime().oldTime().internalField().begin())), "D00");
        c.eqn.eU = dT0;                                             // two old levels exist from here on
    }
    c.eqn.nonLinear = Switch(stressProperties().lookupOrDefault<Switch>("nonLinear", false)) ? 1 : 0;
    c.eqn.thermal = thermalStress() ? 1 : 0;
    {
        ITstream& is = mesh.schemesDict().snGradScheme("snGrad(D)");
        const word name(is);
        if (name != "skewCorrected")
        {
            FatalErrorIn("gpuUnsTotalLagrangianStress::evolve()") << "snGrad(D) " << name << ": skewCorrected <limitCoeff> only" << abort(FatalError);
        }
        c.eqn.limitCoeff = readScalar(is);
    }
    if (c.eqn.thermal)
    {
        FatalErrorIn("gpuUnsTotalLagrangianStress::evolve()") << "thermal stress: upload threeK / alpha / DT / DTb "
            << "(GPUPCG_F_THREEK ...) from the thermal model before evolve(); not wired in this class yet" << abort(FatalError);
    }

    gpupcg_evolve_result r;
    check(gpupcg_evolve_D(solver_, gm_, &c, &r, NULL, NULL, 0), "gpupcg_evolve_D");
    check(gpupcg_sigma_resident(gm_, &c.eqn), "gpupcg_sigma_resident");

    // --- back into the inherited fields (what solvers write, the FSI interface and function objects read)
    check(gpupcg_mesh_get_field(gm_, GPUPCG_F_D, reinterpret_cast<double*>(D().internalField().begin())), "D");
    D().correctBoundaryConditions();
    check(gpupcg_mesh_get_field(gm_, GPUPCG_F_POINTD,
          reinterpret_cast<double*>(const_cast<pointVectorField&>(pointD()).internalField().begin())), "pointD");
    check(gpupcg_mesh_get_field(gm_, GPUPCG_F_SIGMA,
          reinterpret_cast<double*>(const_cast<volSymmTensorField&>(sigma()).internalField().begin())), "sigma");
    check(gpupcg_mesh_get_field(gm_, GPUPCG_F_EPSILON,
          reinterpret_cast<double*>(const_cast<volSymmTensorField&>(epsilon()).internalField().begin())), "epsilon");
    {
        volTensorField& gradD = const_cast<volTensorField&>(mesh.lookupObject<volTensorField>("grad(D)"));
        check(gpupcg_mesh_get_field(gm_, GPUPCG_F_GRADD, reinterpret_cast<double*>(gradD.internalField().begin())), "grad(D)");
        volVectorField& U = const_cast<volVectorField&>(mesh.lookupObject<volVectorField>("U"));
        check(gpupcg_mesh_get_field(gm_, GPUPCG_F_U, reinterpret_cast<double*>(U.internalField().begin())), "U");
    }

    Info<< "Solving for " << D().name() << ", Initial residual = " << r.initialResidual
        << ", Final residual = " << r.lastSolverInitialResidual << ", No outer iterations = " << r.nOuterIterations << endl;

    return r.converged != 0;
}

} // End namespace stressModels
} // End namespace Foam

// ************************************************************************* //
