#include "gpuStressModel.H"

#include <cmath>
#include <cstdlib>
#include <sstream>

namespace Foam
{

// * * * * * * * * * * * * * * * * stressModel  * * * * * * * * * * * * * * * //

std::map<word, stressModel::dictionaryConstructor>& stressModel::dictionaryConstructorTable()
{
    static std::map<word, dictionaryConstructor> t;
    return t;
}

stressModel::stressModel(const word& type, const fvMeshLike& mesh)
:
    mesh_(mesh),
    stressProperties_(mesh.stressProperties.isDict(type + "Coeffs") ? mesh.stressProperties.subDict(type + "Coeffs") : dictionary())
{}

stressModel* stressModel::New(const fvMeshLike& mesh)
{
    const word stressModelTypeName(mesh.stressProperties.lookup("stressModel"));
    std::map<word, dictionaryConstructor>::const_iterator it = dictionaryConstructorTable().find(stressModelTypeName);
    if (it == dictionaryConstructorTable().end())
    {
        std::ostringstream os;
        for (std::map<word, dictionaryConstructor>::const_iterator k = dictionaryConstructorTable().begin();
             k != dictionaryConstructorTable().end(); ++k) os << " " << k->first;
        FatalErrorIn("stressModel::New(const fvMesh&)")
            << "Unknown stressModel type " << stressModelTypeName << "\n\nValid stressModel types are :\n(" << os.str() << " )"
            << exit(FatalError);
    }
    return it->second(mesh);
}

// * * * * * * * * * * * * * * gpuUnsTotalLagrangianStress  * * * * * * * * * * * * * * //

const word gpuUnsTotalLagrangianStress::typeName("gpuUnsTotalLagrangianStress");

static stressModel* newGpuUnsTotalLagrangianStress(const fvMeshLike& mesh) { return new gpuUnsTotalLagrangianStress(mesh); }
// addToRunTimeSelectionTable(stressModel, gpuUnsTotalLagrangianStress, dictionary)   (unsTotalLagrangianStress.C:53-54)
static stressModel::addToTable addgpuUnsTotalLagrangianStressDictionaryConstructorToTable_
(
    gpuUnsTotalLagrangianStress::typeName, newGpuUnsTotalLagrangianStress
);

static scalar dimensionedValue(const dictionary& d, const word& key)
{
    // "rho [1 -3 0 0 0 0 0] 7854" -> 7854  (dimensionedScalar: name, dimensions, value); a bare number passes
    const std::string v = d.lookup(key);
    const size_t p = v.rfind(']');
    return atof((p == std::string::npos ? v : v.substr(p + 1)).c_str());
}

static bool switchValue(const std::string& v)
{
    return v == "yes" || v == "on" || v == "true" || v == "y" || v == "t";
}

static int patchTypeCode(const word& t)
{
    if (t == "tractionDisplacement") return 0;
    if (t == "fixedDisplacement") return 1;
    if (t == "symmetryDisplacement") return 2;
    if (t == "symmetryPlane") return 3;
    if (t == "empty") return 4;
    if (t == "fixedValue") return 5;
    FatalErrorIn("gpuUnsTotalLagrangianStress::gpuUnsTotalLagrangianStress(const fvMesh&)")
        << "boundary condition " << t << " of D is not available on the GPU (supported: tractionDisplacement, "
        << "fixedDisplacement, symmetryDisplacement, symmetryPlane, empty, fixedValue)" << abort(FatalError);
    return -1;
}

static int schemeCode(const dictionary& fvSchemes, const word& group, const word& key)
{
    const dictionary& g = fvSchemes.subDict(group);
    const std::string s = g.found(key) ? g.lookup(key) : g.lookup("default");
    if (s == "steadyState") return 0;
    if (s == "Euler") return 1;
    if (s == "backward") return 2;
    FatalErrorIn("gpuUnsTotalLagrangianStress::gpuUnsTotalLagrangianStress(const fvMesh&)")
        << group << " " << key << " " << s << " is not available on the GPU (steadyState, Euler, backward)" << abort(FatalError);
    return -1;
}

void gpuUnsTotalLagrangianStress::check(int rc, const char* what) const
{
    if (rc != 0)
    {
        const char* m1 = gpupcg_mesh_last_error(gm_);       // NULL handle: the creation error
        const char* m2 = gpupcg_last_error(solver_);
        FatalErrorIn("gpuUnsTotalLagrangianStress") << what << " failed (" << rc << "): " << (m1 && *m1 ? m1 : (m2 ? m2 : ""))
                                                    << abort(FatalError);
    }
}

gpuUnsTotalLagrangianStress::gpuUnsTotalLagrangianStress(const fvMeshLike& mesh)
:
    stressModel(typeName, mesh), gm_(0), solver_(0), nonLinear_(false), thermal_(false), tiD_(1), now_(1)
{
    const dictionary& sp = stressProperties();
    const int nC = mesh.nCells, nIF = mesh.nInternalFaces, nBF = mesh.nFaces - nIF;

    // --- stressProperties (unsTotalLagrangianStress.C:773-810)
    ctl_ = gpupcg_evolve_controls();
    ctl_.nCorrectors = atoi(std::string(sp.lookup("nCorrectors")).c_str());
    ctl_.convergenceTolerance = atof(std::string(sp.lookup("convergenceTolerance")).c_str());
    ctl_.relConvergenceTolerance = sp.found("relConvergenceTolerance") ? atof(std::string(sp.lookup("relConvergenceTolerance")).c_str()) : 0.0;
    nonLinear_ = sp.found("nonLinear") && switchValue(sp.lookup("nonLinear"));
    ctl_.eqn.nonLinear = nonLinear_ ? 1 : 0;
    ctl_.eqn.K = sp.found("K") ? dimensionedValue(sp, "K") : 0.0;
    if (sp.found("g")) FatalErrorIn("gpuUnsTotalLagrangianStress") << "keyword g: pass the body force through fvMeshLike" << abort(FatalError);

    // --- fvSolution: solvers { D { solver PCG|gpuPCG; preconditioner DIC; ... } }  relaxationFactors { D ...; }
    const dictionary& sd = mesh.fvSolution.subDict("solvers").subDict("D");
    const word solverName(sd.lookup("solver"));
    if (solverName != "PCG" && solverName != "gpuPCG")
        FatalErrorIn("gpuUnsTotalLagrangianStress") << "solvers.D.solver " << solverName << ": the resident loop solves with PCG/DIC"
                                                    << abort(FatalError);
    if (sd.found("preconditioner") && word(sd.lookup("preconditioner")) != "DIC")
        FatalErrorIn("gpuUnsTotalLagrangianStress") << "solvers.D.preconditioner must be DIC" << abort(FatalError);
    ctl_.tolerance = 1e-6; ctl_.relTol = 0; ctl_.minIter = 0; ctl_.maxIter = 1000;     // lduMatrix::solver::readControls defaults
    sd.readIfPresent("tolerance", ctl_.tolerance); sd.readIfPresent("relTol", ctl_.relTol);
    sd.readIfPresent("minIter", ctl_.minIter); sd.readIfPresent("maxIter", ctl_.maxIter);
    ctl_.doRelax = 0; ctl_.relaxD = 1.0;
    if (mesh.fvSolution.isDict("relaxationFactors") && mesh.fvSolution.subDict("relaxationFactors").found("D"))
    {
        ctl_.doRelax = 1;
        mesh.fvSolution.subDict("relaxationFactors").readIfPresent("D", ctl_.relaxD);
    }

    // --- fvSchemes: d2dt2(D), ddt(D), snGrad(D) skewCorrected <limitCoeff>
    ctl_.eqn.d2dt2Scheme = schemeCode(mesh.fvSchemes, "d2dt2Schemes", "d2dt2(D)");
    ctl_.eqn.ddtScheme = schemeCode(mesh.fvSchemes, "ddtSchemes", "ddt(D)");
    {
        const std::string sn = mesh.fvSchemes.subDict("snGradSchemes").lookup("snGrad(D)");
        std::istringstream is(sn);
        word name; is >> name;
        ctl_.eqn.limitCoeff = 1.0;
        if (name == "skewCorrected") is >> ctl_.eqn.limitCoeff;
        else
            FatalErrorIn("gpuUnsTotalLagrangianStress") << "snGrad(D) " << sn << ": only `skewCorrected <limitCoeff>` "
                << "(numerics/skewCorrectedVectorSnGrad) is available on the GPU" << abort(FatalError);
    }
    ctl_.eqn.deltaT = ctl_.eqn.deltaT0 = mesh.deltaT;
    ctl_.eqn.thermal = 0;
    for (int d = 0; d < 3; d++) { ctl_.validCmpt[d] = mesh.solutionD[d] != -1 ? 1 : 0; ctl_.g[d] = 0.0; }

    // --- rheologyProperties (constitutiveModel.C:193-303: linearElastic, plane stress / strain)
    const dictionary& rh = mesh.rheologyProperties.subDict("rheology");
    if (word(rh.lookup("type")) != "linearElastic")
        FatalErrorIn("gpuUnsTotalLagrangianStress") << "rheology type " << rh.lookup("type") << " (linearElastic only)" << abort(FatalError);
    const scalar E = dimensionedValue(rh, "E"), nu = dimensionedValue(rh, "nu"), rho = dimensionedValue(rh, "rho");
    const bool planeStress = switchValue(mesh.rheologyProperties.lookup("planeStress"));
    const scalar mu = E/(2.0*(1.0 + nu));
    const scalar lambda = planeStress ? nu*E/((1.0 + nu)*(1.0 - nu)) : nu*E/((1.0 + nu)*(1.0 - 2.0*nu));

    // --- the mesh and the fields on the device
    std::vector<int> ptype(mesh.nPatches);
    for (int p = 0; p < mesh.nPatches; p++) ptype[p] = patchTypeCode(mesh.patchFieldTypes[p]);
    check(gpupcg_mesh_create(&gm_, mesh.device, nC, mesh.nFaces, nIF, mesh.owner, mesh.neighbour, mesh.nPatches, mesh.patchStart,
                             mesh.patchSize, ptype.data(), mesh.Sf, mesh.magSf, mesh.Cf, mesh.C, mesh.V, mesh.weights,
                             mesh.deltaCoeffs), "gpupcg_mesh_create");
    check(gpupcg_mesh_set_points(gm_, mesh.nPoints, mesh.points, mesh.faceStart, mesh.facePoints), "gpupcg_mesh_set_points");
    std::vector<scalar> cellField(nC);
    const struct { int id; scalar v; } cf[3] = {{GPUPCG_F_MU, mu}, {GPUPCG_F_LAMBDA, lambda}, {GPUPCG_F_RHO, rho}};
    for (int k = 0; k < 3; k++)
    {
        for (int c = 0; c < nC; c++) cellField[c] = cf[k].v;
        check(gpupcg_mesh_set_field(gm_, cf[k].id, cellField.data()), "gpupcg_mesh_set_field");
    }
    traction_.assign(mesh.traction, mesh.traction + 3*(size_t)nBF);
    pressure_.assign(mesh.pressure, mesh.pressure + nBF);
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_TRACTION, traction_.data()), "set traction");
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_PRESSURE, pressure_.data()), "set pressure");
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_FIXEDVALUE, mesh.fixedValue), "set fixedValue");
    const struct { int id; size_t n; } zeros[] = {{GPUPCG_F_D, 3*(size_t)nC}, {GPUPCG_F_DOLD, 3*(size_t)nC}, {GPUPCG_F_DOLDOLD, 3*(size_t)nC},
                                                 {GPUPCG_F_GRADD, 9*(size_t)nC}, {GPUPCG_F_GRADDF, 9*(size_t)mesh.nFaces},
                                                 {GPUPCG_F_PATCHGRADIENT, 3*(size_t)nBF}, {GPUPCG_F_POINTD, 3*(size_t)mesh.nPoints},
                                                 {GPUPCG_F_DB, 3*(size_t)nBF}};
    std::vector<scalar> z;
    for (size_t k = 0; k < sizeof(zeros)/sizeof(zeros[0]); k++)
    {
        if (zeros[k].id == GPUPCG_F_POINTD)
        {
            // volToPoint_ before pointD: the device sizes the point fields from it
            check(gpupcg_mesh_set_lsq(gm_, mesh.lsqStart, mesh.lsqEntry, mesh.lsqInv, mesh.lsqOrigin, mesh.lsqMirror), "gpupcg_mesh_set_lsq");
            check(gpupcg_mesh_set_point_constraints(gm_, mesh.nPointConstraints, mesh.pcPoint, mesh.pcStart, mesh.pcKind, mesh.pcVec),
                  "gpupcg_mesh_set_point_constraints");
        }
        z.assign(zeros[k].n > 0 ? zeros[k].n : 1, 0.0);
        check(gpupcg_mesh_set_field(gm_, zeros[k].id, z.data()), "gpupcg_mesh_set_field(0)");
    }
    check(gpupcg_patch_evaluate(gm_), "gpupcg_patch_evaluate");            // D.boundaryField() of the initial field
    check(gpupcg_create(&solver_, mesh.device, nC, nIF, mesh.owner, mesh.neighbour, 0, 0, 0, 0), "gpupcg_create");
}

gpuUnsTotalLagrangianStress::~gpuUnsTotalLagrangianStress()
{
    if (solver_) gpupcg_destroy(solver_);
    if (gm_) gpupcg_mesh_destroy(gm_);
}

void gpuUnsTotalLagrangianStress::setTraction(const label patchID, const scalar* traction)
{
    const int b0 = mesh_.patchStart[patchID] - mesh_.nInternalFaces, n = mesh_.patchSize[patchID];
    if (mesh_.patchFieldTypes[patchID] != "tractionDisplacement")
        FatalErrorIn("gpuUnsTotalLagrangianStress::setTraction(...)") << "Bounary condition on " << mesh_.patchNames[patchID]
            << " is " << mesh_.patchFieldTypes[patchID] << "for patch" << mesh_.patchNames[patchID]
            << ", instead tractionDisplacement" << abort(FatalError);          // unsTotalLagrangianStress.C:528-541
    for (int i = 0; i < 3*n; i++) traction_[3*(size_t)b0 + i] = traction[i];
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_TRACTION, traction_.data()), "set traction");
}

void gpuUnsTotalLagrangianStress::setPressure(const label patchID, const scalar* pressure)
{
    const int b0 = mesh_.patchStart[patchID] - mesh_.nInternalFaces, n = mesh_.patchSize[patchID];
    if (mesh_.patchFieldTypes[patchID] != "tractionDisplacement")
        FatalErrorIn("gpuUnsTotalLagrangianStress::setPressure(...)") << "Bounary condition on " << mesh_.patchNames[patchID]
            << " is " << mesh_.patchFieldTypes[patchID] << ", instead tractionDisplacement" << abort(FatalError);
    for (int i = 0; i < n; i++) pressure_[(size_t)b0 + i] = pressure[i];
    check(gpupcg_mesh_set_field(gm_, GPUPCG_F_PRESSURE, pressure_.data()), "set pressure");
}

// ---- D's old-time levels for the `backward` schemes.  In a foam-extend process GeometricField keeps them (oldTime() creates a
// level on first use, storeOldTimes() shifts them at the first non-const access of a new time step) and the schemes read their
// state: backwardD2dt2Scheme.C:58-76 (GREAT while the 2nd and 3rd old level carry the same timeIndex), [EXT]
// backwardDdtScheme::deltaT0_ (GREAT while vf.nOldTimes() < 2).  Here the levels are device fields and this is their
// bookkeeping (the same as fluidstructureinteraction_b200/timelevels.py).
static const int levelId[4] = {GPUPCG_F_DOLD, GPUPCG_F_DOLDOLD, GPUPCG_F_DOOO, GPUPCG_F_DOOOO};

void gpuUnsTotalLagrangianStress::touchOldTimes()
{
    if (!oldTi_.empty() && tiD_ != now_)
    {
        for (int j = (int)oldTi_.size() - 1; j > 0; j--)
        {
            check(gpupcg_mesh_copy_field(gm_, levelId[j], levelId[j - 1]), "gpupcg_mesh_copy_field");
            oldTi_[j] = oldTi_[j - 1];
        }
        check(gpupcg_mesh_copy_field(gm_, levelId[0], GPUPCG_F_D), "gpupcg_mesh_copy_field");
        oldTi_[0] = tiD_;
    }
    tiD_ = now_;
}

void gpuUnsTotalLagrangianStress::ensureOldTimes(int k)
{
    while ((int)oldTi_.size() < k)
    {
        const int j = (int)oldTi_.size();
        check(gpupcg_mesh_copy_field(gm_, levelId[j], j == 0 ? GPUPCG_F_D : levelId[j - 1]), "gpupcg_mesh_copy_field");
        oldTi_.push_back(j == 0 ? tiD_ : oldTi_[j - 1]);
    }
}

void gpuUnsTotalLagrangianStress::newTimeStep()
{
    if (ctl_.eqn.d2dt2Scheme == 2 || ctl_.eqn.ddtScheme == 2) now_++;       // the levels shift when evolve() first touches D
    else check(gpupcg_mesh_new_time_step(gm_), "gpupcg_mesh_new_time_step");
}

bool gpuUnsTotalLagrangianStress::evolve()
{
    gpupcg_evolve_result r;
    const int cap = ctl_.nCorrectors + 1;
    info_.residualHistory.assign(cap, 0.0);
    info_.pcgIterations.assign(3*(size_t)cap, 0);
    gpupcg_eqn_controls& q = ctl_.eqn;
    const bool backward = q.d2dt2Scheme == 2 || q.ddtScheme == 2;
    if (backward)
    {
        // the uses of one assembly in the reference's order (unsTotalLagrangianStress.C:846-865)
        const double GREAT = 1.0e15, dT0 = q.deltaT0;
        q.eD2 = q.eDdtD = q.eDdtD0 = q.eDdtD00 = q.eDamp = q.eU = dT0;
        touchOldTimes();
        if (q.d2dt2Scheme == 2)
        {
            ensureOldTimes(3);
            q.eD2 = (oldTi_[1] == oldTi_[2]) ? GREAT : dT0;
            q.eDdtD = ((int)oldTi_.size() < 2) ? GREAT : dT0;
            q.eDdtD0 = ((int)oldTi_.size() - 1 < 2) ? GREAT : dT0;
            q.eDdtD00 = ((int)oldTi_.size() - 2 < 2) ? GREAT : dT0;
            ensureOldTimes(4);
        }
        else if (q.d2dt2Scheme == 1) ensureOldTimes(2);
        if (q.K > 1e-15)
        {
            if (q.ddtScheme == 2) { q.eDamp = ((int)oldTi_.size() < 2) ? GREAT : dT0; ensureOldTimes(2); }
            else if (q.ddtScheme == 1) ensureOldTimes(1);
        }
    }
    check(gpupcg_evolve_D(solver_, gm_, &ctl_, &r, info_.residualHistory.data(), info_.pcgIterations.data(), cap), "gpupcg_evolve_D");
    if (q.ddtScheme == 2)       // U = fvc::ddt(D) (:973): residual() has used D.oldTime() by now
    {
        ensureOldTimes(1);
        q.eU = ((int)oldTi_.size() < 2) ? 1.0e15 : q.deltaT0;
        ensureOldTimes(2);
    }
    check(gpupcg_sigma_resident(gm_, &ctl_.eqn), "gpupcg_sigma_resident");
    const int nOut = std::min(cap, r.nOuterIterations + 1);
    int nv = 0; for (int d = 0; d < 3; d++) nv += ctl_.validCmpt[d];
    info_.residualHistory.resize(nOut);
    info_.pcgIterations.resize((size_t)nOut*3);        // three slots per outer iteration, the first nValidCmpts in use
    info_.nValidCmpts = nv;
    info_.nOuterIterations = r.nOuterIterations; info_.totalPcgIterations = r.totalPcgIterations;
    info_.enforceLinear = r.enforceLinear; info_.converged = r.converged;
    info_.initialResidual = r.initialResidual; info_.lastSolverInitialResidual = r.lastSolverInitialResidual;
    info_.maxRes = r.maxRes; info_.res = r.res;
    info_.assemble_ms = r.assemble_ms; info_.solve_ms = r.solve_ms; info_.update_ms = r.update_ms;
    // unsTotalLagrangianStress.C:992-998: the summary evolve() prints
    Info << "Solving for D, Initial residual = " << r.initialResidual << ", Final residual = " << r.lastSolverInitialResidual
         << ", No outer iterations = " << r.nOuterIterations << endl;
    return r.converged != 0;
}

int gpuUnsTotalLagrangianStress::fieldId(const word& name) const
{
    if (name == "D") return GPUPCG_F_D;
    if (name == "pointD") return GPUPCG_F_POINTD;
    if (name == "sigma") return GPUPCG_F_SIGMA;
    if (name == "epsilon") return GPUPCG_F_EPSILON;
    if (name == "gradD") return GPUPCG_F_GRADD;
    if (name == "U") return GPUPCG_F_U;
    FatalErrorIn("gpuUnsTotalLagrangianStress::field(...)") << "unknown field " << name << abort(FatalError);
    return -1;
}

label gpuUnsTotalLagrangianStress::fieldSize(const word& name) const
{
    long long n = 0;
    check(gpupcg_mesh_field_size(gm_, fieldId(name), &n), "gpupcg_mesh_field_size");
    return (label)n;
}

void gpuUnsTotalLagrangianStress::field(const word& name, scalar* out) const
{
    check(gpupcg_mesh_get_field(gm_, fieldId(name), out), "gpupcg_mesh_get_field");
}

} // namespace Foam
