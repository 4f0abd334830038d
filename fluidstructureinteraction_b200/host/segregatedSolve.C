#include "segregatedSolve.H"
#include "gpuPCG.H"

Foam::lduMatrix::solverPerformance Foam::fvMatrixLike::solve
(
    std::vector<scalar>& psi,
    const dictionary& solverDict,
    std::vector<lduMatrix::solverPerformance>* perCmpt
)
{
    static const char* cmptNames[3] = {"x", "y", "z"};
    const label nCells = matrix.lduAddr().size();
    lduMatrix::solverPerformance solverPerfVec("fvMatrix::solve", psiName);

    scalarField saveDiag(matrix.diag());

    // addBoundarySource: non-coupled patches add their boundaryCoeffs to the owner cell
    std::vector<scalar> totalSource(source);
    for (size_t p = 0; p < patches.size(); p++)
    {
        const fvPatchCoeffs& pc = patches[p];
        forAll(pc.faceCells, i)
        {
            for (int c = 0; c < nCmpt; c++) totalSource[pc.faceCells[i]*nCmpt + c] += pc.boundaryCoeffs[i*nCmpt + c];
        }
    }

    // no coupled interfaces in this stand-in
    FieldField<Field, scalar> bouCoeffsCmpt(0), intCoeffsCmpt(0);
    lduInterfaceFieldPtrsList interfaces(0);

    // gpuPCG: all valid components in one call (batchComponents, default yes) -- the systems share upper() and are
    // independent, the device kernels advance them together (gpuPCG::solveVector -> gpupcg_solve_vector)
    bool allValid = nCmpt > 1;
    for (int cmpt = 0; cmpt < nCmpt; cmpt++) allValid = allValid && validComponent[cmpt];
    if (allValid && word(solverDict.lookup("solver")) == gpuPCG::typeName
        && solverDict.lookupOrDefault<word>("batchComponents", "yes") != "no")
    {
        lduMatrix::solver* s = lduMatrix::solver::New(psiName, matrix, bouCoeffsCmpt, intCoeffsCmpt, interfaces, solverDict);
        std::vector<scalar> diagCmpt(size_t(nCells)*nCmpt);
        for (label c = 0; c < nCells; c++)
            for (int cmpt = 0; cmpt < nCmpt; cmpt++) diagCmpt[c*nCmpt + cmpt] = saveDiag[c];
        for (size_t p = 0; p < patches.size(); p++)
        {
            const fvPatchCoeffs& pc = patches[p];
            forAll(pc.faceCells, i)
                for (int cmpt = 0; cmpt < nCmpt; cmpt++) diagCmpt[pc.faceCells[i]*nCmpt + cmpt] += pc.internalCoeffs[i*nCmpt + cmpt];
        }
        std::vector<lduMatrix::solverPerformance> perfs;
        try
        {
            dynamic_cast<const gpuPCG&>(*s).solveVector(nCmpt, &psi[0], &totalSource[0], &diagCmpt[0],
                                                        std::vector<const scalar*>(), perfs);
        }
        catch (...)
        {
            delete s;
            throw;
        }
        delete s;
        for (size_t k = 0; k < perfs.size(); k++)
        {
            perfs[k].print();
            if (perCmpt) perCmpt->push_back(perfs[k]);
            if (k == 0 || (perfs[k].initialResidual() > solverPerfVec.initialResidual() && !perfs[k].singular()))
                solverPerfVec = perfs[k];
        }
        return solverPerfVec;
    }

    scalarField psiCmpt(nCells), sourceCmpt(nCells);
    bool first = true;
    for (int cmpt = 0; cmpt < nCmpt; cmpt++)
    {
        if (!validComponent[cmpt]) continue;

        // addBoundaryDiag(diag(), cmpt)
        scalarField& diag = matrix.diag();
        for (size_t p = 0; p < patches.size(); p++)
        {
            const fvPatchCoeffs& pc = patches[p];
            forAll(pc.faceCells, i) diag[pc.faceCells[i]] += pc.internalCoeffs[i*nCmpt + cmpt];
        }
        for (label c = 0; c < nCells; c++)
        {
            psiCmpt[c] = psi[c*nCmpt + cmpt];
            sourceCmpt[c] = totalSource[c*nCmpt + cmpt];
        }

        const word fieldName = (nCmpt == 1) ? psiName : psiName + cmptNames[cmpt];
        lduMatrix::solver* s = lduMatrix::solver::New(fieldName, matrix, bouCoeffsCmpt, intCoeffsCmpt, interfaces, solverDict);
        lduMatrix::solverPerformance perf;
        try
        {
            perf = s->solve(psiCmpt, sourceCmpt, (direction)cmpt);
        }
        catch (...)
        {
            delete s;
            diag = saveDiag;
            throw;
        }
        delete s;
        perf.print();
        if (perCmpt) perCmpt->push_back(perf);
        if (first || (perf.initialResidual() > solverPerfVec.initialResidual() && !perf.singular()))
        {
            solverPerfVec = perf;
            first = false;
        }
        for (label c = 0; c < nCells; c++) psi[c*nCmpt + cmpt] = psiCmpt[c];

        for (label c = 0; c < nCells; c++) diag[c] = saveDiag[c];
    }
    return solverPerfVec;
}
