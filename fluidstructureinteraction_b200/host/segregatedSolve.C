#include "segregatedSolve.H"

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
