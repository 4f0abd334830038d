// foamlike.C -- implementation of the foam-extend stand-in (see lduMatrix.H).  No solver arithmetic.
#include "lduMatrix.H"

#include <algorithm>
#include <cctype>
#include <cstdio>

namespace Foam
{

FatalErrorType FatalError;

FatalErrorType& FatalErrorInFunction_(const char* where)
{
    FatalError.os.str("");
    FatalError.where = where;
    return FatalError;
}

void operator<<(FatalErrorType& e, const errorManip&)
{
    const std::string msg = "--> FOAM FATAL ERROR:\n" + e.os.str() + "\n\n    From function " + e.where;
    e.os.str("");
    throw error(msg);
}

// ---- dictionary -----------------------------------------------------------------------------
void dictionary::parse(const std::string& t, size_t& p)
{
    auto skip = [&]()
    {
        while (p < t.size())
        {
            if (std::isspace((unsigned char)t[p])) p++;
            else if (t.compare(p, 2, "//") == 0) { while (p < t.size() && t[p] != '\n') p++; }
            else if (t.compare(p, 2, "/*") == 0) { p = t.find("*/", p); p = (p == std::string::npos) ? t.size() : p + 2; }
            else break;
        }
    };
    while (true)
    {
        skip();
        if (p >= t.size()) return;
        if (t[p] == '}') { p++; return; }
        size_t k0 = p;
        while (p < t.size() && !std::isspace((unsigned char)t[p]) && t[p] != '{' && t[p] != ';') p++;
        const word key = t.substr(k0, p - k0);
        skip();
        if (p < t.size() && t[p] == '{')
        {
            p++;
            dictionary sub;
            sub.name_ = key;
            sub.parse(t, p);
            subs_[key] = sub;
        }
        else
        {
            size_t v0 = p;
            while (p < t.size() && t[p] != ';') p++;
            std::string val = t.substr(v0, p - v0);
            while (!val.empty() && std::isspace((unsigned char)val.back())) val.pop_back();
            entries_[key] = val;
            if (p < t.size()) p++;
        }
    }
}

const dictionary& dictionary::subDict(const word& k) const
{
    std::map<word, dictionary>::const_iterator it = subs_.find(k);
    if (it == subs_.end())
    {
        FatalErrorIn("dictionary::subDict(const word&) const")
            << "keyword " << k << " is undefined in dictionary " << name_ << abort(FatalError);
    }
    return it->second;
}

std::string dictionary::lookup(const word& k) const
{
    std::map<word, std::string>::const_iterator it = entries_.find(k);
    if (it == entries_.end())
    {
        FatalErrorIn("dictionary::lookup(const word&) const")
            << "keyword " << k << " is undefined in dictionary " << name_ << abort(FatalError);
    }
    return it->second;
}

std::vector<word> dictionary::toc() const
{
    std::vector<word> r;
    for (auto& e : entries_) r.push_back(e.first);
    for (auto& e : subs_) r.push_back(e.first);
    return r;
}

// ---- lduAddressing derived data ([EXT] lduAddressing::calcOwnerStart/calcLosort/calcLosortStart) --------------
const unallocLabelList& lduAddressing::ownerStartAddr() const
{
    if (ownerStart_.size() == 0)
    {
        ownerStart_.setSize(size_ + 1);
        for (label c = 0; c <= size_; c++) ownerStart_[c] = 0;
        forAll(lower_, f) ownerStart_[lower_[f] + 1]++;
        for (label c = 0; c < size_; c++) ownerStart_[c + 1] += ownerStart_[c];
    }
    return ownerStart_;
}

const unallocLabelList& lduAddressing::losortAddr() const
{
    if (losort_.size() == 0 && upper_.size() > 0)
    {
        const unallocLabelList& ls = losortStartAddr();
        losort_.setSize(upper_.size());
        std::vector<label> cursor(size_);
        for (label c = 0; c < size_; c++) cursor[c] = ls[c];
        forAll(upper_, f) losort_[cursor[upper_[f]]++] = f;
    }
    return losort_;
}

const unallocLabelList& lduAddressing::losortStartAddr() const
{
    if (losortStart_.size() == 0)
    {
        losortStart_.setSize(size_ + 1);
        for (label c = 0; c <= size_; c++) losortStart_[c] = 0;
        forAll(upper_, f) losortStart_[upper_[f] + 1]++;
        for (label c = 0; c < size_; c++) losortStart_[c + 1] += losortStart_[c];
    }
    return losortStart_;
}

// ---- lduMatrix --------------------------------------------------------------------------------
int lduMatrix::debug = 1;
const scalar lduMatrix::great_ = 1.0e+20;
const scalar lduMatrix::small_ = 1.0e-20;

const scalarField& lduMatrix::diag() const
{
    if (!diagPtr_) FatalErrorIn("lduMatrix::diag() const") << "diagPtr_ unallocated" << abort(FatalError);
    return *diagPtr_;
}

const scalarField& lduMatrix::upper() const
{
    if (!lowerPtr_ && !upperPtr_) FatalErrorIn("lduMatrix::upper() const") << "lowerPtr_ or upperPtr_ unallocated" << abort(FatalError);
    return upperPtr_ ? *upperPtr_ : *lowerPtr_;
}

const scalarField& lduMatrix::lower() const
{
    if (!lowerPtr_ && !upperPtr_) FatalErrorIn("lduMatrix::lower() const") << "lowerPtr_ or upperPtr_ unallocated" << abort(FatalError);
    return lowerPtr_ ? *lowerPtr_ : *upperPtr_;
}

void lduMatrix::solverPerformance::print() const
{
    if (debug)
    {
        Info << solverName_ << ":  Solving for " << fieldName_;
        if (singular_) Info << ":  solution singularity" << endl;
        else Info << ", Initial residual = " << initialResidual_ << ", Final residual = " << finalResidual_
                  << ", No Iterations " << noIterations_ << endl;
    }
}

lduMatrix::solver::symMatrixConstructorTable* lduMatrix::solver::symMatrixConstructorTablePtr_ = 0;

lduMatrix::solver::solver(const word& fieldName, const lduMatrix& matrix, const FieldField<Field, scalar>& bou,
                          const FieldField<Field, scalar>& intc, const lduInterfaceFieldPtrsList& interfaces,
                          const dictionary& dict)
: fieldName_(fieldName), matrix_(matrix), coupleBouCoeffs_(bou), coupleIntCoeffs_(intc), interfaces_(interfaces),
  dict_(dict), tolerance_(1e-6), relTolerance_(0), minIter_(0), maxIter_(1000)
{
    readControls();
}

void lduMatrix::solver::readControls()
{
    dict_.readIfPresent("minIter", minIter_);
    dict_.readIfPresent("maxIter", maxIter_);
    dict_.readIfPresent("tolerance", tolerance_);
    dict_.readIfPresent("relTol", relTolerance_);
}

lduMatrix::solver* lduMatrix::solver::New(const word& fieldName, const lduMatrix& matrix,
                                          const FieldField<Field, scalar>& bou, const FieldField<Field, scalar>& intc,
                                          const lduInterfaceFieldPtrsList& interfaces, const dictionary& dict)
{
    const word solverName(dict.lookup("solver"));
    if (matrix.diagonal())
    {
        FatalErrorIn("lduMatrix::solver::New") << "diagonal matrix: foam-extend short-circuits to diagonalSolver, "
            << "which this stand-in does not provide" << abort(FatalError);
    }
    else if (matrix.symmetric())
    {
        symMatrixConstructorTable::iterator it;
        if (!symMatrixConstructorTablePtr_ || (it = symMatrixConstructorTablePtr_->find(solverName)) == symMatrixConstructorTablePtr_->end())
        {
            FatalErrorType& e = FatalIOErrorIn("lduMatrix::solver::New", dict);
            e << "Unknown symmetric matrix solver " << solverName << endl << endl
              << "Valid symmetric matrix solvers are :" << endl;
            if (symMatrixConstructorTablePtr_) for (auto& t : *symMatrixConstructorTablePtr_) e << "    " << t.first << endl;
            e << abort(FatalError);
        }
        return it->second(fieldName, matrix, bou, intc, interfaces, dict);
    }
    else if (matrix.asymmetric())
    {
        FatalErrorIn("lduMatrix::solver::New") << "Unknown asymmetric matrix solver " << solverName
            << " (this stand-in only carries the symMatrix table)" << abort(FatalError);
    }
    FatalErrorIn("lduMatrix::solver::New") << "cannot solve incomplete matrix, no diagonal or off-diagonal coefficient"
        << abort(FatalError);
    return 0;
}

} // namespace Foam
