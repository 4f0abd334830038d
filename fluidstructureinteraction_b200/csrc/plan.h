// plan.h -- host-side analysis of an lduAddressing: the wavefront ("level") ordering of the DIC
// dependency graph and the SELL-32 row storage every kernel of the solver walks.
//
// Reference semantics being scheduled (foam-extend-3.1 [EXT], restated in SURVEY.md Appendix A):
//   Amul           for f ascending: Ax[u[f]] += lower[f]*x[l[f]];  Ax[l[f]] += upper[f]*x[u[f]]
//   DIC factor     for f ascending: rD[u[f]] -= upper[f]*upper[f]/rD[l[f]]
//   DIC forward    for f ascending: wA[u[f]] -= rD[u[f]]*upper[f]*wA[l[f]]
//   DIC backward   for f descending: wA[l[f]] -= rD[l[f]]*upper[f]*wA[u[f]]
// With faces in upper-triangular order (l < u, sorted by l) the per-row operand order of all four
// loops is: neighbour-side faces (u == row) ascending f, then owner-side faces (l == row)
// ascending f.  The plan stores exactly those two lists per row, so a one-thread-per-row gather
// reproduces the face loops bit for bit, whatever order the rows are scheduled in.
#pragma once
#include <vector>
#include <string>

namespace gpupcg
{

struct SliceInfo        // one warp-wide slice of 32 rows (all rows of a slice share a level)
{
    int offU;           // first entry of the owner-side block  (entry = offU + j*32 + lane)
    int widthU;         // owner-side entries per row in this slice (max over its 32 rows)
    int offL;           // first entry of the neighbour-side block
    int widthL;
};

struct HostPlan
{
    int nCells = 0;
    int nFaces = 0;
    int nRows = 0;              // padded
    int nSlices = 0;
    int nLevels = 0;
    int maxLevelRows = 0;
    long long entriesU = 0;
    long long entriesL = 0;

    std::vector<int> rowOld;    // [nRows]  new row -> original cell, -1 for padding rows
    std::vector<int> pos;       // [nCells] original cell -> new row
    std::vector<SliceInfo> slices;
    std::vector<int> levelSliceStart;   // [nLevels+1]

    std::vector<int> Ucol;      // [entriesU] column (new numbering) of owner-side entry, -1 = pad
    std::vector<int> UfaceOld;  // [entriesU] original face id of that entry, -1 = pad
    std::vector<int> Lidx;      // [entriesL] index into the U value array of the same face, -1 = pad
    std::vector<int> Lcol;      // [entriesL] column (new numbering), -1 = pad

    // coupled interfaces, grouped by row: for every row that owns interface faces, its entries in
    // (patch, face-in-patch) order -- the order updateMatrixInterfaces applies them in.
    int nIfaceFaces = 0;
    std::vector<int> ifRow;         // [nIfRows] new row id
    std::vector<int> ifRowStart;    // [nIfRows+1] into ifEntry
    std::vector<int> ifEntry;       // [nIfaceFaces] index into the concatenated patch arrays
    std::vector<int> patchStart;    // [nPatches+1]
    std::vector<int> patchNbrRank;  // [nPatches]
    std::vector<int> ifFaceRow;     // [nIfaceFaces] new row id of faceCells (send gather), concat order
};

// Returns 0 or a negative gpupcg error code; err receives the text.
int build_plan(int nCells, int nFaces, const int* lowerAddr, const int* upperAddr,
               int nPatches, const int* patchSizes, const int* const* patchFaceCells,
               const int* patchNbrRank, HostPlan& plan, std::string& err);

} // namespace gpupcg
