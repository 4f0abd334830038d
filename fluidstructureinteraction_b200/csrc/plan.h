// plan.h -- host-side analysis of an lduAddressing: row tiling + per-tile wavefront schedule for the
// DIC sweeps, and the SELL-32 row storage every kernel of the solver walks.
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
//
// Scheduling.  The triangular sweeps are a dependency chain as long as the mesh is "deep"
// (nx+ny+nz-2 levels on a structured box).  A hand-off between SMs through L2 costs 0.3-4 us on
// B200 (tools/hop_latency.cu, profiles/), a hand-off inside an SM through shared memory ~0.05 us.
// So rows are grouped into TILES (intervals of a topological order => the tile graph is acyclic);
// one CTA sweeps one tile round by round (round = tile-local dependency level) out of shared memory
// and only the edges that cross tiles go through global memory.  Rows are renumbered tile-major,
// round-sorted inside a tile; every vector of the solver lives in that numbering.
#pragma once
#include <vector>
#include <string>

namespace gpupcg
{

struct SliceInfo        // one warp-wide slice of 32 consecutive (new-numbering) rows
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
    int nRows = 0;              // padded (every tile is a whole number of slices)
    int nSlices = 0;
    int nLevels = 0;            // global dependency depth (information only)
    bool bricks = false;        // tiles are bricks of a structured block (tile mode 2 detected one)
    // ---- pencils (tile mode 3 on a structured block): one tile = one pencil, the full length n0 of the block's longest
    // axis times a w1 x w2 (<= 32) cross-section.  Row of cell (q, p1, p2) of a pencil = base + 32*(q + p1 + p2) + (p1 +
    // w1*p2): a warp that marches along the pencil one 32-row step at a time meets every dependency of step t in step
    // t - 1 (forward) / t + 1 (backward) -- own lane (long axis), lane -+1 (axis a1), lane -+w1 (axis a2) -- so a sweep
    // needs no shared-memory round trip inside the pencil: registers + shuffles (kernels_pencil.cuh).
    bool pencil = false;
    int pLong = 0;              // 0 / 1 / 2: the long axis is x / y / z (a1, a2 = the other two in x < y < z order)
    int pStride1 = 0;           // nx: cell-index stride of a y neighbour (x: 1, z: anything else)
    int pW1 = 0, pW2 = 0, pN0 = 0, pSteps = 0;     // cross-section, pencil length, steps = n0 + w1 + w2 - 2
    std::vector<int> pMeta;     // 8 ints per pencil: base row, actual w1, actual w2, tile of the a1-, a2-, a1+, a2+ neighbour (-1), pad
    std::vector<int> pIdxF, pIdxB;  // [nRows*3] U-array index of the coefficient of the row's lower / upper neighbour in
                                    // the order the reference subtracts them (z, y, x), -1: no such neighbour
    int maxLevelRows = 0;
    long long entriesU = 0;
    long long entriesL = 0;

    std::vector<int> rowOld;    // [nRows]  new row -> original cell, -1 for padding rows
    std::vector<int> pos;       // [nCells] original cell -> new row
    std::vector<SliceInfo> slices;

    std::vector<int> Ucol;      // [entriesU] column (new numbering) of owner-side entry, -1 = pad
    std::vector<int> UfaceOld;  // [entriesU] original face id of that entry, -1 = pad
    std::vector<int> Lidx;      // [entriesL] index into the U value array of the same face, -1 = pad
    std::vector<int> Lcol;      // [entriesL] column (new numbering), -1 = pad

    // ---- tiles
    int tileRowsMax = 0;                // largest tile (rows, padded)
    int nTiles = 0;
    int maxRounds = 0;
    int maxTileEntriesL = 0, maxTileEntriesU = 0;       // SELL entries (incl. padding) of the largest tile
    int maxTileExtF = 0, maxTileExtB = 0;
    int maxRowL = 0, maxRowU = 0;                       // most faces of one row on the neighbour / owner side
    std::vector<int> tileStart;         // [nTiles+1] first row of each tile (multiple of 32)
    std::vector<int> tileRoundPtr;      // [nTiles+1] into roundStart (tile t owns K_t+1 entries)
    std::vector<int> roundStart;        // tile-relative first row of each round, last = tile rows
    // work items of the sweep warp: a round cut into passes of <= 32 rows.  item = {first row, count | flags}
    std::vector<int> tileItemPtr;       // [nTiles+1]
    std::vector<int> items;             // 2 ints per item
    int maxItems = 0;
    // per-entry column seen from the tile sweep: >= 0 tile-local row, -1 pad, <= -2 external: -2-k = k-th
    // entry of the tile's external list (values fetched from global memory by the fetch warps)
    std::vector<int> LcolTile;          // [entriesL]  (forward / factor)
    std::vector<int> UcolTile;          // [entriesU]  (backward)
    std::vector<int> extFPtr, extFCol;  // forward: [nTiles+1], global rows in consumption order
    std::vector<int> extFIdx;           // [extForward] U-array index of the coefficient of that cross-tile entry
    // compact 16-bit tile codes (what the kernels stream): 0xFFFF pad, 0x8000|k = k-th cross-tile entry of the
    // tile's list, else tile-local row (col*) / tile-local index into the tile's U block (idxL16)
    std::vector<unsigned short> colL16, idxL16, colU16;
    std::vector<int> extBPtr, extBCol;  // backward

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

// tileRows: target rows per tile (multiple of 32); tileMode: 0 natural-order intervals, 1 greedy compact blobs.
// Returns 0 or a negative gpupcg error code.
int build_plan(int nCells, int nFaces, const int* lowerAddr, const int* upperAddr,
               int nPatches, const int* patchSizes, const int* const* patchFaceCells,
               const int* patchNbrRank, int tileRows, int tileMode, HostPlan& plan, std::string& err);

} // namespace gpupcg
