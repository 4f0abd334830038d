// plan.cpp -- see plan.h.  Pure host code, O(nCells + nFaces).
#include "plan.h"
#include "../../include/gpupcg.h"

#include <algorithm>
#include <climits>
#include <cstdio>

namespace gpupcg
{

static const int W = 32;    // warp width == slice height

int build_plan(int nCells, int nFaces, const int* l, const int* u,
               int nPatches, const int* patchSizes, const int* const* patchFaceCells,
               const int* patchNbrRank, HostPlan& P, std::string& err)
{
    if (nCells < 0 || nFaces < 0 || nPatches < 0 || (nFaces > 0 && (!l || !u)))
    {
        err = "gpupcg_create: negative size or NULL addressing";
        return GPUPCG_ERR_ARG;
    }
    if (nPatches > 0 && (!patchSizes || !patchFaceCells))
    {
        err = "gpupcg_create: nPatches > 0 but patchSizes/patchFaceCells is NULL";
        return GPUPCG_ERR_ARG;
    }

    P = HostPlan();
    P.nCells = nCells;
    P.nFaces = nFaces;

    // ---- validate upper-triangular order (what lduAddressing of an fvMesh guarantees) and
    //      compute the forward dependency level of every row in the same pass:
    //      level(u) = 1 + max level(l) over faces (l,u).  Valid in one ascending sweep because
    //      every face INTO row l has an owner < l and therefore a smaller face index.
    std::vector<int> level(nCells, 0);
    int prevL = -1;
    for (int f = 0; f < nFaces; f++)
    {
        const int lo = l[f], up = u[f];
        if (lo < 0 || up >= nCells || lo >= up)
        {
            char buf[160];
            snprintf(buf, sizeof(buf),
                     "gpupcg_create: face %d has lowerAddr %d, upperAddr %d (need 0 <= l < u < nCells=%d)",
                     f, lo, up, nCells);
            err = buf;
            return lo >= up && lo >= 0 && up < nCells ? GPUPCG_ERR_ORDERING : GPUPCG_ERR_ARG;
        }
        if (lo < prevL)
        {
            char buf[160];
            snprintf(buf, sizeof(buf),
                     "gpupcg_create: lowerAddr not ascending at face %d (%d after %d): "
                     "addressing is not in upper-triangular order", f, lo, prevL);
            err = buf;
            return GPUPCG_ERR_ORDERING;
        }
        prevL = lo;
        if (level[up] < level[lo] + 1) level[up] = level[lo] + 1;
    }

    int nLevels = 0;
    for (int c = 0; c < nCells; c++) nLevels = std::max(nLevels, level[c] + 1);
    P.nLevels = nLevels;

    // ---- counting sort of rows by level (stable in the original cell index), each level padded
    //      to a whole number of slices so that a warp never straddles two levels.
    std::vector<int> levelCount(nLevels + 1, 0);
    for (int c = 0; c < nCells; c++) levelCount[level[c]]++;
    P.levelSliceStart.assign(nLevels + 1, 0);
    std::vector<long long> levelRowStart(nLevels + 1, 0);
    for (int L = 0; L < nLevels; L++)
    {
        const int nSl = (levelCount[L] + W - 1)/W;
        P.levelSliceStart[L + 1] = P.levelSliceStart[L] + nSl;
        levelRowStart[L + 1] = levelRowStart[L] + (long long)nSl*W;
        P.maxLevelRows = std::max(P.maxLevelRows, levelCount[L]);
    }
    if (levelRowStart[nLevels] > (long long)INT_MAX - W)
    {
        err = "gpupcg_create: padded row count exceeds int32";
        return GPUPCG_ERR_ARG;
    }
    P.nRows = (int)levelRowStart[nLevels];
    P.nSlices = P.nRows/W;
    P.rowOld.assign(P.nRows, -1);
    P.pos.assign(nCells, -1);
    {
        std::vector<long long> cursor(levelRowStart.begin(), levelRowStart.end() - 1);
        for (int c = 0; c < nCells; c++)
        {
            const int r = (int)cursor[level[c]]++;
            P.rowOld[r] = c;
            P.pos[c] = r;
        }
    }
    std::vector<int>().swap(level);

    // ---- per-row entry counts and per-slice widths
    std::vector<unsigned char> nU(P.nRows, 0), nL(P.nRows, 0);
    for (int f = 0; f < nFaces; f++)
    {
        unsigned char& a = nU[P.pos[l[f]]];
        unsigned char& b = nL[P.pos[u[f]]];
        if (a == 255 || b == 255)
        {
            err = "gpupcg_create: a row has more than 255 faces on one side";
            return GPUPCG_ERR_UNSUPPORTED;
        }
        a++; b++;
    }
    P.slices.resize(P.nSlices);
    long long totU = 0, totL = 0;
    for (int s = 0; s < P.nSlices; s++)
    {
        int wU = 0, wL = 0;
        for (int k = 0; k < W; k++)
        {
            wU = std::max<int>(wU, nU[s*W + k]);
            wL = std::max<int>(wL, nL[s*W + k]);
        }
        if (totU + (long long)wU*W > INT_MAX || totL + (long long)wL*W > INT_MAX)
        {
            err = "gpupcg_create: SELL storage exceeds int32 indexing";
            return GPUPCG_ERR_ARG;
        }
        P.slices[s].offU = (int)totU;
        P.slices[s].widthU = wU;
        P.slices[s].offL = (int)totL;
        P.slices[s].widthL = wL;
        totU += (long long)wU*W;
        totL += (long long)wL*W;
    }
    P.entriesU = totU;
    P.entriesL = totL;

    // ---- fill the two lists, faces visited in ascending f so every row keeps the reference order
    P.Ucol.assign(totU, -1);
    P.UfaceOld.assign(totU, -1);
    P.Lidx.assign(totL, -1);
    P.Lcol.assign(totL, -1);
    std::vector<int> faceU(nFaces);
    std::fill(nU.begin(), nU.end(), 0);
    std::fill(nL.begin(), nL.end(), 0);
    for (int f = 0; f < nFaces; f++)
    {
        const int r = P.pos[l[f]];
        const int j = nU[r]++;
        const int idx = P.slices[r/W].offU + j*W + (r % W);
        P.Ucol[idx] = P.pos[u[f]];
        P.UfaceOld[idx] = f;
        faceU[f] = idx;
    }
    for (int f = 0; f < nFaces; f++)
    {
        const int r = P.pos[u[f]];
        const int j = nL[r]++;
        const int idx = P.slices[r/W].offL + j*W + (r % W);
        P.Lidx[idx] = faceU[f];
        P.Lcol[idx] = P.pos[l[f]];
    }

    // ---- coupled interfaces
    P.patchStart.assign(nPatches + 1, 0);
    P.patchNbrRank.assign(nPatches, -1);
    for (int p = 0; p < nPatches; p++)
    {
        if (patchSizes[p] < 0 || (patchSizes[p] > 0 && !patchFaceCells[p]))
        {
            err = "gpupcg_create: bad interface patch size / NULL faceCells";
            return GPUPCG_ERR_ARG;
        }
        P.patchStart[p + 1] = P.patchStart[p] + patchSizes[p];
        P.patchNbrRank[p] = patchNbrRank ? patchNbrRank[p] : -1;
    }
    P.nIfaceFaces = P.patchStart[nPatches];
    P.ifFaceRow.resize(P.nIfaceFaces);
    {
        std::vector<std::pair<int, int> > byRow;   // (row, concat index), stable => (patch, i) order
        byRow.reserve(P.nIfaceFaces);
        for (int p = 0; p < nPatches; p++)
        {
            for (int i = 0; i < patchSizes[p]; i++)
            {
                const int c = patchFaceCells[p][i];
                if (c < 0 || c >= nCells)
                {
                    err = "gpupcg_create: interface faceCells out of range";
                    return GPUPCG_ERR_ARG;
                }
                const int e = P.patchStart[p] + i;
                P.ifFaceRow[e] = P.pos[c];
                byRow.push_back(std::make_pair(P.pos[c], e));
            }
        }
        std::stable_sort(byRow.begin(), byRow.end(),
                         [](const std::pair<int, int>& a, const std::pair<int, int>& b)
                         { return a.first < b.first; });
        P.ifRowStart.push_back(0);
        for (size_t k = 0; k < byRow.size(); k++)
        {
            if (k == 0 || byRow[k].first != byRow[k - 1].first)
            {
                if (k) P.ifRowStart.push_back((int)k);
                P.ifRow.push_back(byRow[k].first);
            }
            P.ifEntry.push_back(byRow[k].second);
        }
        P.ifRowStart.push_back((int)byRow.size());
        if (byRow.empty()) P.ifRowStart.assign(1, 0);
    }
    return GPUPCG_OK;
}

} // namespace gpupcg
