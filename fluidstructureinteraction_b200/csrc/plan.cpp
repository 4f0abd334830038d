// plan.cpp -- see plan.h.  Pure host code, O(nCells + nFaces).
#include "plan.h"
#include "../../include/gpupcg.h"
#define GPUPCG_ERR_TILE -100    /* internal: retry with smaller tiles */

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <queue>

namespace gpupcg
{

static const int W = 32;    // warp width == slice height

int build_plan(int nCells, int nFaces, const int* l, const int* u,
               int nPatches, const int* patchSizes, const int* const* patchFaceCells,
               const int* patchNbrRank, int tileRows, int tileMode, HostPlan& P, std::string& err)
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
    if (tileRows < W) tileRows = W;
    tileRows = (tileRows/W)*W;

    P = HostPlan();
    P.nCells = nCells;
    P.nFaces = nFaces;

    // ---- validate upper-triangular order (what lduAddressing of an fvMesh guarantees) and
    //      compute the global dependency depth in the same pass (information only):
    //      level(u) = 1 + max level(l) over faces (l,u).  Valid in one ascending sweep because
    //      every face INTO row l has an owner < l and therefore a smaller face index.
    std::vector<int> level(nCells, 0);
    {
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
        std::vector<int> cnt(nLevels + 1, 0);
        for (int c = 0; c < nCells; c++) cnt[level[c]]++;
        for (int L = 0; L < nLevels; L++) P.maxLevelRows = std::max(P.maxLevelRows, cnt[L]);
    }

    // ---- tiles: intervals of a TOPOLOGICAL ORDER of the rows => the tile graph is acyclic (a tile only ever
    //      waits on earlier tiles).  Which topological order decides the tile shapes:
    //        mode 0  natural cell order (l < u): tiles are index ranges -- pencils/slabs on a structured mesh,
    //                ~1 cross-tile hop per mesh plane on the sweep's critical path;
    //        mode 1  greedy breadth-first growth over READY rows (all lower neighbours already placed), seeded
    //                at the lowest-level ready row: compact blobs, several times fewer hops on the critical path
    //                and tiles created in wavefront order (the resident CTAs are the ones the wavefront needs).
    //        mode 2  (auto when the numbering is a lexicographic structured block, i.e. every face joins c and
    //                c+1, c+nx or c+nx*ny as blockMesh numbers a block): bricks bx x by x bz in brick-major order.
    //                A monotone path crosses sum(n_d/b_d) brick faces instead of ~nz + ny/5 index ranges.
    std::vector<int> order(nCells);
    std::vector<int> cuts;                          // first position (in `order`) of every tile, + nCells
    bool bricks = false;
    const bool wantPencils = (tileMode == 3);
    if (tileMode == 3) tileMode = 2;
    if (tileMode == 2 && nFaces > 0)
    {
        // detect: at most three distinct strides u-l, 1 | nx | nx*ny, box completely filled
        long long d[3] = {0, 0, 0};
        int nd = 0;
        bool ok = true;
        for (int f = 0; f < nFaces && ok; f++)
        {
            const long long s = (long long)u[f] - l[f];
            int k = 0;
            while (k < nd && d[k] != s) k++;
            if (k == nd) { if (nd == 3) ok = false; else d[nd++] = s; }
        }
        if (ok)
        {
            std::sort(d, d + nd);
            long long nx = nCells, ny = 1, nz = 1;
            if (nd >= 1 && d[0] != 1) ok = false;
            if (ok && nd >= 2) { nx = d[1]; ny = (nd == 3) ? d[2]/d[1] : nCells/nx; if (nd == 3 && d[2] % d[1]) ok = false; }
            if (ok) { nz = nCells/(nx*ny); if (nx*ny*nz != nCells || nx < 1 || ny < 1) ok = false; }
            // every x-face must stay inside a mesh row etc.: count faces against the box formula
            if (ok)
            {
                const long long expect = (nx - 1)*ny*nz + nx*(ny - 1)*nz + nx*ny*(nz - 1);
                if (expect != nFaces) ok = false;
            }
            for (int f = 0; f < nFaces && ok; f++)
            {
                const long long s = (long long)u[f] - l[f], c = l[f];
                if (s == 1 && nx > 1) { if (c % nx == nx - 1) ok = false; }
                else if (s == nx && ny > 1) { if ((c/nx) % ny == ny - 1) ok = false; }
            }
            if (ok)
            {
                // Brick shape.  One warp sweeps a brick level by level, so a level must not be wider than a warp: the
                // cross-section (the two shorter edges) holds <= 32 rows, 4 x 8 with the 8 along the longer of the two
                // axes, and the brick is long (tileRows/32) along the longest axis: few hops per dependency path AND one
                // pass per round.  (Near-cubic bricks of the same volume have levels of ~40 rows = two passes per round:
                // 0.254 vs 0.205 ms per precondition at 1 M cells, profiles/r1c_dic_latency_study.md.)
                long long n3[3] = {nx, ny, nz};
                int ax[3] = {0, 1, 2};
                std::sort(ax, ax + 3, [&](int a, int b) { return n3[a] != n3[b] ? n3[a] > n3[b] : a < b; });
                if (wantPencils && n3[ax[0]] >= 16 && (long long)nCells >= 4096)
                {
                    // pencils: long axis = the longest (ties: x before y before z, so that the own-lane dependency comes
                    // last in the reference's z, y, x chain); a1 < a2 are the other two axes in x < y < z order
                    const int a0 = ax[0], a1 = (a0 == 0) ? 1 : 0, a2 = (a0 == 2) ? 1 : 2;
                    const int w1 = (int)std::min<long long>(n3[a1], 4), w2 = (int)std::min<long long>(n3[a2], 32/w1);
                    const int nP1 = (int)((n3[a1] + w1 - 1)/w1), nP2 = (int)((n3[a2] + w2 - 1)/w2);
                    const int n0 = (int)n3[a0], steps = ((n0 + w1 + w2 - 2 + 3)/4)*4;      // whole ring stages (PEN_SS = 4)
                    if ((long long)nP1*nP2*steps*32 < (long long)INT_MAX - 64)
                    {
                        P.pencil = true; P.pStride1 = (int)nx; P.pLong = a0; P.pW1 = w1; P.pW2 = w2; P.pN0 = n0; P.pSteps = steps;
                        struct Pen { int s, b, a; };
                        std::vector<Pen> pl;
                        for (int b = 0; b < nP2; b++) for (int a = 0; a < nP1; a++) pl.push_back(Pen{a + b, b, a});
                        std::sort(pl.begin(), pl.end(), [](const Pen& x, const Pen& y)
                                  { return x.s != y.s ? x.s < y.s : (x.b != y.b ? x.b < y.b : x.a < y.a); });
                        std::vector<int> tileOf((size_t)nP1*nP2);
                        for (size_t t = 0; t < pl.size(); t++) tileOf[(size_t)pl[t].b*nP1 + pl[t].a] = (int)t;
                        P.nTiles = (int)pl.size();
                        P.tileStart.assign(P.nTiles + 1, 0);
                        P.pos.assign(nCells, -1);
                        P.pMeta.assign((size_t)8*P.nTiles, -1);
                        const long long str[3] = {1, nx, nx*ny};
                        for (int t = 0; t < P.nTiles; t++)
                        {
                            const int a = pl[t].a, b = pl[t].b;
                            const int base = t*steps*32;
                            const int w1a = (int)std::min<long long>(w1, n3[a1] - (long long)a*w1), w2a = (int)std::min<long long>(w2, n3[a2] - (long long)b*w2);
                            P.tileStart[t] = base;
                            int* M = &P.pMeta[(size_t)8*t];
                            M[0] = base; M[1] = w1a; M[2] = w2a;
                            M[3] = a > 0 ? tileOf[(size_t)b*nP1 + a - 1] : -1;
                            M[4] = b > 0 ? tileOf[(size_t)(b - 1)*nP1 + a] : -1;
                            M[5] = a + 1 < nP1 ? tileOf[(size_t)b*nP1 + a + 1] : -1;
                            M[6] = b + 1 < nP2 ? tileOf[(size_t)(b + 1)*nP1 + a] : -1;
                            M[7] = 0;
                            for (int p2 = 0; p2 < w2a; p2++)
                            for (int p1 = 0; p1 < w1a; p1++)
                            for (int q = 0; q < n0; q++)
                            {
                                const long long c = q*str[a0] + ((long long)a*w1 + p1)*str[a1] + ((long long)b*w2 + p2)*str[a2];
                                P.pos[(size_t)c] = base + 32*(q + p1 + p2) + (p1 + w1*p2);
                            }
                        }
                        P.tileStart[P.nTiles] = P.nTiles*steps*32;
                        P.nRows = P.tileStart[P.nTiles];
                        P.nSlices = P.nRows/W;
                        P.tileRowsMax = steps*32;
                        bricks = true;          // skips the generic tile construction below
                    }
                }
                if (!P.pencil)
                {
                int bb[3];
                bb[ax[2]] = (int)std::min<long long>(n3[ax[2]], 4);
                bb[ax[1]] = (int)std::min<long long>(n3[ax[1]], 32/bb[ax[2]]);
                bb[ax[0]] = (int)std::max<long long>(1, std::min<long long>(n3[ax[0]], tileRows/(bb[ax[1]]*bb[ax[2]])));
                if (const char* ev = getenv("GPUPCG_BRICK"))       // experiment hook: explicit brick edges "bx,by,bz"
                {
                    int e0, e1, e2;
                    if (sscanf(ev, "%d,%d,%d", &e0, &e1, &e2) == 3 && e0 > 0 && e1 > 0 && e2 > 0)
                    {
                        bb[0] = (int)std::min<long long>(e0, nx); bb[1] = (int)std::min<long long>(e1, ny); bb[2] = (int)std::min<long long>(e2, nz);
                    }
                }
                const int bx = bb[0], by = bb[1], bz = bb[2];
                // bricks in WAVEFRONT order (a+b+c ascending, then lexicographic): still topological at tile level,
                // and the CTAs resident at any time are consecutive brick hyperplanes -- the ones the sweep is working
                // on -- instead of a few z-layers whose tiles mostly wait (8 M cells: 5.5 ms -> see profiles/).
                struct Brick { int s, c, b, a; };
                std::vector<Brick> bl;
                for (int c = 0; c*(long long)bz < nz; c++)
                for (int b = 0; b*(long long)by < ny; b++)
                for (int a = 0; a*(long long)bx < nx; a++) bl.push_back(Brick{a + b + c, c, b, a});
                std::sort(bl.begin(), bl.end(), [](const Brick& x, const Brick& y)
                          { return x.s != y.s ? x.s < y.s : (x.c != y.c ? x.c < y.c : (x.b != y.b ? x.b < y.b : x.a < y.a)); });
                int p = 0;
                for (const Brick& B : bl)
                {
                    const long long i0 = (long long)B.a*bx, j0 = (long long)B.b*by, k0 = (long long)B.c*bz;
                    cuts.push_back(p);
                    for (long long k = k0; k < std::min(nz, k0 + bz); k++)
                    for (long long j = j0; j < std::min(ny, j0 + by); j++)
                    for (long long i = i0; i < std::min(nx, i0 + bx); i++)
                        order[p++] = (int)(i + nx*(j + ny*k));
                }
                cuts.push_back(p);
                bricks = true;
                P.bricks = true;
                }
            }
        }
    }
    if (bricks)
    {
    }
    else if (tileMode == 0 || nFaces == 0)

    {
        for (int c = 0; c < nCells; c++) order[c] = c;
    }
    else
    {
        std::vector<int> ownerStart(nCells + 1, 0), indeg(nCells, 0);
        for (int f = 0; f < nFaces; f++) { ownerStart[l[f] + 1]++; indeg[u[f]]++; }
        for (int c = 0; c < nCells; c++) ownerStart[c + 1] += ownerStart[c];
        typedef std::pair<int, int> Key;            // (level, cell)
        std::priority_queue<Key, std::vector<Key>, std::greater<Key> > heap;
        for (int c = 0; c < nCells; c++) if (indeg[c] == 0) heap.push(Key(level[c], c));
        std::vector<int> q;                         // FIFO of the tile being grown
        q.reserve(4*tileRows);
        size_t head = 0;
        int placed = 0, inTile = 0;
        while (placed < nCells)
        {
            if (head == q.size())
            {
                q.clear(); head = 0;
                q.push_back(heap.top().second);     // never empty here: the DAG always has a ready row
                heap.pop();
            }
            const int r = q[head++];
            order[placed++] = r;
            for (int f = ownerStart[r]; f < ownerStart[r + 1]; f++)
            {
                if (--indeg[u[f]] == 0) q.push_back(u[f]);
            }
            if (++inTile == tileRows)
            {
                for (; head < q.size(); head++) heap.push(Key(level[q[head]], q[head]));
                q.clear(); head = 0;
                inTile = 0;
            }
        }
    }
    if (!bricks)
    {
        for (int p = 0; p < nCells; p += tileRows) cuts.push_back(p);
        cuts.push_back(nCells);
    }
    const int nTiles = P.pencil ? P.nTiles : (int)cuts.size() - 1;
    P.nTiles = nTiles;
    if (P.pencil)
    {
        // one dummy round per tile: the tile kernels are not used on pencils
        std::vector<int>().swap(level);
        std::vector<int>().swap(order);
        P.tileRoundPtr.assign(nTiles + 1, 0);
        P.tileItemPtr.assign(nTiles + 1, 0);
        P.maxRounds = P.pSteps;
    }
    else
    {


    // Round of a row = rank of its GLOBAL dependency level among the levels present in its tile.  Using the
    // global level (not the tile-local depth) keeps all tiles marching in lock step with the global wavefront:
    // a tile at level L only ever needs level L-1 of its neighbours, so a consumer tile trails its producer by
    // one hand-off, however the tile boundaries cut the mesh rows (tile-local depths made tiles wait for values
    // their predecessor produces o(tile) rounds later -- measured 10x slower, profiles/r1_dic_sweep_design.md).
    std::vector<int> lvl(nCells, 0);
    {
        std::vector<int> present;
        for (int t = 0; t < nTiles; t++)
        {
            const int i0 = cuts[t], i1 = cuts[t + 1];
            present.clear();
            for (int i = i0; i < i1; i++) present.push_back(level[order[i]]);
            std::sort(present.begin(), present.end());
            present.erase(std::unique(present.begin(), present.end()), present.end());
            for (int i = i0; i < i1; i++)
                lvl[order[i]] = (int)(std::lower_bound(present.begin(), present.end(), level[order[i]]) - present.begin());
        }
    }
    std::vector<int>().swap(level);

    // ---- new numbering: tile-major, rows of a tile sorted by round (stable), tile padded to x32
    P.tileStart.assign(nTiles + 1, 0);
    P.tileRoundPtr.assign(nTiles + 1, 0);
    P.pos.assign(nCells, -1);
    {
        long long row = 0;
        std::vector<int> cnt;
        for (int t = 0; t < nTiles; t++)
        {
            const int i0 = cuts[t], i1 = cuts[t + 1];
            int K = 0;
            for (int i = i0; i < i1; i++) K = std::max(K, lvl[order[i]] + 1);
            cnt.assign(K + 1, 0);
            for (int i = i0; i < i1; i++) cnt[lvl[order[i]] + 1]++;
            for (int k = 0; k < K; k++) cnt[k + 1] += cnt[k];
            const int padded = ((i1 - i0 + W - 1)/W)*W;
            P.tileStart[t] = (int)row;
            for (int k = 0; k < K; k++) P.roundStart.push_back(cnt[k]);
            P.roundStart.push_back(padded);         // padding rows ride in the last round
            P.tileRoundPtr[t + 1] = (int)P.roundStart.size();
            P.maxRounds = std::max(P.maxRounds, K);
            P.tileRowsMax = std::max(P.tileRowsMax, padded);
            for (int i = i0; i < i1; i++) P.pos[order[i]] = (int)row + cnt[lvl[order[i]]]++;
            row += padded;
            if (row > (long long)INT_MAX - W)
            {
                err = "gpupcg_create: padded row count exceeds int32";
                return GPUPCG_ERR_ARG;
            }
        }
        P.tileStart[nTiles] = (int)row;
        P.nRows = (int)row;
        P.nSlices = P.nRows/W;
    }
    std::vector<int>().swap(order);
    // work items: every round cut into passes of <= 32 rows; bit 16 = first pass of its round, bit 17 = last
    P.tileItemPtr.assign(nTiles + 1, 0);
    for (int t = 0; t < nTiles; t++)
    {
        const int* rs = &P.roundStart[P.tileRoundPtr[t]];
        const int K = P.tileRoundPtr[t + 1] - P.tileRoundPtr[t] - 1;
        for (int k = 0; k < K; k++)
        {
            for (int a = rs[k]; a < rs[k + 1]; a += W)
            {
                const int cnt = std::min(W, rs[k + 1] - a);
                int flags = 0;
                if (a == rs[k]) flags |= 1 << 16;
                if (a + W >= rs[k + 1]) flags |= 1 << 17;
                P.items.push_back(a);
                P.items.push_back(cnt | flags);
            }
        }
        P.tileItemPtr[t + 1] = (int)P.items.size()/2;
        P.maxItems = std::max(P.maxItems, P.tileItemPtr[t + 1] - P.tileItemPtr[t]);
    }
    std::vector<int>().swap(lvl);
    }   // !pencil
    P.rowOld.assign(P.nRows, -1);
    for (int c = 0; c < nCells; c++) P.rowOld[P.pos[c]] = c;

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
    for (int r = 0; r < P.nRows; r++) { P.maxRowU = std::max<int>(P.maxRowU, nU[r]); P.maxRowL = std::max<int>(P.maxRowL, nL[r]); }
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
    {
        std::vector<int> faceU(nFaces);
        std::vector<unsigned char> cu(P.nRows, 0), cl(P.nRows, 0);
        for (int f = 0; f < nFaces; f++)
        {
            const int r = P.pos[l[f]];
            const int j = cu[r]++;
            const int idx = P.slices[r/W].offU + j*W + (r % W);
            P.Ucol[idx] = P.pos[u[f]];
            P.UfaceOld[idx] = f;
            faceU[f] = idx;
        }
        for (int f = 0; f < nFaces; f++)
        {
            const int r = P.pos[u[f]];
            const int j = cl[r]++;
            const int idx = P.slices[r/W].offL + j*W + (r % W);
            P.Lidx[idx] = faceU[f];
            P.Lcol[idx] = P.pos[l[f]];
        }
    }

    if (P.pencil)
    {
        // coefficient index tables of the pencil sweeps: per row the U index of the face to the lower (forward / factor)
        // and upper (backward) neighbour along z, y, x -- the order the reference's face loops subtract them in
        P.pIdxF.assign((size_t)3*P.nRows, -1);
        P.pIdxB.assign((size_t)3*P.nRows, -1);
        for (int r = 0; r < P.nRows; r++)
        {
            const int c = P.rowOld[r];
            if (c < 0) continue;
            const SliceInfo& si = P.slices[r/W];
            for (int j = 0; j < nL[r]; j++)
            {
                const int e = si.offL + j*W + (r % W);
                const long long d = (long long)c - P.rowOld[P.Lcol[e]];
                const int slot = (d == 1) ? 2 : ((P.pStride1 == d) ? 1 : 0);
                P.pIdxF[(size_t)3*r + slot] = P.Lidx[e];
            }
            for (int j = 0; j < nU[r]; j++)
            {
                const int e = si.offU + j*W + (r % W);
                const long long d = (long long)P.rowOld[P.Ucol[e]] - c;
                const int slot = (d == 1) ? 2 : ((P.pStride1 == d) ? 1 : 0);
                P.pIdxB[(size_t)3*r + slot] = e;
            }
        }
        P.extFPtr.assign(nTiles + 1, 0);
        P.extBPtr.assign(nTiles + 1, 0);
    }
    else
    {
    // ---- tile view of the columns + external (cross-tile) lists in consumption order
    P.LcolTile.assign(totL, -1);
    P.UcolTile.assign(totU, -1);
    P.colL16.assign(totL, 0xFFFF);
    P.idxL16.assign(totL, 0xFFFF);
    P.colU16.assign(totU, 0xFFFF);
    P.extFPtr.assign(nTiles + 1, 0);
    P.extBPtr.assign(nTiles + 1, 0);
    for (int t = 0; t < nTiles; t++)
    {
        const int r0 = P.tileStart[t], r1 = P.tileStart[t + 1];
        const int s0 = r0/W, s1 = r1/W;
        const int eL = (s1 < P.nSlices ? P.slices[s1].offL : (int)totL) - P.slices[s0].offL;
        const int eU = (s1 < P.nSlices ? P.slices[s1].offU : (int)totU) - P.slices[s0].offU;
        P.maxTileEntriesL = std::max(P.maxTileEntriesL, eL);
        P.maxTileEntriesU = std::max(P.maxTileEntriesU, eU);
        // forward: rounds ascending = rows ascending, entries ascending
        const int baseF = (int)P.extFCol.size();
        for (int r = r0; r < r1; r++)
        {
            const SliceInfo& si = P.slices[r/W];
            for (int j = 0; j < nL[r]; j++)
            {
                const int e = si.offL + j*W + (r % W);
                const int col = P.Lcol[e];
                if (col >= r0 && col < r1)
                {
                    P.LcolTile[e] = col - r0;
                    P.colL16[e] = (unsigned short)(col - r0);
                    P.idxL16[e] = (unsigned short)(P.Lidx[e] - P.slices[s0].offU);
                }
                else
                {
                    const int k = (int)P.extFCol.size() - baseF;
                    P.LcolTile[e] = -2 - k;
                    P.colL16[e] = (unsigned short)(0x8000 | k);
                    P.extFCol.push_back(col);
                    P.extFIdx.push_back(P.Lidx[e]);
                }
            }
        }
        P.extFPtr[t + 1] = (int)P.extFCol.size();
        P.maxTileExtF = std::max(P.maxTileExtF, P.extFPtr[t + 1] - baseF);
        // backward: rounds descending, rows of a round ascending, entries descending
        const int baseB = (int)P.extBCol.size();
        const int* rs = &P.roundStart[P.tileRoundPtr[t]];
        const int K = P.tileRoundPtr[t + 1] - P.tileRoundPtr[t] - 1;
        for (int k = K - 1; k >= 0; k--)
        {
            for (int r = r0 + rs[k]; r < r0 + rs[k + 1]; r++)
            {
                const SliceInfo& si = P.slices[r/W];
                for (int j = nU[r] - 1; j >= 0; j--)
                {
                    const int e = si.offU + j*W + (r % W);
                    const int col = P.Ucol[e];
                    if (col >= r0 && col < r1)
                    {
                        P.UcolTile[e] = col - r0;
                        P.colU16[e] = (unsigned short)(col - r0);
                    }
                    else
                    {
                        const int k = (int)P.extBCol.size() - baseB;
                        P.UcolTile[e] = -2 - k;
                        P.colU16[e] = (unsigned short)(0x8000 | k);
                        P.extBCol.push_back(col);
                    }
                }
            }
        }
        P.extBPtr[t + 1] = (int)P.extBCol.size();
        P.maxTileExtB = std::max(P.maxTileExtB, P.extBPtr[t + 1] - baseB);
    }

    if (P.tileRowsMax > 0x8000 || P.maxTileEntriesU > 0x8000 || P.maxTileExtF > 0x7FFE || P.maxTileExtB > 0x7FFE)
    {
        err = "gpupcg_create: tile too large for the 16-bit tile codes";
        return GPUPCG_ERR_TILE;
    }
    }   // !pencil

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
