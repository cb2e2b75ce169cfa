// slab_host.cpp - brute-force check of the index arithmetic of the slab-ordered M kernel (thunder_b200/csrc/thb_slab.cuh),
// compiled for the host by tests/test_slab_host.py.  For one rotation and one slab thickness: every pixel of the list must be a
// candidate of the slab that holds its exact cell base z0 (exact = slice_coord -> fold -> floor of thb_math.cuh, what the
// scatter itself computes), and no pixel may be a candidate twice within one slab.
#include <vector>
#include <algorithm>
#include "thb_slab.cuh"

using namespace thb;

extern "C" int slab_check(int pf, int nPxl, const int* iCol, const int* iRow, const double* quat, int vdim, int th,
                          long long* nCandidates, long long* nHits)
{
    // row-major runs, as thb_set_insert_pixels builds them
    std::vector<int> perm(nPxl);
    for (int i = 0; i < nPxl; ++i) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int l, int r) { return iRow[l] != iRow[r] ? iRow[l] < iRow[r] : iCol[l] < iCol[r]; });
    std::vector<Seg> segs;
    for (int k = 0; k < nPxl; ++k) {
        const int x = iCol[perm[k]], y = iRow[perm[k]];
        if (!segs.empty() && segs.back().j == y && segs.back().iFirst + segs.back().count == x) segs.back().count++;
        else segs.push_back(Seg{y, x, 1, k});
    }
    const Rot2 rot = quat_to_rot2(quat);
    std::vector<int> z0of(nPxl);
    for (int k = 0; k < nPxl; ++k) {
        float x, y, z, xd, yd, zd;
        int x0, y0, z0;
        slice_coord(rot, (double)(pf * iCol[perm[k]]), (double)(pf * iRow[perm[k]]), x, y, z);
        fold_floor(x, y, z, x0, y0, z0, xd, yd, zd);
        z0of[k] = z0;
    }
    const int zMin = -(vdim / 2), nSlab = (vdim + th - 1) / th;
    std::vector<int> seen(nPxl);
    long long cand = 0, hits = 0;
    for (int s = 0; s < nSlab; ++s) {
        const int zlo = zMin + s * th, zhi = zlo + th;
        std::fill(seen.begin(), seen.end(), 0);
        for (const Seg& sg : segs) {
            SlabRec rec;
            const int n = seg_intervals(rot, pf, sg, zlo, zhi, rec);
            for (int u = 0; u < n; ++u) {
                const int p = slab_element(rec, u);
                if (p < sg.start || p >= sg.start + sg.count) return -1;      // outside its own segment
                if (seen[p]++) return -2;                                      // candidate twice in one slab
                ++cand;
            }
        }
        for (int k = 0; k < nPxl; ++k) {
            const bool in = z0of[k] >= zlo && z0of[k] < zhi;
            if (in && !seen[k]) return -3;                                     // a sample of this slab was missed
            hits += in;
        }
    }
    for (int k = 0; k < nPxl; ++k)
        if (z0of[k] < zMin || z0of[k] >= zMin + nSlab * th) return -4;        // outside every slab
    *nCandidates = cand;
    *nHits = hits;
    return 0;
}
