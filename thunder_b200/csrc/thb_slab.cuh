// thb_slab.cuh - index arithmetic of the slab-ordered M kernel (thb_insert2.cuh), host+device inline so that tests/ can
// compile it with g++ and check it by brute force against the exact coordinate arithmetic of the scatter.
//
// A "segment" is a run of the M pixel list: one image row j, consecutive columns i = iFirst .. iFirst + count - 1, stored
// at positions start .. of the (row-major) list.  Along a segment the slice coordinate of Reconstructor::insertP
// (src/Reconstructor.cpp:809-815: R (pf i, pf j, 0)) is linear in i, so the pixels whose cell base z0 = floor(+-z) lies in the
// slab [zlo, zhi) form at most two index intervals - one per side of the Hermitian fold x = 0 (src/Image/Volume.cpp:340-357) -
// plus the few pixels next to the fold, which are always handed to the exact test.  The intervals are CONSERVATIVE: margins of
// 1e-2 voxel (the fp32 rounding of the exact coordinate is < 1e-4 at |k| < 2048) and one index on each side.
#pragma once
#include <math.h>
#include "thb_math.cuh"

namespace thb {

struct Seg { int j, iFirst, count, start; };
// up to three runs of list positions: [p1, p1 + c1) | [pm, pm + cm) | [p2, ...); c1 | cm << 16 packed
struct SlabRec { int p1, c1cm, pm, p2; };

THB_HD int slab_clampi(double v, int lo, int hi)
{
    if (!(v > (double)lo)) return lo;      // also NaN
    if (v > (double)hi) return hi;
    return (int)v;
}

// indices i in [pa, pb] with zlo <= sg (alpha i + beta) < zhi, widened by the margin mz in z and one index each side
THB_HD void slab_z_interval(double alpha, double beta, double sg, double zlo, double zhi, int& pa, int& pb)
{
    if (pb < pa) return;
    const double g = sg * alpha, d = sg * beta, mz = 1e-2;
    if (fabs(g) < 1e-9) {
        // z is constant along the run up to ~1e-6
        if (!(d >= zlo - mz && d < zhi + mz)) pb = pa - 1;
        return;
    }
    const double t0 = (zlo - mz - d) / g, t1 = (zhi + mz - d) / g;
    const double lo = fmin(t0, t1), hi = fmax(t0, t1);
    const int a = slab_clampi(floor(lo) - 1.0, pa, pb + 1), b = slab_clampi(ceil(hi) + 1.0, pa - 1, pb);
    pa = a;
    pb = b;
}

// returns the number of candidate pixels of (rotation r, segment sg) for the slab z0 in [zlo, zhi)
THB_HD int seg_intervals(const Rot2& r, int pf, const Seg sg, int zlo, int zhi, SlabRec& rec)
{
    const int i0 = sg.iFirst, i1 = sg.iFirst + sg.count - 1;
    const double b = (double)(pf * sg.j);
    const double kap = (double)pf * r.c0[0], lam = r.c1[0] * b;        // x(i) = kap i + lam
    const double alp = (double)pf * r.c0[2], bet = r.c1[2] * b;        // z(i) = alp i + bet
    // [i0, A) sign sL | [A, B) around the fold, unconditional | [B, i1] sign sR
    int A, B;
    double sL, sR;
    const double mx = 1e-2;
    if (fabs(kap) < 1e-9) {
        if (fabs(lam) <= mx) { A = i0; B = i1 + 1; sL = sR = 1.0; }           // x ~ 0 along the whole run: exact test only
        else { A = B = i1 + 1; sL = sR = lam > 0 ? 1.0 : -1.0; }
    } else {
        const double ta = (-mx - lam) / kap, tb = (mx - lam) / kap;
        A = slab_clampi(floor(fmin(ta, tb)), i0, i1 + 1);
        B = slab_clampi(ceil(fmax(ta, tb)) + 1.0, A, i1 + 1);
        sR = kap > 0 ? 1.0 : -1.0;
        sL = -sR;
    }
    int la = i0, lb = A - 1, ra = B, rb = i1;
    slab_z_interval(alp, bet, sL, (double)zlo, (double)zhi, la, lb);
    slab_z_interval(alp, bet, sR, (double)zlo, (double)zhi, ra, rb);
    // next to the fold either sign may apply: the hull of the two intervals (the exact test decides)
    int ma = A, mb = B - 1, na = A, nb = B - 1;
    slab_z_interval(alp, bet, 1.0, (double)zlo, (double)zhi, ma, mb);
    slab_z_interval(alp, bet, -1.0, (double)zlo, (double)zhi, na, nb);
    if (mb < ma) { ma = na; mb = nb; }
    else if (nb >= na) { ma = ma < na ? ma : na; mb = mb > nb ? mb : nb; }
    const int c1 = lb - la + 1 > 0 ? lb - la + 1 : 0, cm = mb - ma + 1 > 0 ? mb - ma + 1 : 0, c2 = rb - ra + 1 > 0 ? rb - ra + 1 : 0;
    const int p0 = sg.start - i0;
    rec.p1 = p0 + la; rec.c1cm = c1 | (cm << 16); rec.pm = p0 + ma; rec.p2 = p0 + ra;
    return c1 + cm + c2;
}

// list position of candidate u (0 <= u < the count seg_intervals returned)
THB_HD int slab_element(const SlabRec rec, int u)
{
    const int c1 = rec.c1cm & 0xffff, cm = rec.c1cm >> 16;
    if (u < c1) return rec.p1 + u;
    u -= c1;
    if (u < cm) return rec.pm + u;
    return rec.p2 + (u - cm);
}

}  // namespace thb
