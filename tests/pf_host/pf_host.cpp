// Test-only host build of the particle-filter operators in thunder_b200/csrc/thb_pf.cuh (the same
// source the CUDA kernels compile, one particle per thread there, one particle per call here), so
// that the CPU suite can compare them with the reference's Particle / DirectionalStat classes.
// Compiled by tests/test_pf_host.py with g++; never part of libthunder_b200.so.
#include <cstring>
#include <vector>
#include <algorithm>
#include "thb_pf.cuh"
#include "thb_pf2d.cuh"

using namespace thb;

extern "C" {

// arrays are component-major for one particle: r[c*mLR+i], t[c*mLT+i]; scal[20]; uR/uT double in/out
int pfh_run_s(int op, double arg, int mLR, int mLT, double* r, double* t, double* wR, double* wT, double* uR, double* uT,
              const float* uRf, const float* uTf, double* scal, double transS, double transQ, unsigned long long seed,
              unsigned long long stream, unsigned long long epoch);
int pfh_run(int op, double arg, int mLR, int mLT, double* r, double* t, double* wR, double* wT, double* uR, double* uT,
            const float* uRf, const float* uTf, double* scal, double transS, double transQ, unsigned long long seed,
            unsigned long long epoch)
{
    return pfh_run_s(op, arg, mLR, mLT, r, t, wR, wT, uR, uT, uRf, uTf, scal, transS, transQ, seed, 0, epoch);
}

// the same with the random stream (= particle index on the device) given
int pfh_run_s(int op, double arg, int mLR, int mLT, double* r, double* t, double* wR, double* wT, double* uR, double* uT,
              const float* uRf, const float* uTf, double* scal, double transS, double transQ, unsigned long long seed,
              unsigned long long stream, unsigned long long epoch)
{
    std::vector<double> r2(4 * mLR), t2(2 * mLT), w2(mLR > mLT ? mLR : mLT), w3(w2.size()), w4(w2.size());
    pf::View v;
    v.r = r; v.t = t; v.wR = wR; v.wT = wT; v.uR = uR; v.uT = uT; v.scal = scal;
    v.r2 = r2.data(); v.t2 = t2.data(); v.w2 = w2.data(); v.w3 = w3.data(); v.w4 = w4.data();
    v.n = 1; v.p = 0; v.mLR = mLR; v.mLT = mLT; v.lane = -1;
    v.mLD = 0; v.d = v.wD = v.uD = nullptr; v.mode2D = 0;
    pf::Rng g;
    g.init(seed, stream, epoch);
    switch (op) {
        case 1: pf::perturb_R(v, arg, g); break;
        case 2: pf::perturb_T(v, arg, transS, transQ, g); break;
        case 3: pf::set_u_keep_peak(v, uRf, uTf); break;
        case 4: pf::rank1st(v); break;
        case 5: pf::cal_vari(v, g); break;
        case 6: pf::resample_R(v, g); pf::resample_T(v, g); pf::norm_w(v); break;
        case 7: pf::balance_R(v); break;
        case 8: pf::balance_T(v); break;
        case 100: {   // load: arg unused; scal[0..2] = k, scal[3..4] = s, scal[6..9] = q, scal[10..11] = t on input
            double q[4] = {scal[6], scal[7], scal[8], scal[9]}, tt[2] = {scal[10], scal[11]};
            pf::load(v, q, scal[0], scal[1], scal[2], tt, scal[3], scal[4], g);
            break;
        }
        case 101: return pf::stop_rule(v, (int)arg, 3, 0.95, 1) ? 1 : 0;
        case 102: {   // Particle::rand x arg (<= mLR): (class, rotation, translation, defocus) indices; uR[i] = 1000 rotation + translation
            const int m = (int)arg;
            for (int i = 0; i < m && i < mLR; ++i) {
                (void)g.uniform_int(1);
                const double a = (double)g.uniform_int((uint32_t)mLR);
                const double b = (double)g.uniform_int((uint32_t)mLT);
                (void)g.uniform_int(1);
                uR[i] = 1000.0 * a + b;
            }
            break;
        }
        default: return -1;
    }
    return 0;
}

void pfh_infer_acg(int mLR, double* r, double* A16, double* mean4)
{
    pf::View v;
    memset(&v, 0, sizeof(v));
    v.r = r; v.n = 1; v.p = 0; v.mLR = mLR; v.lane = -1;
    pf::infer_acg(v, A16);
    if (mean4) pf::sym4_top_eigvec(A16, mean4);
}

void pfh_rng(unsigned long long seed, unsigned long long stream, unsigned long long epoch, int n, double* uni, double* nor)
{
    pf::Rng g;
    g.init(seed, stream, epoch);
    for (int i = 0; i < n; i++) uni[i] = g.uniform();
    for (int i = 0; i < n; i++) nor[i] = g.normal();
}

// the defocus dimension (CTF search): d[mLD + 1] (last = the top one), wD / uD [mLD], scal[20]
int pfh_run_d(int op, double arg, int mLD, double* d, double* wD, double* uD, const float* uDf, double* scal, unsigned long long seed,
              unsigned long long stream, unsigned long long epoch)
{
    const int mw = mLD > 4 ? mLD : 4;
    std::vector<double> r2(4 * mw), w2(mw), w3(mw), w4(mw), wR(mw, 1.0), wT(mw, 1.0);
    pf::View v;
    memset(&v, 0, sizeof(v));
    v.d = d; v.wD = wD; v.uD = uD; v.scal = scal; v.r2 = r2.data(); v.w2 = w2.data(); v.w3 = w3.data(); v.w4 = w4.data();
    v.wR = wR.data(); v.wT = wT.data();
    v.n = 1; v.p = 0; v.mLR = mw; v.mLT = 1; v.mLD = mLD; v.lane = -1;
    pf::Rng g;
    g.init(seed, stream, epoch);
    switch (op) {
        case 200: pf::init_D(v, arg, g); break;
        case 201: pf::perturb_D(v, arg, g); break;
        case 202: pf::cal_vari_D(v); break;
        case 203: pf::resample_D(v, g); pf::norm_w(v); break;
        case 204: pf::balance_D(v); break;
        case 205: for (int i = 0; i < mLD; ++i) v.UD(i) = (double)uDf[i]; pf::rank1st_D(v); break;
        default: return -1;
    }
    return 0;
}

// the hand-over from the global scan (thb_pf2d.cuh: from_scan) for one image; r / t / wR / wT / scal as in pfh_run
int pfh_from_scan(int mode2D, int nK, int nR, int nT, const double* gridR, const double* gridT, const float* wC, const float* wR,
                  const float* wT, int mLR, int mLT, double kFloor, double sFloor, double* r, double* t, double* owR, double* owT,
                  double* scal, unsigned long long seed, unsigned long long stream, unsigned long long epoch)
{
    const int mw = mLR > mLT ? mLR : mLT;
    const int nMax = std::max(std::max(nR, nT), 4 * nK);
    std::vector<double> r2(4 * mLR), t2(2 * mLT), w2(mw), w3(mw), w4(mw), uR(mLR), uT(mLT), sc(3 * (size_t)nMax);
    std::vector<int> ix(mLR + mLT);
    pf::View v;
    memset(&v, 0, sizeof(v));
    v.r = r; v.t = t; v.wR = owR; v.wT = owT; v.uR = uR.data(); v.uT = uT.data(); v.scal = scal;
    v.r2 = r2.data(); v.t2 = t2.data(); v.w2 = w2.data(); v.w3 = w3.data(); v.w4 = w4.data();
    v.n = 1; v.p = 0; v.mLR = mLR; v.mLT = mLT; v.lane = -1; v.mode2D = mode2D;
    pf::Rng g;
    g.init(seed, stream, epoch);
    return pf::from_scan(v, g, mode2D, nK, nR, nT, gridR, mode2D ? 2 : 4, gridT, wC, wR, (size_t)nR, wT, (size_t)nT, kFloor, sFloor, sc.data(),
                         sc.data() + nMax, sc.data() + 2 * nMax, ix.data(), ix.data() + mLR);
}

// the GSL entry points restated in pf::Rng, by kind as ref_rng_draw of oracle/ref_harness.cpp
void pfh_rng_draw(unsigned long long seed, unsigned long long stream, unsigned long long epoch, int kind, int n, double a, double b,
                  double* out)
{
    pf::Rng g;
    g.init(seed, stream, epoch);
    for (int i = 0; i < n; i++)
        switch (kind) {
            case 0: out[i] = g.uniform(); break;
            case 1: out[i] = g.gaussian(a); break;
            case 2: out[i] = (double)g.uniform_int((uint32_t)a); break;
            case 3: out[i] = g.flat(a, b); break;
            case 4: g.bivariate_gaussian(a, b, 0.3, out[2 * i], out[2 * i + 1]); break;
        }
}

// ---- MODE_2D rotation operators (thb_pf2d.cuh)
void pfh_sample_vms(unsigned long long seed, double k, int n, double* cs)
{
    pf::Rng g;
    g.init(seed, 0, 1);
    for (int i = 0; i < n; i++) pf::sample_vms(g, k, cs[2 * i], cs[2 * i + 1]);
}

void pfh_infer_vms(int mLR, double* r, double* mu2, double* k)
{
    pf::View v;
    memset(&v, 0, sizeof(v));
    v.r = r; v.n = 1; v.p = 0; v.mLR = mLR; v.lane = -1;
    pf::infer_vms(v, mu2, *k);
}

double pfh_pdf_vms(const double* x2, const double* mu2, double k) { return pf::pdf_vms(x2, mu2, k); }

void pfh_perturb_r_2d(int mLR, double* r, double k1, double pfac, unsigned long long seed)
{
    std::vector<double> scal(20, 0.0);
    pf::View v;
    memset(&v, 0, sizeof(v));
    v.r = r; v.scal = scal.data(); v.n = 1; v.p = 0; v.mLR = mLR; v.lane = -1;
    scal[pf::S_K1] = k1;
    pf::Rng g;
    g.init(seed, 0, 2);
    pf::perturb_R_2d(v, pfac, g);
}

void pfh_balance_r_2d(int mLR, double* r, double* wR)
{
    pf::View v;
    memset(&v, 0, sizeof(v));
    v.r = r; v.wR = wR; v.n = 1; v.p = 0; v.mLR = mLR; v.lane = -1;
    pf::balance_R_2d(v);
}

int pfh_resample_c(int nIn, int* c, double* wC, double* uC, int nOut, int* cOut, double* wOut, unsigned long long seed)
{
    pf::Rng g;
    g.init(seed, 0, 3);
    return pf::resample_C(c, wC, uC, nIn, nOut, cOut, wOut, g);
}

}  // extern "C"
