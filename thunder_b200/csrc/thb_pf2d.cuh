// thb_pf2d.cuh - MODE_2D rotation operators of the particle filter (host + device inline), the in-plane twins of the ACG
// operators in thb_pf.cuh.  NOT wired into a kernel yet: the 2D runs of this round drive the
// filter on the host; this header is the verified building block for the device version (tests/test_pf_host.py compares it
// with the reference's DirectionalStat functions on CPU).
//
// In MODE_2D a rotation is the unit vector (cos phi, sin phi) kept in the first two components of the particle's
// quaternion slots (src/Particle.cpp:100-120, 1013-1016, 1160-1175):
//   perturb   r_i <- r_i (x) d_i,  d_i ~ von Mises-like VMS((1, 0), k = min(PERTURB_K_MAX, k1 * pf))   (quaternion_mul of two
//             (c, s, 0, 0) vectors = complex multiplication)
//   calVari   k1 = 1 - | mean(r_i) |                                                                   (inferVMS)
//   sampleVMS / inferVMS / pdfVMS: src/Geometry/DirectionalStat.cpp:252-384; the concentration the reference samples with is
//   kappa(k) = (1 - k)(1 + 2k - k^2) / (k (2 - k)), uniform on the circle below kappa = 0.1, Best-Fisher rejection above.
#pragma once
#include "thb_pf.cuh"

namespace thb {
namespace pf {

constexpr double PERTURB_K_MAX_2D = 1.0;     // include/Particle.h:64

THB_HD double vms_kappa(double k) { return (1.0 - k) * (1.0 + 2.0 * k - k * k) / k / (2.0 - k); }

// one draw about mu = (1, 0): returns (c, s)
THB_HD void sample_vms(Rng& g, double k, double& c, double& s)
{
    const double kappa = vms_kappa(k);
    if (kappa < 1e-1) {                           // gsl_ran_dir_2d: uniform on the circle
        double sn, cs;
        sincos(6.283185307179586 * g.uniform(), &sn, &cs);
        c = cs; s = sn;
        return;
    }
    const double a = 1.0 + sqrt(1.0 + 4.0 * kappa * kappa);
    const double b = (a - sqrt(2.0 * a)) / (2.0 * kappa);
    const double r = (1.0 + b * b) / (2.0 * b);
    double f;
    for (int it = 0; it < 10000; ++it) {          // acceptance probability is > 0.65 for every kappa
        const double z = cos(3.14159265358979323846 * g.uniform());
        f = (1.0 + r * z) / (r + z);
        const double cc = kappa * (r - f);
        const double u2 = g.uniform();
        if (cc * (2.0 - cc) > u2) break;
        if (log(cc / u2) + 1.0 - cc >= 0.0) break;
    }
    const double d = sqrt((1.0 - f) * (f + 1.0));   // mu = (1, 0): delta0 = 0, delta1 = d
    c = f;
    s = g.uniform() > 0.5 ? -d : d;
}

// inferVMS: mean direction and k = 1 - R of the mLR unit vectors in components 0, 1
THB_HD void infer_vms(const View& v, double mu[2], double& k)
{
    double m0 = 0.0, m1 = 0.0;
    for (int i = 0; i < v.mLR; ++i) { m0 += v.R(i, 0); m1 += v.R(i, 1); }
    const double nrm = sqrt(m0 * m0 + m1 * m1);
    const double R = nrm / v.mLR;
    mu[0] = m0 / nrm; mu[1] = m1 / nrm;
    k = 1.0 - R;
}

THB_HD double pdf_vms(const double x[2], const double mu[2], double k)
{
    const double kappa = vms_kappa(k);
    if (kappa < 5.0) {
        // I0 by its power series (converges in < 30 terms for kappa < 5)
        double i0 = 1.0, t = 1.0;
        for (int j = 1; j < 60; ++j) { t *= (kappa / (2.0 * j)) * (kappa / (2.0 * j)); i0 += t; if (t < 1e-17 * i0) break; }
        return exp(kappa * (x[0] * mu[0] + x[1] * mu[1])) / (6.283185307179586 * i0);
    }
    const double dx = x[0] - mu[0], dy = x[1] - mu[1], sd = sqrt(1.0 / kappa), dist = sqrt(dx * dx + dy * dy);
    return exp(-0.5 * dist * dist / (sd * sd)) / (sd * 2.5066282746310002);
}

// Particle::perturb(pf, PAR_R) in MODE_2D
THB_HD void perturb_R_2d(const View& v, double pfac, Rng& g)
{
    const double k = fmin(PERTURB_K_MAX_2D, v.S(S_K1) * pfac);
    for (int i = 0; i < v.mLR; ++i) {
        double c, s;
        sample_vms(g, k, c, s);
        const double a = v.R(i, 0), b = v.R(i, 1);
        v.R(i, 0) = a * c - b * s;                  // quaternion_mul((a, b, 0, 0), (c, s, 0, 0))
        v.R(i, 1) = a * s + b * c;
        v.R(i, 2) = 0.0; v.R(i, 3) = 0.0;
    }
}

// Particle::calVari(PAR_R) in MODE_2D
THB_HD void cal_vari_R_2d(const View& v)
{
    double mu[2], k;
    infer_vms(v, mu, k);
    v.S(S_K1) = k;
}

// Particle::balanceWeight(PAR_R) in MODE_2D (src/Particle.cpp:2315-2329): w_i = 1 / pdfVMS(r_i; mu, k), (mu, k) inferred from
// the support itself
THB_HD void balance_R_2d(const View& v)
{
    double mu[2], k;
    infer_vms(v, mu, k);
    for (int i = 0; i < v.mLR; ++i) {
        const double x[2] = {v.R(i, 0), v.R(i, 1)};
        v.WR(i) = 1.0 / pdf_vms(x, mu, k);
    }
}

// Particle::resample(nOut, PAR_C) (src/Particle.cpp:1296-1341) on plain arrays: shuffle (gsl_ran_shuffle of the identity gives s,
// then new[s(i)] = old[i], src/Particle.cpp:2208-2228), top class = the class of the largest uC, w *= u, systematic resampling of
// nOut classes with one uniform u0 in [0, 1/nOut), new prior 1 / uC of the source (PARTICLE_PRIOR_ONE).  c / wC / uC [nIn] are
// permuted in place (tmp: 4 nIn doubles of scratch); returns the top class.
THB_HD int resample_C(int* c, double* wC, double* uC, int nIn, int nOut, int* cOut, double* wOut, Rng& g, double* tmp = nullptr)
{
    double local[4 * 32];
    if (!tmp) tmp = local;                      // nIn <= 32 without caller scratch (THB_MAX_SLOTS classes)
    double* sidx = tmp; double* c2 = tmp + nIn; double* w2 = tmp + 2 * nIn; double* u2 = tmp + 3 * nIn;
    for (int i = 0; i < nIn; ++i) sidx[i] = (double)i;
    for (int i = nIn - 1; i > 0; --i) {
        const int j = (int)g.uniform_int((uint32_t)(i + 1));
        const double x = sidx[i]; sidx[i] = sidx[j]; sidx[j] = x;
    }
    for (int i = 0; i < nIn; ++i) { const int d = (int)sidx[i]; c2[d] = (double)c[i]; w2[d] = wC[i]; u2[d] = uC[i]; }
    for (int i = 0; i < nIn; ++i) { c[i] = (int)c2[i]; wC[i] = w2[i]; uC[i] = u2[i]; }
    int top = 0;
    for (int i = 1; i < nIn; ++i) if (uC[i] > uC[top]) top = i;
    const int topC = c[top];
    double s = 0.0;
    for (int i = 0; i < nIn; ++i) { wC[i] *= uC[i]; s += wC[i]; }
    double cum = 0.0;
    for (int i = 0; i < nIn; ++i) { wC[i] /= s; cum += wC[i]; w2[i] = cum; }     // w2 = cdf
    const double last = w2[nIn - 1];
    const double u0 = g.flat(0.0, 1.0 / nOut);
    int i = 0;
    for (int j = 0; j < nOut; ++j) {
        const double uj = u0 + j * 1.0 / nOut;
        while (i < nIn - 1 && uj > w2[i] / last) ++i;
        cOut[j] = c[i];
        wOut[j] = 1.0 / uC[i];
    }
    return topC;
}

// ------------------------------------------------------------------------------------------------
// From the global scan to the support of the local phases: the per-image logic of src/Optimiser.cpp:921-1075 after the scan
// (one shared grid of nR rotations x nT translations, uniform priors of Particle::reset; the scan's marginal weights wC[nK],
// wR[nR], wT[nT] of this image):
//   setUC, setPeakFactor(PAR_C) (PARTICLE_PEAK_FACTOR_C: 0.99), keepHalfHeightPeak(PAR_C), resample(nK, PAR_C), rand(cls)
//   setUR / setUT from the marginals of THAT class, setPeakFactor(PAR_R) (the ratio of the value at rank nR / 2 (2D), nR / 8 (3D)
//   to the largest, clamped to [1e-3, 0.5]), keepHalfHeightPeak(PAR_R)                    (OPTIMISER_PEAK_FACTOR_T is off)
//   resample(mLR, PAR_R), resample(mLT, PAR_T): shuffle of the grid, systematic resampling down to the support sizes
//   calVari(PAR_R), calVari(PAR_T), floors on k1..k3 / s0, s1 (OPTIMISER_SCAN_SET_MIN_STD_WITH_PERTURB)
// The view v receives the support (R, T, WR, WT, scalars); scratch: perm / cdf / u of max(nR, nT, 4 nK) doubles each.
// Returns the class.  Random draws in the reference's order (GSL's algorithms, see thb_pf.cuh).
// ------------------------------------------------------------------------------------------------
THB_HD double kth_largest(double* a, int n, int k)      // value of rank k (0 = largest); a is permuted
{
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const double piv = a[(lo + hi) >> 1];
        int i = lo, j = hi;
        while (i <= j) {
            while (a[i] > piv) ++i;
            while (a[j] < piv) --j;
            if (i <= j) { const double x = a[i]; a[i] = a[j]; a[j] = x; ++i; --j; }
        }
        if (k <= j) hi = j; else if (k >= i) lo = i; else return a[k];
    }
    return a[k];
}

// shuffle + systematic resampling of nIn grid points (prior 1 / nIn each, likelihood weights u[nIn]) down to nOut indices into
// the grid (the new priors are 1 / u of the chosen points, PARTICLE_PRIOR_ONE), *topIdx = grid index of the largest weight
THB_HD void resample_grid(const double* u, int nIn, int nOut, int* outIdx, int* topIdx, double* perm, double* cdf, Rng& g)
{
    // shuffle: s = gsl_ran_shuffle(identity); new[s(i)] = old[i]  ->  perm[pos] = old index at that position
    for (int i = 0; i < nIn; ++i) cdf[i] = (double)i;                // cdf doubles as s
    for (int i = nIn - 1; i > 0; --i) {
        const int j = (int)g.uniform_int((uint32_t)(i + 1));
        const double x = cdf[i]; cdf[i] = cdf[j]; cdf[j] = x;
    }
    for (int i = 0; i < nIn; ++i) perm[(int)cdf[i]] = (double)i;
    int top = 0;
    for (int i = 1; i < nIn; ++i) if (u[(int)perm[i]] > u[(int)perm[top]]) top = i;
    *topIdx = (int)perm[top];
    double s = 0.0;
    for (int i = 0; i < nIn; ++i) s += (1.0 / nIn) * u[(int)perm[i]];
    double cum = 0.0;
    for (int i = 0; i < nIn; ++i) { cum += (1.0 / nIn) * u[(int)perm[i]] / s; cdf[i] = cum; }
    const double last = cdf[nIn - 1];
    const double u0 = g.flat(0.0, 1.0 / nOut);
    int i = 0;
    for (int j = 0; j < nOut; ++j) {
        const double uj = u0 + j * 1.0 / nOut;
        while (i < nIn - 1 && uj > cdf[i] / last) ++i;
        outIdx[j] = (int)perm[i];
    }
}

THB_HD int from_scan(const View& v, Rng& g, int mode2D, int nK, int nR, int nT, const double* gridR, int qc, const double* gridT,
                     const float* wC, const float* wRk, size_t strideRk, const float* wTk, size_t strideTk, double kFloor, double sFloor,
                     double* perm, double* cdf, double* u, int* idxR /* [mLR] */, int* idxT /* [mLT] */)
{
    // ---- class
    int cls = 0;
    {
        int c[32], cOut[32];
        double wc[32], uc[32], wOut[32];
        for (int k = 0; k < nK; ++k) { c[k] = k; wc[k] = 1.0 / nK; uc[k] = (double)wC[k]; }
        int top = 0;
        for (int k = 1; k < nK; ++k) if (uc[k] > uc[top]) top = k;
        const double hh = uc[top] * (1.0 - 1e-2);                  // PEAK_FACTOR_C
        for (int k = 0; k < nK; ++k) uc[k] = uc[k] < hh ? 0.0 : uc[k] - hh;
        resample_C(c, wc, uc, nK, nK, cOut, wOut, g, u);
        cls = cOut[g.uniform_int((uint32_t)nK)];
    }
    // ---- rotations: likelihood weights of the chosen class, peak factor from the rank statistics, half-height cut
    const float* wR = wRk + (size_t)cls * strideRk;
    const float* wT = wTk + (size_t)cls * strideTk;
    for (int i = 0; i < nR; ++i) { u[i] = (double)wR[i]; cdf[i] = u[i]; }
    const double umax = kth_largest(cdf, nR, 0);
    const int rank = mode2D ? nR / 2 : nR / 8;
    double pfR = kth_largest(cdf, nR, rank) / umax;
    pfR = fmax(1e-3, fmin(0.5, pfR));
    v.S(S_PEAKR) = pfR;
    {
        const double hh = umax * pfR;
        for (int i = 0; i < nR; ++i) u[i] = u[i] < hh ? 0.0 : u[i] - hh;
    }
    int topR = 0, topT = 0;
    resample_grid(u, nR, v.mLR, idxR, &topR, perm, cdf, g);
    for (int j = 0; j < v.mLR; ++j) {
        const int gi = idxR[j];
        for (int c = 0; c < 4; ++c) v.R(j, c) = c < qc ? gridR[(size_t)gi * qc + c] : 0.0;
        v.WR(j) = 1.0 / u[gi];
    }
    for (int c = 0; c < 4; ++c) v.S(S_TOPR + c) = c < qc ? gridR[(size_t)topR * qc + c] : 0.0;
    // ---- translations (no peak factor: OPTIMISER_PEAK_FACTOR_T is off)
    for (int i = 0; i < nT; ++i) u[i] = (double)wT[i];
    resample_grid(u, nT, v.mLT, idxT, &topT, perm, cdf, g);
    for (int j = 0; j < v.mLT; ++j) {
        const int gi = idxT[j];
        v.T(j, 0) = gridT[2 * (size_t)gi]; v.T(j, 1) = gridT[2 * (size_t)gi + 1];
        v.WT(j) = 1.0 / u[gi];
    }
    v.S(S_TOPT) = gridT[2 * (size_t)topT]; v.S(S_TOPT + 1) = gridT[2 * (size_t)topT + 1];
    norm_w(v);
    // ---- variances and their floors
    if (mode2D) cal_vari_R_2d(v); else cal_vari_R(v, g);
    cal_vari_T(v);
    // setK1(TSGSL_MAX_RFLOAT(floor, k1())) ...: the reference takes the maximum in RFLOAT, i.e. the variances pass through fp32 here
    v.S(S_K1) = (double)fmaxf((float)kFloor, (float)v.S(S_K1));
    if (!mode2D) { v.S(S_K2) = (double)fmaxf((float)kFloor, (float)v.S(S_K2)); v.S(S_K3) = (double)fmaxf((float)kFloor, (float)v.S(S_K3)); }
    v.S(S_S0) = (double)fmaxf((float)sFloor, (float)v.S(S_S0));
    v.S(S_S1) = (double)fmaxf((float)sFloor, (float)v.S(S_S1));
    v.S(S_SCORE) = 1.0; v.S(S_NPHASE) = 0.0; v.S(S_NODEC) = 0.0;
    v.S(S_VARIR) = 1.79769313486231570e308; v.S(S_VARIT) = 1.79769313486231570e308; v.S(S_VARID) = 1.79769313486231570e308;
    return cls;
}

}  // namespace pf
}  // namespace thb
