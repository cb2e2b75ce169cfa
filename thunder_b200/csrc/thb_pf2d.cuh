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

// Particle::resample(nOut, PAR_C) (src/Particle.cpp:1296-1341) on plain arrays: shuffle (gsl_ran_shuffle = Fisher-Yates), top
// class = the class of the largest uC, w *= u, systematic resampling of nOut classes with one uniform u0 in [0, 1/nOut), new
// prior 1 / uC of the source (PARTICLE_PRIOR_ONE).  c / wC / uC [nIn] are permuted in place; returns the top class.
THB_HD int resample_C(int* c, double* wC, double* uC, int nIn, int nOut, int* cOut, double* wOut, Rng& g)
{
    for (int i = nIn - 1; i > 0; --i) {
        const int j = (int)g.uniform_int((uint32_t)(i + 1));
        if (j != i) {
            const int ci = c[i]; c[i] = c[j]; c[j] = ci;
            double x = wC[i]; wC[i] = wC[j]; wC[j] = x;
            x = uC[i]; uC[i] = uC[j]; uC[j] = x;
        }
    }
    int top = 0;
    for (int i = 1; i < nIn; ++i) if (uC[i] > uC[top]) top = i;
    const int topC = c[top];
    double s = 0.0;
    for (int i = 0; i < nIn; ++i) { wC[i] *= uC[i]; s += wC[i]; }
    for (int i = 0; i < nIn; ++i) wC[i] /= s;
    double last = 0.0;
    for (int i = 0; i < nIn; ++i) last += wC[i];
    const double u0 = g.uniform() * (1.0 / nOut);
    int i = 0;
    double cum = wC[0];
    for (int j = 0; j < nOut; ++j) {
        const double uj = u0 + j * 1.0 / nOut;
        while (i < nIn - 1 && uj > cum / last) { ++i; cum += wC[i]; }
        cOut[j] = c[i];
        wOut[j] = 1.0 / uC[i];
    }
    return topC;
}

}  // namespace pf
}  // namespace thb
