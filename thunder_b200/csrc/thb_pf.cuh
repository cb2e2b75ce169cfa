// thb_pf.cuh - the per-image particle filter (host side of the reference: class Particle),
// restated as serial per-particle code that runs one particle per CUDA thread on the device-
// resident SoA state.  Everything is host+device inline so that tests/ can compile the same
// operators with g++ and compare them with the reference's Particle / DirectionalStat.
//
// reference (paths relative to the THUNDER tree):
//   perturb            src/Particle.cpp:1149-1289        resample         src/Particle.cpp:1291-1478
//   calVari            src/Particle.cpp:1004-1142        keepHalfHeightPeak :1964-2002
//   calRank1st         src/Particle.cpp:990-1002         balanceWeight    :2309-2410
//   reCentre           src/Particle.cpp:2473-2495        shuffle          :2202-2300
//   rand               src/Particle.cpp:2109-2200        load             :401-556
//   variR/variT/compressR :611-667
//   sampleACG / inferACG / pdfACG   src/Geometry/DirectionalStat.cpp:19-250
//   quaternion_mul     src/Geometry/Euler.cpp:13-26
// Default Config.h switches in force: PARTICLE_PRIOR_ONE, PARTICLE_RECENTRE(_TRANSQ),
// PARTICLE_ROT_MEAN_USING_STAT_CAL_VARI / _PERTURB, PARTICLE_BALANCE_WEIGHT_R/T; PARTICLE_RHO off.
// Random numbers.  The BIT generator is a counter-based Philox4x32-10 keyed by (seed, particle, epoch) instead of the
// reference's thread-local urandom-seeded mt19937 (src/Functions/Random.cpp:51-100); everything ABOVE the bit generator
// is GSL 2.4's published algorithm, the one the reference calls (external/packages/gsl-2.4, pinned by the reference tree):
//   gsl_rng_uniform        get() / 2^32 for a 32-bit generator          rng/gsl_rng.h:164-168, rng/mt.c (mt_get_double)
//   gsl_rng_uniform_pos    rejects 0                                     rng/gsl_rng.h:170-181
//   gsl_rng_uniform_int    scale = range / n, reject k >= n              rng/gsl_rng.h:189-212
//   gsl_ran_gaussian       polar Box-Muller, returns sigma y sqrt(..)    randist/gauss.c:47-65
//   gsl_ran_bivariate_gaussian  one polar pair for both coordinates     randist/bigauss.c:34-57
//   gsl_ran_flat           a (1 - u) + b u                               randist/flat.c:32-40
//   gsl_ran_shuffle        i = n-1 .. 1: swap(i, uniform_int(i + 1))     randist/shuffle.c:66-78
// and the operators below consume them in the reference's order.  With the reference's engine swapped for the same bit
// generator (the test harness of the reference build does that) the two particle filters therefore see the SAME random numbers, and
// every operator - stochastic ones included - is compared exactly (tests/test_pf_host.py, tests/test_gpu_iteration.py).
#pragma once
#include <math.h>
#include <stdint.h>
#include "thb_math.cuh"

namespace thb {
namespace pf {

enum { S_K1 = 0, S_K2, S_K3, S_S0, S_S1, S_RHO, S_TOPR, S_TOPT = 10, S_SCORE = 12, S_NPHASE = 13, S_VARIR = 14, S_VARIT = 15,
       S_PEAKR = 16, S_NODEC = 17, S_VARID = 18, S_SD = 19 /* sigma of the defocus factors (_s) */, S_SPARE = 19, S_COUNT = 20 };

// ------------------------------------------------------------------------------------------------
struct Rng {
    uint32_t k0, k1;
    uint32_t c[4];
    uint32_t o[4];
    int have;

    THB_HD void init(uint64_t seed, uint64_t stream, uint64_t epoch)
    {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        c[0] = 0; c[1] = (uint32_t)epoch; c[2] = (uint32_t)stream; c[3] = (uint32_t)(stream >> 32) ^ (uint32_t)(epoch >> 32);
        have = 0;
    }
    THB_HD static void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
    {
        const uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
    }
    THB_HD void block()
    {
        uint32_t x0 = c[0], x1 = c[1], x2 = c[2], x3 = c[3], a = k0, b = k1;
        for (int r = 0; r < 10; ++r) {
            uint32_t h0, l0, h1, l1;
            mulhilo(0xD2511F53u, x0, h0, l0);
            mulhilo(0xCD9E8D57u, x2, h1, l1);
            const uint32_t y0 = h1 ^ x1 ^ a, y1 = l1, y2 = h0 ^ x3 ^ b, y3 = l0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        o[0] = x0; o[1] = x1; o[2] = x2; o[3] = x3;
        c[0]++;
        have = 4;
    }
    // block b of the stream without touching the state (the generator is counter-based: draw d is element 3 - d % 4 of block d / 4)
    THB_HD void block_at(uint32_t b, uint32_t out[4]) const
    {
        uint32_t x0 = b, x1 = c[1], x2 = c[2], x3 = c[3], ka = k0, kb = k1;
        for (int r = 0; r < 10; ++r) {
            uint32_t h0, l0, h1, l1;
            mulhilo(0xD2511F53u, x0, h0, l0);
            mulhilo(0xCD9E8D57u, x2, h1, l1);
            const uint32_t y0 = h1 ^ x1 ^ ka, y1 = l1, y2 = h0 ^ x3 ^ kb, y3 = l0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
        out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
    }
    THB_HD uint32_t position() const { return 4u * c[0] - (uint32_t)have; }     // draws consumed so far
    THB_HD void seek(uint32_t pos)                                              // continue with draw `pos`
    {
        const uint32_t b = pos >> 2, e = pos & 3u;
        if (e == 0) { c[0] = b; have = 0; return; }
        block_at(b, o);
        c[0] = b + 1;
        have = 4 - (int)e;
    }
    THB_HD uint32_t u32()
    {
        if (have == 0) block();
        return o[--have];
    }
    THB_HD double uniform() { return (double)u32() * (1.0 / 4294967296.0); }   // gsl_rng_uniform: [0, 1), 32 bits
    THB_HD double uniform_pos()
    {
        double x;
        do { x = uniform(); } while (x == 0.0);
        return x;
    }
    THB_HD uint32_t uniform_int(uint32_t n)     // gsl_rng_uniform_int: consumes at least one draw, also for n = 1
    {
        const uint32_t scale = 0xffffffffu / n;
        uint32_t k;
        do { k = u32() / scale; } while (k >= n);
        return k;
    }
    THB_HD double flat(double a, double b)
    {
        const double u = uniform();
        return a * (1.0 - u) + b * u;
    }
    THB_HD double gaussian(double sigma)
    {
        double x, y, r2;
        do {
            x = -1.0 + 2.0 * uniform_pos();
            y = -1.0 + 2.0 * uniform_pos();
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);
        return sigma * y * sqrt(-2.0 * log(r2) / r2);
    }
    THB_HD void bivariate_gaussian(double sx, double sy, double rho, double& x, double& y)
    {
        double u, v, r2;
        do {
            u = -1.0 + 2.0 * uniform();
            v = -1.0 + 2.0 * uniform();
            r2 = u * u + v * v;
        } while (r2 > 1.0 || r2 == 0.0);
        const double scale = sqrt(-2.0 * log(r2) / r2);
        x = sx * u * scale;
        y = sy * (rho * u + sqrt(1.0 - rho * rho) * v) * scale;
    }
    THB_HD double normal() { return gaussian(1.0); }
};

// ------------------------------------------------------------------------------------------------
// strided view of one particle in the SoA state
struct View {
    double* r; double* t; double* wR; double* wT; double* uR; double* uT; double* scal;
    double* r2; double* t2; double* w2;      // resampling scratch, same layout as r / t / wR
    double* w3; double* w4;                  // two more rows of max(mLR, mLT) (the shuffle's scatter)
    double* d; double* wD; double* uD;       // CTF search: mLD defocus factors (+ the top one in row mLD of d), null when mLD == 0
    int mLD;
    int mode2D;                              // MODE_2D: rotations are unit vectors (cos, sin, 0, 0), von Mises-like statistics (thb_pf2d.cuh)
    long long n;   // stride between consecutive samples/components = number of particles
    long long p;   // particle index
    int mLR, mLT;
    int lane;      // >= 0: a whole warp executes this particle redundantly and `lane` is this thread's lane (device only)
    THB_HD double& R(int i, int c) const { return r[((long long)c * mLR + i) * n + p]; }
    THB_HD double& T(int i, int c) const { return t[((long long)c * mLT + i) * n + p]; }
    THB_HD double& R2(int i, int c) const { return r2[((long long)c * mLR + i) * n + p]; }
    THB_HD double& T2(int i, int c) const { return t2[((long long)c * mLT + i) * n + p]; }
    THB_HD double& W2(int i) const { return w2[(long long)i * n + p]; }
    THB_HD double& W3(int i) const { return w3[(long long)i * n + p]; }
    THB_HD double& W4(int i) const { return w4[(long long)i * n + p]; }
    THB_HD double& WR(int i) const { return wR[(long long)i * n + p]; }
    THB_HD double& WT(int i) const { return wT[(long long)i * n + p]; }
    THB_HD double& UR(int i) const { return uR[(long long)i * n + p]; }
    THB_HD double& UT(int i) const { return uT[(long long)i * n + p]; }
    THB_HD double& S(int k) const { return scal[(long long)k * n + p]; }
    THB_HD double& D(int i) const { return d[(long long)i * n + p]; }          // i == mLD: the most likely one (_topD)
    THB_HD double& WD(int i) const { return wD[(long long)i * n + p]; }
    THB_HD double& UD(int i) const { return uD[(long long)i * n + p]; }
};

THB_HD void quat_mul(double d[4], const double a[4], const double b[4])
{
    const double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    const double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    const double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    const double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    d[0] = w; d[1] = x; d[2] = y; d[3] = z;
}

// 4x4 inverse by cofactors (row-major); returns the determinant
THB_HD double inv4(const double m[16], double inv[16])
{
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    const double id = 1.0 / det;
    for (int i = 0; i < 16; ++i) inv[i] *= id;
    return det;
}

THB_HD double quad_form(const double Ai[16], const double x[4])
{
    double s = 0.0;
    for (int i = 0; i < 4; ++i) {
        double row = 0.0;
        for (int j = 0; j < 4; ++j) row += Ai[i * 4 + j] * x[j];
        s += x[i] * row;
    }
    return s;
}

// inferACG(dmat44& dst, const dmat4& src): fixed-point iteration, returns the LAST-BUT-ONE iterate
// exactly as the reference does (dst = A).  DirectionalStat.cpp:93-145
THB_HD void infer_acg_serial(const View& v, double A[16])
{
    double B[16];
    for (int i = 0; i < 16; ++i) B[i] = (i % 5 == 0) ? 1.0 : 0.0;
    double diff;
    int iter = 0;
    do {
        for (int i = 0; i < 16; ++i) A[i] = B[i];
        double Ai[16];
        inv4(A, Ai);
        for (int i = 0; i < 16; ++i) B[i] = 0.0;
        double nf = 0.0;
        for (int s = 0; s < v.mLR; ++s) {
            const double x[4] = {v.R(s, 0), v.R(s, 1), v.R(s, 2), v.R(s, 3)};
            const double u = quad_form(Ai, x);
            const double iu = 1.0 / u;
            for (int j = 0; j < 4; ++j)
                for (int k = 0; k < 4; ++k) B[j * 4 + k] += x[j] * x[k] / u;
            nf += iu;
        }
        const double sc = 4.0 / nf;
        diff = 0.0;
        for (int i = 0; i < 16; ++i) {
            B[i] *= sc;
            diff += fabs(A[i] - B[i]);
        }
    } while (diff > 1e-3 && ++iter < 500);
}

#if defined(__CUDA_ARCH__)
// The same iteration with the sum over the support points spread over the 32 lanes of the warp that owns the
// particle (v.lane >= 0): every lane keeps up to 4 support points in registers, the 10 unique entries of the
// symmetric B and the normalisation are butterfly-reduced, all lanes hold identical results.
__device__ inline void infer_acg_warp(const View& v, double A[16])
{
    const int lane = v.lane;
    double xs[4][4];
    int cnt = 0;
    for (int s = lane; s < v.mLR && cnt < 4; s += 32, ++cnt)
        for (int c = 0; c < 4; ++c) xs[cnt][c] = v.R(s, c);
    double B[16];
    for (int i = 0; i < 16; ++i) B[i] = (i % 5 == 0) ? 1.0 : 0.0;
    double diff;
    int iter = 0;
    do {
        for (int i = 0; i < 16; ++i) A[i] = B[i];
        double Ai[16];
        inv4(A, Ai);
        double acc[11];   // 00 01 02 03 11 12 13 22 23 33, sum 1/u
        for (int i = 0; i < 11; ++i) acc[i] = 0.0;
        for (int q = 0; q < cnt; ++q) {
            const double* x = xs[q];
            const double iu = 1.0 / quad_form(Ai, x);
            int e = 0;
            for (int j = 0; j < 4; ++j)
                for (int k = j; k < 4; ++k) acc[e++] += x[j] * x[k] * iu;
            acc[10] += iu;
        }
        for (int s = lane + 128; s < v.mLR; s += 32) {   // supports larger than 128: the rest straight from memory
            const double x[4] = {v.R(s, 0), v.R(s, 1), v.R(s, 2), v.R(s, 3)};
            const double iu = 1.0 / quad_form(Ai, x);
            int e = 0;
            for (int j = 0; j < 4; ++j)
                for (int k = j; k < 4; ++k) acc[e++] += x[j] * x[k] * iu;
            acc[10] += iu;
        }
        for (int i = 0; i < 11; ++i)
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        const double sc = 4.0 / acc[10];
        int e = 0;
        for (int j = 0; j < 4; ++j)
            for (int k = j; k < 4; ++k) {
                B[j * 4 + k] = acc[e] * sc;
                B[k * 4 + j] = acc[e] * sc;
                ++e;
            }
        diff = 0.0;
        for (int i = 0; i < 16; ++i) diff += fabs(A[i] - B[i]);
    } while (diff > 1e-3 && ++iter < 500);
}
#endif

THB_HD void infer_acg(const View& v, double A[16])
{
#if defined(__CUDA_ARCH__)
    if (v.lane >= 0) {
        infer_acg_warp(v, A);
        return;
    }
#endif
    infer_acg_serial(v, A);
}

// eigenvector of the largest eigenvalue of a symmetric 4x4 (cyclic Jacobi); unit norm.
// (the reference uses Eigen::SelfAdjointEigenSolver; the sign of the vector is immaterial, see perturb)
THB_HD void sym4_top_eigvec(const double Ain[16], double vec[4])
{
    double a[16], V[16];
    for (int i = 0; i < 16; ++i) { a[i] = Ain[i]; V[i] = (i % 5 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < 4; ++i)
            for (int j = i + 1; j < 4; ++j) off += a[i * 4 + j] * a[i * 4 + j];
        if (off < 1e-300) break;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 4; ++q) {
                const double apq = a[p * 4 + q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (a[q * 4 + q] - a[p * 4 + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {
                    const double akp = a[k * 4 + p], akq = a[k * 4 + q];
                    a[k * 4 + p] = c * akp - s * akq;
                    a[k * 4 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; ++k) {
                    const double apk = a[p * 4 + k], aqk = a[q * 4 + k];
                    a[p * 4 + k] = c * apk - s * aqk;
                    a[q * 4 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k * 4 + p], vkq = V[k * 4 + q];
                    V[k * 4 + p] = c * vkp - s * vkq;
                    V[k * 4 + q] = s * vkp + c * vkq;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < 4; ++i)
        if (a[i * 4 + i] > a[best * 4 + best]) best = i;
    double nrm = 0.0;
    for (int k = 0; k < 4; ++k) { vec[k] = V[k * 4 + best]; nrm += vec[k] * vec[k]; }
    nrm = 1.0 / sqrt(nrm);
    for (int k = 0; k < 4; ++k) vec[k] *= nrm;
}

THB_HD void acg_mean(const View& v, double mean[4])
{
    double A[16];
    infer_acg(v, A);
    sym4_top_eigvec(A, mean);
}

// left-multiply every rotation by q (conj = use the conjugate of q)
THB_HD void left_mul_all(const View& v, const double q[4], bool conj)
{
    double a[4] = {q[0], conj ? -q[1] : q[1], conj ? -q[2] : q[2], conj ? -q[3] : q[3]};
    for (int i = 0; i < v.mLR; ++i) {
        double x[4] = {v.R(i, 0), v.R(i, 1), v.R(i, 2), v.R(i, 3)}, y[4];
        quat_mul(y, a, x);
        for (int c = 0; c < 4; ++c) v.R(i, c) = y[c];
    }
}

THB_HD void norm_w(const View& v)
{
    double s = 0.0;
    for (int i = 0; i < v.mLR; ++i) s += v.WR(i);
    for (int i = 0; i < v.mLR; ++i) v.WR(i) /= s;
    s = 0.0;
    for (int i = 0; i < v.mLT; ++i) s += v.WT(i);
    for (int i = 0; i < v.mLT; ++i) v.WT(i) /= s;
    if (v.mLD > 0) {
        s = 0.0;
        for (int i = 0; i < v.mLD; ++i) s += v.WD(i);
        for (int i = 0; i < v.mLD; ++i) v.WD(i) /= s;
    }
}

// balanceWeight(PAR_R): wR = 1 / pdfACG(r, A), A = inferACG(r); pdfACG = det^-1/2 (x' A^-1 x)^-2
THB_HD void balance_R(const View& v)
{
    double A[16], Ai[16];
    infer_acg(v, A);
    const double det = inv4(A, Ai);
    const double c = pow(det, -0.5);
    for (int i = 0; i < v.mLR; ++i) {
        const double x[4] = {v.R(i, 0), v.R(i, 1), v.R(i, 2), v.R(i, 3)};
        const double u = quad_form(Ai, x);
        v.WR(i) = 1.0 / (c * pow(u, -2.0));
    }
    norm_w(v);
}

THB_HD void mean_sd(const View& v, int c, double& m, double& sd)
{
    double s = 0.0;
    for (int i = 0; i < v.mLT; ++i) s += v.T(i, c);
    m = s / v.mLT;
    double q = 0.0;
    for (int i = 0; i < v.mLT; ++i) { const double d = v.T(i, c) - m; q += d * d; }
    sd = sqrt(q / (v.mLT - 1));   // gsl_stats_sd_m: N-1 denominator
}

// balanceWeight(PAR_T): wT = 1 / bivariate_gaussian_pdf(t - mean; s0, s1, rho = 0)
THB_HD void balance_T(const View& v)
{
    double m0, m1, s0, s1;
    mean_sd(v, 0, m0, s0);
    mean_sd(v, 1, m1, s1);
    for (int i = 0; i < v.mLT; ++i) {
        const double u = (v.T(i, 0) - m0) / s0, w = (v.T(i, 1) - m1) / s1;
        const double pdf = exp(-(u * u + w * w) / 2.0) / (2.0 * 3.14159265358979323846 * s0 * s1);
        v.WT(i) = 1.0 / pdf;
    }
    norm_w(v);
}

#if defined(__CUDA_ARCH__)
// The same draws and the same arithmetic as the serial loop below, spread over the warp that executes the particle: the polar
// method consumes the stream two draws per attempt, whatever the outcome, so attempt a uses draws pos + 2a, pos + 2a + 1 and the
// k-th ACCEPTED attempt is the k-th Gaussian.  32 attempts are evaluated at a time (one per lane), a ballot ranks the accepted ones.
// A zero draw (uniform_pos re-draws it: probability 2^-32) shifts the pairing; it is left to the serial loop (returns false with
// the state untouched).  The 4 mLR Gaussians are the serial chain that dominated the particle-filter kernel.
__device__ inline bool sample_acg_r2_warp(const View& v, double k1, double k2, double k3, Rng& g)
{
    const unsigned full = 0xffffffffu;
    const int lane = v.lane, need = 4 * v.mLR;
    const uint32_t pos = g.position();
    int accepted = 0;
    uint32_t base = 0, lastAttempt = 0;
    while (accepted < need) {
        const uint32_t a = base + (uint32_t)lane;
        const uint32_t d0 = pos + 2u * a;
        const int e0 = (int)(d0 & 3u);
        uint32_t o0[4];
        g.block_at(d0 >> 2, o0);
        const uint32_t u0 = o0[3 - e0];
        uint32_t u1;
        if (e0 == 3) {
            uint32_t o1[4];
            g.block_at((d0 >> 2) + 1u, o1);
            u1 = o1[3];
        } else
            u1 = o0[2 - e0];
        if (__any_sync(full, u0 == 0u || u1 == 0u)) return false;
        const double x = -1.0 + 2.0 * ((double)u0 * (1.0 / 4294967296.0));
        const double y = -1.0 + 2.0 * ((double)u1 * (1.0 / 4294967296.0));
        const double r2 = x * x + y * y;
        const bool ok = !(r2 > 1.0 || r2 == 0.0);
        const unsigned mask = __ballot_sync(full, ok);
        const int gi = accepted + __popc(mask & ((1u << lane) - 1u));
        if (ok && gi < need) v.R2(gi >> 2, gi & 3) = 1.0 * y * sqrt(-2.0 * log(r2) / r2);
        const unsigned lastMask = __ballot_sync(full, ok && gi == need - 1);
        if (lastMask) lastAttempt = base + (uint32_t)(__ffs((int)lastMask) - 1);
        accepted += __popc(mask);
        base += 32u;
    }
    g.seek(pos + 2u * (lastAttempt + 1u));
    __syncwarp();
    const double l1 = sqrt(k1), l2 = sqrt(k2), l3 = sqrt(k3);
    for (int i = lane; i < v.mLR; i += 32) {
        double d[4];
        d[0] = v.R2(i, 0); d[1] = v.R2(i, 1); d[2] = v.R2(i, 2); d[3] = v.R2(i, 3);
        d[1] *= l1; d[2] *= l2; d[3] *= l3;
        const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
        for (int c = 0; c < 4; ++c) v.R2(i, c) = d[c] / nrm;
    }
    __syncwarp();
    return true;
}
#endif

// sampleACG(dst, k1, k2, k3, n) into the scratch rows R2 (src/Geometry/DirectionalStat.cpp:39-62): L = diag(1, sqrt k) of the LLT
THB_HD void sample_acg_r2(const View& v, double k1, double k2, double k3, Rng& g)
{
#if defined(__CUDA_ARCH__)
    if (v.lane >= 0 && sample_acg_r2_warp(v, k1, k2, k3, g)) return;
#endif
    const double l1 = sqrt(k1), l2 = sqrt(k2), l3 = sqrt(k3);
    for (int i = 0; i < v.mLR; ++i) {
        double d[4];
        d[0] = g.gaussian(1.0); d[1] = g.gaussian(1.0); d[2] = g.gaussian(1.0); d[3] = g.gaussian(1.0);
        d[1] *= l1; d[2] *= l2; d[3] *= l3;
        const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
        for (int c = 0; c < 4; ++c) v.R2(i, c) = d[c] / nrm;
    }
}

// perturb(pf, PAR_R), MODE_3D
THB_HD void perturb_R(const View& v, double pfac, Rng& g)
{
    sample_acg_r2(v, pfac * pfac * fmin(1.0, v.S(S_K1)), pfac * pfac * fmin(1.0, v.S(S_K2)), pfac * pfac * fmin(1.0, v.S(S_K3)), g);
    double mean[4];
    acg_mean(v, mean);
    const double mc[4] = {mean[0], -mean[1], -mean[2], -mean[3]};
    for (int i = 0; i < v.mLR; ++i) {
        const double d[4] = {v.R2(i, 0), v.R2(i, 1), v.R2(i, 2), v.R2(i, 3)};
        double x[4] = {v.R(i, 0), v.R(i, 1), v.R(i, 2), v.R(i, 3)}, y[4];
        quat_mul(y, mc, x);      // quat = conj(mean) * quat
        quat_mul(x, d, y);       // quat = pert * quat
        quat_mul(y, mean, x);    // quat = mean * quat
        for (int c = 0; c < 4; ++c) v.R(i, c) = y[c];
    }
    balance_R(v);
}

// perturb(pf, PAR_T) + reCentre + balanceWeight(PAR_T)
THB_HD void perturb_T(const View& v, double pfac, double transS, double transQ, Rng& g)
{
    const double s0 = v.S(S_S0), s1 = v.S(S_S1), rho = v.S(S_RHO);
    for (int i = 0; i < v.mLT; ++i) {
        double x, y;
        g.bivariate_gaussian(s0, s1, rho / s0 / s1, x, y);
        v.T(i, 0) += x * pfac;
        v.T(i, 1) += y * pfac;
    }
    const double transM = transS * (-2.0 * log(transQ));   // transS * gsl_cdf_chisq_Qinv(transQ, 2)
    for (int i = 0; i < v.mLT; ++i)
        if (hypot(v.T(i, 0), v.T(i, 1)) > transM) {
            double x, y;
            g.bivariate_gaussian(transS, transS, 0.0, x, y);
            v.T(i, 0) = x;
            v.T(i, 1) = y;
        }
    balance_T(v);
}

THB_HD int argmax_uR(const View& v)
{
    int b = 0;
    for (int i = 1; i < v.mLR; ++i)
        if (v.UR(i) > v.UR(b)) b = i;
    return b;
}
THB_HD int argmax_uT(const View& v)
{
    int b = 0;
    for (int i = 1; i < v.mLT; ++i)
        if (v.UT(i) > v.UT(b)) b = i;
    return b;
}

// setUR/setUT from the E kernel's float marginals + keepHalfHeightPeak(PAR_R)
// (OPTIMISER_PEAK_FACTOR_R on, _T off: src/Optimiser.cpp:1408-1421)
THB_HD void set_u_keep_peak(const View& v, const float* uR, const float* uT)
{
    for (int i = 0; i < v.mLR; ++i) v.UR(i) = (double)uR[i];
    for (int i = 0; i < v.mLT; ++i) v.UT(i) = (double)uT[i];
    const double hh = v.UR(argmax_uR(v)) * v.S(S_PEAKR);
    for (int i = 0; i < v.mLR; ++i) v.UR(i) = v.UR(i) < hh ? 0.0 : v.UR(i) - hh;
}

THB_HD void rank1st(const View& v)
{
    const int a = argmax_uR(v), b = argmax_uT(v);
    for (int c = 0; c < 4; ++c) v.S(S_TOPR + c) = v.R(a, c);
    for (int c = 0; c < 2; ++c) v.S(S_TOPT + c) = v.T(b, c);
}

// calVari(PAR_R), MODE_3D
THB_HD void cal_vari_R(const View& v, Rng& g)
{
    (void)g.uniform_int((uint32_t)v.mLR);   // the reference draws an (unused under C1) anchor index
    double mean[4];
    acg_mean(v, mean);
    left_mul_all(v, mean, true);
    double A[16];
    infer_acg(v, A);
    v.S(S_K1) = A[5] / A[0];
    v.S(S_K2) = A[10] / A[0];
    v.S(S_K3) = A[15] / A[0];
    left_mul_all(v, mean, false);
}

// calVari(PAR_T): PARTICLE_RHO off -> rho = 0
THB_HD void cal_vari_T(const View& v)
{
    double m, sd;
    mean_sd(v, 0, m, sd); v.S(S_S0) = sd;
    mean_sd(v, 1, m, sd); v.S(S_S1) = sd;
    v.S(S_RHO) = 0.0;
}

THB_HD void cal_vari(const View& v, Rng& g)
{
    cal_vari_R(v, g);
    cal_vari_T(v);
}

// resample(n = mLR, PAR_R) and (mLT, PAR_T): shuffle, top, prior x likelihood, systematic resampling
THB_HD void resample_R(const View& v, Rng& g)
{
    const int n = v.mLR;
    // Particle::shuffle (src/Particle.cpp:2202-2300): gsl_ran_shuffle of the identity gives s, then new[s(i)] = old[i]
    for (int i = 0; i < n; ++i) v.W2(i) = (double)i;
    for (int i = n - 1; i > 0; --i) {
        const int j = (int)g.uniform_int((uint32_t)(i + 1));
        const double x = v.W2(i); v.W2(i) = v.W2(j); v.W2(j) = x;
    }
    for (int i = 0; i < n; ++i) {
        const int d = (int)v.W2(i);
        for (int c = 0; c < 4; ++c) v.R2(d, c) = v.R(i, c);
        v.W3(d) = v.WR(i);
        v.W4(d) = v.UR(i);
    }
    for (int i = 0; i < n; ++i) {
        for (int c = 0; c < 4; ++c) v.R(i, c) = v.R2(i, c);
        v.WR(i) = v.W3(i);
        v.UR(i) = v.W4(i);
    }
    const int top = argmax_uR(v);
    for (int c = 0; c < 4; ++c) v.S(S_TOPR + c) = v.R(top, c);
    double s = 0.0;
    for (int i = 0; i < n; ++i) { v.WR(i) *= v.UR(i); s += v.WR(i); }
    double cum = 0.0;
    for (int i = 0; i < n; ++i) { v.WR(i) /= s; cum += v.WR(i); v.W2(i) = cum; }   // W2 = cdf
    const double last = v.W2(n - 1);
    const double u0 = g.flat(0.0, 1.0 / n);
    int i = 0;
    for (int j = 0; j < n; ++j) {
        const double uj = u0 + j * 1.0 / n;
        while (i < n - 1 && uj > v.W2(i) / last) ++i;
        for (int c = 0; c < 4; ++c) v.R2(j, c) = v.R(i, c);
        v.WR(j) = -1.0 / v.UR(i);       // PARTICLE_PRIOR_ONE: prior 1/u; stored negated until the copy-back below
        // (WR(j) for j <= processed is no longer needed: cdf lives in W2; UR is read-only here)
    }
    for (int j = 0; j < n; ++j) {
        for (int c = 0; c < 4; ++c) v.R(j, c) = v.R2(j, c);
        v.WR(j) = -v.WR(j);
    }
}

THB_HD void resample_T(const View& v, Rng& g)
{
    const int n = v.mLT;
    for (int i = 0; i < n; ++i) v.W2(i) = (double)i;
    for (int i = n - 1; i > 0; --i) {
        const int j = (int)g.uniform_int((uint32_t)(i + 1));
        const double x = v.W2(i); v.W2(i) = v.W2(j); v.W2(j) = x;
    }
    for (int i = 0; i < n; ++i) {
        const int d = (int)v.W2(i);
        for (int c = 0; c < 2; ++c) v.T2(d, c) = v.T(i, c);
        v.W3(d) = v.WT(i);
        v.W4(d) = v.UT(i);
    }
    for (int i = 0; i < n; ++i) {
        for (int c = 0; c < 2; ++c) v.T(i, c) = v.T2(i, c);
        v.WT(i) = v.W3(i);
        v.UT(i) = v.W4(i);
    }
    const int top = argmax_uT(v);
    for (int c = 0; c < 2; ++c) v.S(S_TOPT + c) = v.T(top, c);
    double s = 0.0;
    for (int i = 0; i < n; ++i) { v.WT(i) *= v.UT(i); s += v.WT(i); }
    double cum = 0.0;
    for (int i = 0; i < n; ++i) { v.WT(i) /= s; cum += v.WT(i); v.W2(i) = cum; }
    const double last = v.W2(n - 1);
    const double u0 = g.flat(0.0, 1.0 / n);
    int i = 0;
    for (int j = 0; j < n; ++j) {
        const double uj = u0 + j * 1.0 / n;
        while (i < n - 1 && uj > v.W2(i) / last) ++i;
        for (int c = 0; c < 2; ++c) v.T2(j, c) = v.T(i, c);
        v.WT(j) = -1.0 / v.UT(i);
    }
    for (int j = 0; j < n; ++j) {
        for (int c = 0; c < 2; ++c) v.T(j, c) = v.T2(j, c);
        v.WT(j) = -v.WT(j);
    }
}

THB_HD double vari_R(const View& v) { return v.mode2D ? v.S(S_K1) : pow(v.S(S_K1) * v.S(S_K2) * v.S(S_K3), 1.0 / 6); }       // Particle::variR
THB_HD double vari_T(const View& v) { return sqrt(v.S(S_S0) * v.S(S_S0) * v.S(S_S1) * v.S(S_S1)); }   // rho = 0
THB_HD double compress_R(const View& v) { return v.mode2D ? 1.0 / v.S(S_K1) : pow(v.S(S_K1) * v.S(S_K2) * v.S(S_K3), -1.0 / 6); }

// Particle::load (src/Particle.cpp:401-556), in the reference's order of random draws: all ACG samples, then one sign per
// support point, balanceWeight / calVari of the rotations, the translations (one bivariate draw each), their balanceWeight /
// calVari, and the nD defocus factors (d + gaussian(s): drawn even when there is no CTF search, nD = 1); peak factor reset
// (PEAK_FACTOR_MIN = 1e-3).
THB_HD void load(const View& v, const double q[4], double k1, double k2, double k3, const double t[2], double s0,
                 double s1, Rng& g, int nD = 1, double sD = 0.0)
{
    v.S(S_K1) = k1; v.S(S_K2) = k2; v.S(S_K3) = k3;
    for (int c = 0; c < 4; ++c) v.S(S_TOPR + c) = q[c];
    sample_acg_r2(v, k1, k2, k3, g);
    for (int i = 0; i < v.mLR; ++i) {
        const double sgn = g.flat(-1.0, 1.0) >= 0 ? 1.0 : -1.0;
        const double qq[4] = {sgn * q[0], sgn * q[1], sgn * q[2], sgn * q[3]};
        const double d[4] = {v.R2(i, 0), v.R2(i, 1), v.R2(i, 2), v.R2(i, 3)};
        double y[4];
        quat_mul(y, d, qq);
        for (int c = 0; c < 4; ++c) v.R(i, c) = y[c];
        v.WR(i) = 1.0 / v.mLR;
        v.UR(i) = 1.0 / v.mLR;
    }
    for (int i = 0; i < v.mLT; ++i) { v.WT(i) = 1.0 / v.mLT; v.UT(i) = 1.0 / v.mLT; }
    balance_R(v);
    cal_vari_R(v, g);
    v.S(S_S0) = s0; v.S(S_S1) = s1;
    for (int c = 0; c < 2; ++c) v.S(S_TOPT + c) = t[c];
    for (int i = 0; i < v.mLT; ++i) {
        double x, y;
        g.bivariate_gaussian(s0, s1, 0.0, x, y);
        v.T(i, 0) = x + t[0];
        v.T(i, 1) = y + t[1];
    }
    balance_T(v);
    cal_vari_T(v);
    for (int i = 0; i < nD; ++i) (void)g.gaussian(sD);
    v.S(S_SCORE) = 1.0;
    v.S(S_NPHASE) = 0.0;
    v.S(S_VARIR) = 1.79769313486231570e308;
    v.S(S_VARIT) = 1.79769313486231570e308;
    v.S(S_VARID) = 1.79769313486231570e308;
    v.S(S_PEAKR) = 1e-3;
    v.S(S_NODEC) = 0.0;
    v.S(S_SPARE) = 0.0;
}

// ------------------------------------------------------------------------------------------------ the defocus dimension (CTF search)
// gsl_stats_mean / gsl_stats_sd_m in GSL's recurrence form (statistics/mean_source.c, variance_source.c)
THB_HD double mean_D(const View& v)
{
    double mean = 0.0;
    for (int i = 0; i < v.mLD; ++i) mean += (v.D(i) - mean) / (i + 1);
    return mean;
}
THB_HD double sd_m_D(const View& v, double mean)
{
    double variance = 0.0;
    for (int i = 0; i < v.mLD; ++i) {
        const double delta = v.D(i) - mean;
        variance += (delta * delta - variance) / (i + 1);
    }
    return sqrt(variance * ((double)v.mLD / (double)(v.mLD - 1)));
}

THB_HD void balance_D(const View& v);

// Particle::initD (src/Particle.cpp:280-310, PARTICLE_DEFOCUS_INIT_GAUSSIAN, PARTICLE_BALANCE_WEIGHT_D)
THB_HD void init_D(const View& v, double sD, Rng& g)
{
    for (int i = 0; i < v.mLD; ++i) v.D(i) = 1.0 + g.gaussian(sD);
    for (int i = 0; i < v.mLD; ++i) { v.WD(i) = 1.0 / v.mLD; v.UD(i) = 1.0 / v.mLD; }
    balance_D(v);
}

// balanceWeight(PAR_D) (src/Particle.cpp:2374-2409): wD = 1 / gaussian_pdf(d - mean, sd), then normW
THB_HD void balance_D(const View& v)
{
    const double m = mean_D(v);
    const double s = v.mLD == 1 ? 0.0 : sd_m_D(v, m);
    if (s == 0) {
        for (int i = 0; i < v.mLD; ++i) v.WD(i) = 1.0;
    } else {
        for (int i = 0; i < v.mLD; ++i) {
            const double u = (v.D(i) - m) / fabs(s);
            const double pdf = (1.0 / (sqrt(2.0 * 3.14159265358979323846) * fabs(s))) * exp(-u * u / 2.0);   // gsl_ran_gaussian_pdf
            v.WD(i) = 1.0 / pdf;
        }
    }
    norm_w(v);
}

// perturb(pf, PAR_D) (src/Particle.cpp:1278-1288)
THB_HD void perturb_D(const View& v, double pfac, Rng& g)
{
    const double s = v.S(S_SD);
    for (int i = 0; i < v.mLD; ++i) v.D(i) += g.gaussian(s) * pfac;
    balance_D(v);
}

// calVari(PAR_D) (src/Particle.cpp:1120-1141): gsl_stats_sd
THB_HD void cal_vari_D(const View& v) { v.S(S_SD) = v.mLD == 1 ? 0.0 : sd_m_D(v, mean_D(v)); }

// calRank1st(PAR_D) + resample(mLD, PAR_D) (src/Particle.cpp:1430-1475); W2 / W3 / W4 are long enough (max(mLR, mLT) >= mLD is
// required by the caller)
THB_HD void resample_D(const View& v, Rng& g)
{
    const int n = v.mLD;
    for (int i = 0; i < n; ++i) v.W2(i) = (double)i;
    for (int i = n - 1; i > 0; --i) {
        const int j = (int)g.uniform_int((uint32_t)(i + 1));
        const double x = v.W2(i); v.W2(i) = v.W2(j); v.W2(j) = x;
    }
    for (int i = 0; i < n; ++i) {
        const int d = (int)v.W2(i);
        v.R2(d, 0) = v.D(i);          // the rotation scratch doubles as the defocus scratch
        v.W3(d) = v.WD(i);
        v.W4(d) = v.UD(i);
    }
    for (int i = 0; i < n; ++i) { v.D(i) = v.R2(i, 0); v.WD(i) = v.W3(i); v.UD(i) = v.W4(i); }
    int top = 0;
    for (int i = 1; i < n; ++i)
        if (v.UD(i) > v.UD(top)) top = i;
    v.D(n) = v.D(top);                // _topD
    double s = 0.0;
    for (int i = 0; i < n; ++i) { v.WD(i) *= v.UD(i); s += v.WD(i); }
    double cum = 0.0;
    for (int i = 0; i < n; ++i) { v.WD(i) /= s; cum += v.WD(i); v.W2(i) = cum; }
    const double last = v.W2(n - 1);
    const double u0 = g.flat(0.0, 1.0 / n);
    int i = 0;
    for (int j = 0; j < n; ++j) {
        const double uj = u0 + j * 1.0 / n;
        while (i < n - 1 && uj > v.W2(i) / last) ++i;
        v.R2(j, 0) = v.D(i);
        v.W3(j) = 1.0 / v.UD(i);
    }
    for (int j = 0; j < n; ++j) { v.D(j) = v.R2(j, 0); v.WD(j) = v.W3(j); }
}

THB_HD void rank1st_D(const View& v)
{
    int top = 0;
    for (int i = 1; i < v.mLD; ++i)
        if (v.UD(i) > v.UD(top)) top = i;
    v.D(v.mLD) = v.D(top);
}

// the stop rule of the phase loop (src/Optimiser.cpp:1510-1615, OPTIMISER_COMPRESS_CRITERIA);
// returns true when the particle is finished.  variD is the constant defocus sigma (0 here).
THB_HD bool stop_rule(const View& v, int phase, int minPhase, double decreaseFactor, int noDecreaseLimit)
{
    if (phase < minPhase) return false;
    const double r = vari_R(v), t = vari_T(v), d = v.mLD > 0 ? v.S(S_SD) : 0.0;     // variD() = _s
    if (r < v.S(S_VARIR) * decreaseFactor || t < v.S(S_VARIT) * decreaseFactor || d < v.S(S_VARID) * decreaseFactor)
        v.S(S_NODEC) = 0.0;
    else
        v.S(S_NODEC) += 1.0;
    if (r < v.S(S_VARIR)) v.S(S_VARIR) = r;
    if (t < v.S(S_VARIT)) v.S(S_VARIT) = t;
    if (d < v.S(S_VARID)) v.S(S_VARID) = d;
    return v.S(S_NODEC) == (double)noDecreaseLimit;
}

}  // namespace pf
}  // namespace thb
