/*
 * thb_oracle.c - plain-C CPU restatement of the reference's algorithm for the Optimiser hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load this; nothing under thunder_b200/ links or calls it.  Each function cites the reference
 * file:line it follows (paths relative to the THUNDER tree).  It is pinned (tests/test_oracle.py)
 * against (i) oracle/_ref/libthunder_ref.so = the reference's own classes compiled from
 * /root/reference, run in this container, and (ii) the fixtures under tests/golden/ generated
 * from that library by tests/golden/make_golden.py.  The reference's own tests hold no golden
 * vectors for this path, so those two are the pin.
 *
 * Single precision build of the reference: RFLOAT = float, coordinates/quaternions double.
 * Compiled with -O2 and no -ffast-math / FMA contraction so that operation order is the
 * reference's.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#pragma STDC FP_CONTRACT OFF

typedef struct { float re, im; } cpx;

/* ---- a1: Optimiser::allocPreCalIdx, src/Optimiser.cpp:7991-8041; loop macro include/Image/Image.h:68-70 */
int orc_pixel_list(int N, int pf, float rU, float rL, int* iCol, int* iRow, int* iPxl, int* iSig, int* iColPad,
                   int* iRowPad)
{
    float rU2 = (float)((double)rU * (double)rU), rL2 = (float)((double)rL * (double)rL); /* TSGSL_pow_2 -> RFLOAT */
    float R = rU + 1;
    int nColFT = N / 2 + 1, n = 0;
    for (long j = (long)(-R); j < R; j++)
        for (long i = 0; i <= R; i++) {
            if (i == 0 && j < 0) continue;
            float u = (float)((double)i * (double)i + (double)j * (double)j); /* QUAD = gsl_pow_2 + gsl_pow_2 */
            if (u < rU2 && u >= rL2) {
                int v = (int)rint(hypot((double)i, (double)j)); /* AROUND(NORM(i, j)) */
                if (v < rU && v >= rL) {
                    if (iPxl) iPxl[n] = (int)((j >= 0 ? j : j + N) * nColFT + i); /* Image::iFTHalf */
                    if (iCol) iCol[n] = (int)i;
                    if (iRow) iRow[n] = (int)j;
                    if (iSig) iSig[n] = v;
                    if (iColPad) iColPad[n] = (int)i * pf;
                    if (iRowPad) iRowPad[n] = (int)j * pf;
                    n++;
                }
            }
        }
    return n;
}

/* ---- a10: rotate3D(dmat33&, const dvec4&), src/Geometry/Euler.cpp:181-189; column-major output */
void orc_rotate3D(const double* q, double* m)
{
    double A[3][3] = {{0, -q[3], q[2]}, {q[3], 0, -q[1]}, {-q[2], q[1], 0}};
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double aa = 0;
            for (int k = 0; k < 3; k++) aa += A[r][k] * A[k][c];
            m[c * 3 + r] = (r == c ? 1.0 : 0.0) + 2 * q[0] * A[r][c] + 2 * aa;
        }
}

/* ---- a3: translate(Complex*, tx, ty, ...), src/Image/ImageFunctions.cpp:233-252 and (dst, src) :471-492 */
void orc_translate(cpx* dst, const cpx* src, float tx, float ty, int N, const int* iCol, const int* iRow, int nPxl)
{
    float rCol = tx / N, rRow = ty / N;
    for (int i = 0; i < nPxl; i++) {
        float phase = (float)(6.28318530717959 * (iCol[i] * rCol + iRow[i] * rRow));
        float c = cosf(-phase), s = sinf(-phase); /* COMPLEX_POLAR(-phase), include/Complex.h:33 */
        if (src) {
            dst[i].re = src[i].re * c - src[i].im * s;
            dst[i].im = src[i].re * s + src[i].im * c;
        } else {
            dst[i].re = c;
            dst[i].im = s;
        }
    }
}

/* ---- CTF(RFLOAT* dst, ...), src/CTF.cpp:118-151 */
void orc_ctf(float* dst, float pixelSize, float voltage, float defocusU, float defocusV, float theta, float Cs,
             float amplitudeContrast, float phaseShift, int nCol, int nRow, const int* iCol, const int* iRow, int nPxl)
{
    float lambda = (float)(12.2643247 / sqrt(voltage * (1 + voltage * 0.978466e-6)));
    float w1 = sqrtf(1 - (float)((double)amplitudeContrast * (double)amplitudeContrast)); /* TS_SQRT(1 - TSGSL_pow_2) */
    float w2 = amplitudeContrast;
    float K1 = (float)(M_PI * lambda);
    float K2 = (float)(M_PI_2 * Cs * (float)((double)lambda * lambda * lambda)); /* TSGSL_pow_3 -> RFLOAT */
    for (int i = 0; i < nPxl; i++) {
        float u = (float)hypot(iCol[i] / (pixelSize * nCol), iRow[i] / (pixelSize * nRow));
        float angle = (float)(atan2((double)iRow[i], (double)iCol[i]) - theta);
        float defocus = -(defocusU + defocusV + (defocusU - defocusV) * cosf(2 * angle)) / 2;
        double u2d = (double)u * u;
        float u2 = (float)u2d, u4 = (float)(u2d * u2d); /* TSGSL_pow_2 / TSGSL_pow_4 return RFLOAT */
        float ki = K1 * defocus * u2 + K2 * u4 - phaseShift;
        dst[i] = -w1 * sinf(ki) + w2 * cosf(ki);
    }
}

/* ---- a5: Volume::getByInterpolationFT, src/Image/Volume.cpp:314-338; conjHalf include/Image/Volume.h:135-147;
 *      WG_TRI_INTERP_LINEAR include/Functions/Interpolation.h:163-200; getFTHalf(w, x0) Volume.cpp:491-563;
 *      iFTHalf include/Image/Volume.h:567-575 (per-corner wrap covers both the fast and the -1 path) */
static inline size_t ift_half(long i, long j, long k, int n, int nColFT)
{
    return (size_t)(k >= 0 ? k : k + n) * nColFT * n + (size_t)(j >= 0 ? j : j + n) * nColFT + i;
}

static inline int cell_setup(float x[3], long x0[3], float w[2][2][2])
{
    int conj = 0;
    if (!(x[0] >= 0)) { x[0] *= -1; x[1] *= -1; x[2] *= -1; conj = 1; }
    float xd[3], v[3][2];
    for (int a = 0; a < 3; a++) {
        x0[a] = (long)floorf(x[a]);
        xd[a] = x[a] - x0[a];
        v[a][0] = 1 - xd[a];
        v[a][1] = xd[a];
    }
    for (int k = 0; k < 2; k++)
        for (int j = 0; j < 2; j++)
            for (int i = 0; i < 2; i++) w[k][j][i] = v[0][i] * v[1][j] * v[2][k];
    return conj;
}

cpx orc_interp_ft(const cpx* vol, int n, float xx, float yy, float zz)
{
    float x[3] = {xx, yy, zz}, w[2][2][2];
    long x0[3];
    int nColFT = n / 2 + 1;
    int conj = cell_setup(x, x0, w);
    cpx r = {0, 0};
    for (int k = 0; k < 2; k++)
        for (int j = 0; j < 2; j++)
            for (int i = 0; i < 2; i++) {
                const cpx* v = vol + ift_half(x0[0] + i, x0[1] + j, x0[2] + k, n, nColFT);
                r.re += v->re * w[k][j][i];
                r.im += v->im * w[k][j][i];
            }
    if (conj) r.im = -r.im;
    return r;
}

/* ---- a4: Projector::project(Complex*, const dmat33&, iCol, iRow, nPxl, nThread), src/Projector.cpp:356-374 */
void orc_project(cpx* dst, const cpx* vol, int n, int pf, const double* mat, const int* iCol, const int* iRow, int nPxl)
{
    for (int i = 0; i < nPxl; i++) {
        double a = (double)(iCol[i] * pf), b = (double)(iRow[i] * pf);
        double ox = mat[0] * a + mat[3] * b + mat[6] * 0.0;
        double oy = mat[1] * a + mat[4] * b + mat[7] * 0.0;
        double oz = mat[2] * a + mat[5] * b + mat[8] * 0.0;
        dst[i] = orc_interp_ft(vol, n, (float)ox, (float)oy, (float)oz);
    }
}

/* ---- a6: logDataVSPrior_m_huabin, src/Optimiser.cpp:9187-9213 (scalar form) */
float orc_logDataVSPrior(const cpx* dat, const cpx* pri, const float* ctf, const float* sigRcp, int m)
{
    float result = 0;
    for (int i = 0; i < m; i++) {
        float tr = ctf[i] * pri[i].re, ti = ctf[i] * pri[i].im;
        float dr = dat[i].re - tr, di = dat[i].im - ti;
        float t2 = dr * dr + di * di;
        result += t2 * sigRcp[i];
    }
    return result;
}

/* ---- a7: logDataVSPrior_m_n_huabin, src/Optimiser.cpp:9931-9973: n images (pixel-major) against one template */
void orc_logDataVSPrior_m_n(const cpx* dat, const cpx* pri, const float* ctf, const float* sigRcp, int n, int m,
                            float* result)
{
    for (int l = 0; l < n; l++) result[l] = 0;
    for (int i = 0; i < m; i++)
        for (int l = 0; l < n; l++) {
            size_t k = (size_t)i * n + l;
            float tr = ctf[k] * pri[i].re, ti = ctf[k] * pri[i].im;
            float dr = dat[k].re - tr, di = dat[k].im - ti;
            result[l] += (dr * dr + di * di) * sigRcp[k];
        }
}

/* ---- a8 + the per-image body of the phase loop, src/Optimiser.cpp:1216-1402 (k = 1, no CTF search):
 *      translate -> project -> priAllP = traP * priRotP -> logDataVSPrior -> running-baseline weights.
 *      quat[nR][4], tran[nT][2], priors wR[nR], wT[nT] (wC = wD = 1).
 *      Outputs uR[nR], uT[nT], uC, base, logL[nR][nT] (any may be NULL). */
void orc_expect_local(const cpx* vol, int n, int pf, int N, const int* iCol, const int* iRow, int nPxl, const cpx* dat,
                      const float* ctf, const float* sigRcp, int nR, int nT, const double* quat, const double* tran,
                      const double* wR, const double* wT, float* uR, float* uT, float* uC, float* base, float* logL)
{
    cpx* traP = (cpx*)malloc(sizeof(cpx) * (size_t)nT * nPxl);
    cpx* priRotP = (cpx*)malloc(sizeof(cpx) * nPxl);
    cpx* priAllP = (cpx*)malloc(sizeof(cpx) * nPxl);
    float* aR = (float*)calloc(nR, sizeof(float));
    float* aT = (float*)calloc(nT, sizeof(float));
    float aC = 0, baseLine = NAN;
    for (int t = 0; t < nT; t++)
        orc_translate(traP + (size_t)t * nPxl, NULL, (float)tran[2 * t], (float)tran[2 * t + 1], N, iCol, iRow, nPxl);
    for (int r = 0; r < nR; r++) {
        double mat[9];
        orc_rotate3D(quat + 4 * r, mat);
        orc_project(priRotP, vol, n, pf, mat, iCol, iRow, nPxl);
        for (int t = 0; t < nT; t++) {
            const cpx* tr = traP + (size_t)t * nPxl;
            for (int i = 0; i < nPxl; i++) { /* Complex operator*, include/Complex.h */
                priAllP[i].re = tr[i].re * priRotP[i].re - tr[i].im * priRotP[i].im;
                priAllP[i].im = tr[i].re * priRotP[i].im + tr[i].im * priRotP[i].re;
            }
            float w = orc_logDataVSPrior(dat, priAllP, ctf, sigRcp, nPxl);
            if (logL) logL[(size_t)r * nT + t] = w;
            if (isnan(baseLine)) baseLine = w;
            if (w > baseLine) {
                float nf = expf(baseLine - w);
                aC *= nf;
                for (int k = 0; k < nR; k++) aR[k] *= nf;
                for (int k = 0; k < nT; k++) aT[k] *= nf;
                baseLine = w;
            }
            float s = expf(w - baseLine);
            aC = (float)(aC + s * (wR[r] * wT[t] * 1.0));
            aR[r] = (float)(aR[r] + s * (1.0 * wT[t] * 1.0));
            aT[t] = (float)(aT[t] + s * (1.0 * wR[r] * 1.0));
        }
    }
    if (uR) memcpy(uR, aR, sizeof(float) * nR);
    if (uT) memcpy(uT, aT, sizeof(float) * nT);
    if (uC) *uC = aC;
    if (base) *base = baseLine;
    free(traP); free(priRotP); free(priAllP); free(aR); free(aT);
}

/* ---- a12/a13: Reconstructor::insertP (src/Reconstructor.cpp:782-863, sig == NULL) with Volume::addFT
 *      (src/Image/Volume.cpp:340-375, 565-712).  F complex half-volume, T real (the reference keeps a
 *      complex T whose real part is used).  iColPad/iRowPad already multiplied by pf. */
void orc_insertP(cpx* F, float* T, int n, const cpx* src, const float* ctf, const double* mat, float w, const int* iColPad,
                 const int* iRowPad, int nPxl)
{
    int nColFT = n / 2 + 1;
    for (int p = 0; p < nPxl; p++) {
        double ox = mat[0] * iColPad[p] + mat[3] * iRowPad[p];
        double oy = mat[1] * iColPad[p] + mat[4] * iRowPad[p];
        double oz = mat[2] * iColPad[p] + mat[5] * iRowPad[p];
        cpx val;
        val.re = src[p].re * ctf[p] * 1 * w;
        val.im = src[p].im * ctf[p] * 1 * w;
        float tval = (float)((double)ctf[p] * (double)ctf[p]) * 1 * w; /* TSGSL_pow_2(ctf) -> RFLOAT */
        float x[3] = {(float)ox, (float)oy, (float)oz}, wg[2][2][2];
        long x0[3];
        if (cell_setup(x, x0, wg)) val.im = -val.im;
        for (int k = 0; k < 2; k++)
            for (int j = 0; j < 2; j++)
                for (int i = 0; i < 2; i++) {
                    size_t idx = ift_half(x0[0] + i, x0[1] + j, x0[2] + k, n, nColFT);
                    F[idx].re += val.re * wg[k][j][i];
                    F[idx].im += val.im * wg[k][j][i];
                    T[idx] += tval * wg[k][j][i];
                }
    }
}

/* ---- a11 + a14: insert loop of Optimiser::reconstructRef (src/Optimiser.cpp:7036-7241) with explicit
 *      draws nr[nImg][mReco][4], nt[nImg][mReco][2]; insertDir (src/Reconstructor.cpp:407-422).
 *      iCol/iRow: unpadded (for translate); the padded ones are iCol*pf. */
void orc_insert_loop(cpx* F, float* T, double* O, int* counter, int n, int pf, int N, const cpx* dat, const float* ctf,
                     const float* w, const double* offS, const double* nr, const double* nt, int nImg, int mReco,
                     const int* iCol, const int* iRow, int nPxl)
{
    int* a = (int*)malloc(sizeof(int) * nPxl);
    int* b = (int*)malloc(sizeof(int) * nPxl);
    cpx* tmp = (cpx*)malloc(sizeof(cpx) * nPxl);
    for (int i = 0; i < nPxl; i++) { a[i] = iCol[i] * pf; b[i] = iRow[i] * pf; }
    for (int l = 0; l < nImg; l++)
        for (int m = 0; m < mReco; m++) {
            const double* q = nr + ((size_t)l * mReco + m) * 4;
            const double* t = nt + ((size_t)l * mReco + m) * 2;
            double tx = t[0] - (offS ? offS[2 * l] : 0), ty = t[1] - (offS ? offS[2 * l + 1] : 0);
            double mat[9];
            orc_rotate3D(q, mat);
            orc_translate(tmp, dat + (size_t)l * nPxl, (float)-tx, (float)-ty, N, iCol, iRow, nPxl);
            orc_insertP(F, T, n, tmp, ctf + (size_t)l * nPxl, mat, w[l], a, b, nPxl);
            O[0] += -(mat[0] * tx + mat[3] * ty);
            O[1] += -(mat[1] * tx + mat[4] * ty);
            O[2] += -(mat[2] * tx + mat[5] * ty);
            *counter += 1;
        }
    free(a); free(b); free(tmp);
}

/* ---- a15: the normalisation inside allReduceT, src/Reconstructor.cpp:2458-2483 (sf = 1 / Re T[0]) */
void orc_normalise_TF(cpx* F, float* T, size_t nVox)
{
    float sf = (float)(1.0 / T[0]);
    for (size_t i = 0; i < nVox; i++) { T[i] *= sf; F[i].re *= sf; F[i].im *= sf; }
}
